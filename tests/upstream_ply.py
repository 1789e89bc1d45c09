"""Binary little-endian PLY writer in the layout of the reference's depth-derived meshes (double vertices, uchar-uint face lists)."""
import numpy as np


def write_ply(path, verts, tris):
    verts = np.asarray(verts); tris = np.asarray(tris)
    with open(path, "wb") as f:
        f.write((f"ply\nformat binary_little_endian 1.0\nelement vertex {len(verts)}\nproperty double x\nproperty double y\nproperty double z\n"
                 f"element face {len(tris)}\nproperty list uchar uint vertex_indices\nend_header\n").encode())
        f.write(verts.astype("<f8").tobytes())
        rec = np.empty(len(tris), dtype=np.dtype([("n", "u1"), ("i", "<u4", 3)]))
        rec["n"] = 3; rec["i"] = tris
        f.write(rec.tobytes())
