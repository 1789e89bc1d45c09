"""G-buffer extraction: replaces the ray-traced primary visibility of the reference's depth-derived height-field PLY
(myutils/mesh_recon.py:41-74, written at inverse_img_w_mi.py:721-727) by the per-pixel (position, normal, valid)
buffers the fused kernels shade.

The reference's mesh has vertex k = row*W + col sitting exactly on pixel (col,row)'s centre ray (SURVEY §2 #13,
[PROBE]: reprojection error 1e-13 px), followed by extra 'curtain' vertices; so the first H*W vertices ARE the
position buffer.  Normals are the normalised cross product of the central differences of neighbouring vertices,
flipped towards the camera (the reference shades flat triangle normals; per-triangle primary hits are the §8f
'next' row).  Small readers for the file formats involved live here (binary PLY; Radiance .hdr / OpenEXR via imageio)
too so that the shipped scenes (output_imgs/*) can be relit without Mitsuba's mi.Bitmap.
"""
import os

import numpy as np

from .scene import Camera


def read_ply_vertices(path):
    """Vertices (n,3) float64 of an ASCII-header PLY with binary_little_endian body (what Open3D writes)."""
    with open(path, "rb") as f:
        header = b""
        while not header.endswith(b"end_header\n"):
            line = f.readline()
            if not line:
                raise ValueError("PLY: end_header not found")
            header += line
        text = header.decode("ascii", "replace").splitlines()
        if not any(l.strip() == "format binary_little_endian 1.0" for l in text):
            raise ValueError("PLY: only binary_little_endian 1.0 is supported")
        n, props, in_vertex = 0, [], False
        for l in text:
            t = l.split()
            if t[:2] == ["element", "vertex"]:
                n, in_vertex = int(t[2]), True
            elif t[:1] == ["element"]:
                in_vertex = False
            elif t[:1] == ["property"] and in_vertex:
                props.append((t[2], {"double": "<f8", "float": "<f4", "uchar": "u1", "int": "<i4", "uint": "<u4"}[t[1]]))
        data = np.frombuffer(f.read(n * np.dtype(props).itemsize), dtype=np.dtype(props), count=n)
    return np.stack([data["x"], data["y"], data["z"]], 1).astype(np.float64)


def gbuffer_from_positions(pos, camera=None, depth_eps=1e-6):
    """pos (H,W,3) world positions -> (pos, nrm, valid) float32/bool. Pixels at the camera origin are invalid."""
    H, W, _ = pos.shape
    cam = camera or Camera(width=W, height=H)
    p = pos.astype(np.float64)
    dx = np.zeros_like(p); dy = np.zeros_like(p)
    dx[:, 1:-1] = p[:, 2:] - p[:, :-2]; dx[:, 0] = p[:, 1] - p[:, 0]; dx[:, -1] = p[:, -1] - p[:, -2]
    dy[1:-1] = p[2:] - p[:-2]; dy[0] = p[1] - p[0]; dy[-1] = p[-1] - p[-2]
    n = np.cross(dx, dy)
    norm = np.linalg.norm(n, axis=-1, keepdims=True)
    valid = (norm[..., 0] > 0) & (np.linalg.norm(p - cam.to_world[:3, 3], axis=-1) > depth_eps) & np.isfinite(p).all(-1)
    n = n / np.maximum(norm, 1e-30)
    flip = (n * (cam.to_world[:3, 3] - p)).sum(-1) < 0
    n[flip] = -n[flip]
    n[~valid] = (0, 0, 1)
    return p.astype(np.float32), n.astype(np.float32), valid


def gbuffer_from_ply(path, H=512, W=512, camera=None):
    v = read_ply_vertices(path)
    if v.shape[0] < H * W:
        raise ValueError(f"PLY has {v.shape[0]} vertices, expected at least {H * W} (vertex k <-> pixel k)")
    return gbuffer_from_positions(v[:H * W].reshape(H, W, 3), camera)


def read_image(path):
    """Radiance .hdr / OpenEXR -> float32 RGB(A) array in linear file values through the library's own readers
    (imageio.read_bitmap, csrc/mb200_io.cu), PNG (bg.png / mask.png of the editing scripts) -> [0, 1] floats the same way;
    anything else through OpenCV."""
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    if path.lower().endswith((".hdr", ".exr", ".png")):
        from .imageio import read_bitmap
        return read_bitmap(path)
    import cv2
    img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    if img is None:
        raise FileNotFoundError(path)
    if img.ndim == 3:
        img = img[..., [2, 1, 0] + ([3] if img.shape[2] == 4 else [])]
    if img.dtype == np.uint8:
        return img.astype(np.float32) / 255.0
    if img.dtype == np.uint16:
        return img.astype(np.float32) / 65535.0
    return img.astype(np.float32)


def load_estimated_brdf(mat_dir):
    """mi_plugin.py:701-739: albedo / roughness / metallic / normal EXRs of a `best_results` folder; note the
    reference's `roughness * 0.95 + 0.05` (:716)."""
    out = {"albedo": read_image(os.path.join(mat_dir, "albedo.exr"))[..., :3]}
    r = read_image(os.path.join(mat_dir, "roughness.exr")); m = read_image(os.path.join(mat_dir, "metallic.exr"))
    out["roughness"] = (r[..., :1] if r.ndim == 3 else r[..., None]) * 0.95 + 0.05
    out["metallic"] = m[..., :1] if m.ndim == 3 else m[..., None]
    npath = os.path.join(mat_dir, "normal.exr")
    if os.path.exists(npath):
        out["normal"] = read_image(npath)[..., :3]
    out = {k: np.ascontiguousarray(v, dtype=np.float32) for k, v in out.items()}
    # the editing inputs (mi_plugin.py:717-735): background image (resized to the maps, bilinear / align_corners like the
    # reference), edit mask (first channel as bool), optimised envmap
    bpath, mpath, epath = (os.path.join(mat_dir, n) for n in ("bg.png", "mask.png", "envmap.hdr"))
    if os.path.exists(bpath):
        bg = read_image(bpath)[..., :3]
        H, W = out["albedo"].shape[:2]
        if bg.shape[0] != H:
            import torch
            import torch.nn.functional as NF
            bg = NF.interpolate(torch.from_numpy(np.ascontiguousarray(bg))[None].permute(0, 3, 1, 2), size=(H, W), mode="bilinear",
                                align_corners=True)[0].permute(1, 2, 0).numpy()
        out["bg"] = np.ascontiguousarray(bg, dtype=np.float32)
    if os.path.exists(mpath):
        mk = read_image(mpath)
        out["mask"] = np.ascontiguousarray((mk[..., 0] if mk.ndim == 3 else mk) != 0)
    if os.path.exists(epath):
        out["envmap"] = np.ascontiguousarray(read_image(epath)[..., :3], dtype=np.float32)
    return out
