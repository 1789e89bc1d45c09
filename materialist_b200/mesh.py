"""Device-side triangle mesh of the render operator's mesh mode — the replacement for Mitsuba's `ply` shape + OptiX
acceleration structure in the scene of inverse_img_w_mi.py:40-56 / render_final.py:32-53.

`Mesh(verts, tris)` uploads the index/vertex buffers and runs mb200_mesh_build (bounds, Morton radix sort, angle-weighted
vertex normals, implicit 4-ary BVH — all on the GPU, current stream).  The mesh is static over an optimisation, so it
is built once per scene.
"""
import ctypes as C

import numpy as np
import torch

from . import _abi


def read_ply_mesh(path):
    """(verts (nv,3) float32, tris (nt,3) int32) of a binary_little_endian PLY with triangle faces (what Open3D writes
    at inverse_img_w_mi.py:721-727)."""
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
        types = {"double": "<f8", "float": "<f4", "uchar": "u1", "int": "<i4", "uint": "<u4", "float32": "<f4", "float64": "<f8",
                 "int32": "<i4", "uint32": "<u4", "uint8": "u1", "short": "<i2", "ushort": "<u2", "char": "i1"}
        nv = nf = 0; vprops = []; cur = None; face_list = None
        for l in text:
            t = l.split()
            if t[:2] == ["element", "vertex"]:
                nv, cur = int(t[2]), "vertex"
            elif t[:2] == ["element", "face"]:
                nf, cur = int(t[2]), "face"
            elif t[:1] == ["element"]:
                cur = None
            elif t[:1] == ["property"] and cur == "vertex":
                vprops.append((t[2], types[t[1]]))
            elif t[:2] == ["property", "list"] and cur == "face":
                face_list = (types[t[2]], types[t[3]])
        vd = np.dtype(vprops)
        v = np.frombuffer(f.read(nv * vd.itemsize), dtype=vd, count=nv)
        verts = np.stack([v["x"], v["y"], v["z"]], 1).astype(np.float32)
        if nf == 0 or face_list is None:
            return verts, np.zeros((0, 3), np.int32)
        fd = np.dtype([("n", face_list[0]), ("i", face_list[1], (3,))])
        fc = np.frombuffer(f.read(nf * fd.itemsize), dtype=fd, count=nf)
        if not (fc["n"] == 3).all():
            raise ValueError("PLY: only triangle faces are supported")
    return np.ascontiguousarray(verts), np.ascontiguousarray(fc["i"].astype(np.int32))


class Mesh:
    def __init__(self, verts, tris, face_normals=False, device="cuda"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("materialist_b200 meshes live on a CUDA device (no CPU fallback)")
        verts = torch.as_tensor(np.ascontiguousarray(verts, dtype=np.float32) if not isinstance(verts, torch.Tensor) else verts)
        tris = torch.as_tensor(np.ascontiguousarray(tris, dtype=np.int32) if not isinstance(tris, torch.Tensor) else tris)
        if verts.ndim != 2 or verts.shape[1] != 3 or tris.ndim != 2 or tris.shape[1] != 3:
            raise ValueError("verts must be (nv,3) and tris (nt,3)")
        if tris.shape[0] == 0:
            raise ValueError("mesh has no triangles")
        if int(tris.min()) < 0 or int(tris.max()) >= verts.shape[0]:
            raise ValueError("triangle index out of range")
        self.verts = verts.to(self.device, torch.float32).contiguous()
        self.tris = tris.to(self.device, torch.int32).contiguous()
        self.nv, self.nt = int(verts.shape[0]), int(tris.shape[0])
        self.desc = _abi.MeshDesc()
        _abi.check(_abi.lib.mb200_mesh_describe(self.nv, self.nt, int(bool(face_normals)), C.byref(self.desc)), "mb200_mesh_describe")
        with torch.cuda.device(self.device):
            sbytes = _abi.lib.mb200_mesh_scratch_bytes(self.nv, self.nt, int(bool(face_normals)))
            if sbytes == 0:
                raise _abi.MB200Error("mb200_mesh_scratch_bytes failed: " + _abi.lib.mb200_last_cuda_error().decode())
            self.buf = torch.empty(self.desc.total_bytes // 8 + 1, dtype=torch.float64, device=self.device)
            scratch = torch.empty(sbytes // 8 + 1, dtype=torch.float64, device=self.device)
            _abi.check(_abi.lib.mb200_mesh_build(_abi.ptr(self.verts), _abi.ptr(self.tris), C.byref(self.desc), _abi.ptr(self.buf),
                                                 _abi.ptr(scratch), _abi.stream_ptr()), "mb200_mesh_build")
            scratch.record_stream(torch.cuda.current_stream())

    def header(self):
        """(centre.xyz, radius, bbox lo, bbox hi) as written by the build."""
        h = self.buf.view(torch.float32)[self.desc.off_header // 4: self.desc.off_header // 4 + 12].cpu().numpy()
        return h[:3], float(h[3]), h[4:7], h[8:11]

    def corner_normals(self):
        """(tri_id (n_slots,), normals (n_slots, 3, 3)) as stored for the shading frames: per Morton-ordered slot the original triangle id
        (-1: padding) and the angle-weighted vertex normal of each of its three corners (None with face normals)."""
        if self.desc.face_normals:
            return None
        n = self.desc.n_slots
        f = self.buf.view(torch.float32)
        tv = f[self.desc.off_tv // 4: self.desc.off_tv // 4 + n * 12].view(n, 3, 4)
        tn = f[self.desc.off_tn // 4: self.desc.off_tn // 4 + n * 12].view(n, 3, 4)
        return tv[:, 0, 3].contiguous().view(torch.int32), tn[..., :3].contiguous()

    def intersect(self, o, d, maxt=None, any_hit=False):
        o = torch.as_tensor(o, dtype=torch.float32).to(self.device).contiguous()
        d = torch.as_tensor(d, dtype=torch.float32).to(self.device).contiguous()
        n = o.shape[0]
        mt = None if maxt is None else torch.as_tensor(maxt, dtype=torch.float32).to(self.device).contiguous()
        tri = torch.empty(n, dtype=torch.int32, device=self.device); tuv = torch.empty(n, 3, device=self.device)
        _abi.check(_abi.lib.mb200_mesh_intersect(C.byref(self.desc), _abi.ptr(self.buf), _abi.ptr(o), _abi.ptr(d), _abi.ptr(mt), n,
                                                 int(any_hit), _abi.ptr(tri), _abi.ptr(tuv), _abi.stream_ptr()), "mb200_mesh_intersect")
        return tri, tuv

    def primary(self, cfg, jx=0.5, jy=0.5):
        """Primary visibility through film offset (jx, jy) of each pixel: gpos (H,W,4), gnrm (H,W,4), tri (H,W), flat (H,W)."""
        H, W = cfg.H, cfg.W
        gpos = torch.empty(H, W, 4, device=self.device); gnrm = torch.empty(H, W, 4, device=self.device)
        tri = torch.empty(H, W, dtype=torch.int32, device=self.device); flat = torch.empty(H, W, dtype=torch.int32, device=self.device)
        _abi.check(_abi.lib.mb200_mesh_primary(C.byref(cfg), C.byref(self.desc), _abi.ptr(self.buf), float(jx), float(jy), _abi.ptr(gpos),
                                               _abi.ptr(gnrm), _abi.ptr(tri), _abi.ptr(flat), _abi.stream_ptr()), "mb200_mesh_primary")
        return gpos, gnrm, tri, flat
