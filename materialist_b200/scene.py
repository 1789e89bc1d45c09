"""Scene handle of the B200 render operator — the replacement for the `mi.load_dict({...})` scene of
inverse_img_w_mi.py:30-56 / render_final.py:19-97 and for `mi.traverse(scene)`.

A scene is: a per-pixel G-buffer (position, geometric normal, validity) in place of the ray-traced PLY
height field, the camera meta MatDiffBSDF reads (myutils/default_cam.json, mi_plugin.py:1259-1275), the
material maps a/r/m/n exposed by MatDiffBSDF.traverse (mi_plugin.py:1464-1469), the envmap emitter data,
and the integrator / film settings (path max_depth, hdrfilm gaussian rfilter — all Mitsuba defaults).
"""
import ctypes as C
import json
import math
import os

import numpy as np
import torch

from . import _abi

from .camera import Camera  # noqa: F401  (re-exported: `from materialist_b200.scene import Camera` keeps working)


class TransSettings:
    """State of the `TransBSDF` plugin (myutils/mi_plugin.py:1477-1492): ior (default 1.3), specTrans (0.8), the background
    image `bg` (0.5) and the edit `mask` (empty), and refract_distance = 100 when the plugin was created with a
    keep_albedo_color property ("scale factor for real scene"), else 1."""

    def __init__(self, H, W, device, ior=1.3, keep_albedo_color=None):
        self.ior = float(ior)
        self.keep_albedo_color = bool(keep_albedo_color) if keep_albedo_color is not None else False
        self.refract_distance = 100.0 if keep_albedo_color is not None else 1.0
        self.spec_trans = 0.8
        self.bg = torch.full((H, W, 3), 0.5, device=device)
        self.mask = torch.zeros(H, W, dtype=torch.uint8, device=device)

    def desc(self):
        return _abi.Trans(self.ior, self.spec_trans, self.refract_distance, 0, self.bg.data_ptr(), self.mask.data_ptr())


class Scene:
    """G-buffer scene handle. Tensors live on one CUDA device; everything is fp32 and contiguous."""

    def __init__(self, pos, nrm, valid=None, camera=None, envmap=None, use_mesh_normal=True, max_depth=4,
                 rfilter="gaussian", device=None, flags=None, mesh=None):
        """mesh: optional materialist_b200.mesh.Mesh.  With a mesh the operator TRACES it (per-sample triangle hits,
        shadow rays, bounces — what the reference renders); without, it shades the per-pixel G-buffer (pos, nrm)."""
        pos = torch.as_tensor(pos, dtype=torch.float32)
        nrm = torch.as_tensor(nrm, dtype=torch.float32)
        if pos.ndim != 3 or pos.shape[-1] != 3 or nrm.shape != pos.shape:
            raise ValueError("pos and nrm must be (H, W, 3)")
        H, W, _ = pos.shape
        self.device = torch.device(device if device is not None else "cuda")
        if self.device.type != "cuda":
            raise ValueError("materialist_b200 scenes live on a CUDA device (no CPU fallback)")
        self.H, self.W = H, W
        self.camera = camera or Camera(width=W, height=H)
        if (self.camera.width, self.camera.height) != (W, H):
            raise ValueError("camera film size does not match the G-buffer")
        v = torch.ones(H, W, 1) if valid is None else torch.as_tensor(valid).reshape(H, W, 1).float()
        self.gpos = torch.cat([pos.to(self.device), v.to(self.device)], -1).contiguous()
        self.gnrm = torch.cat([nrm.to(self.device), torch.zeros(H, W, 1, device=self.device)], -1).contiguous()
        self.mesh = mesh
        if mesh is not None and mesh.buf.device != self.gpos.device:
            raise ValueError("mesh and scene must live on the same device")
        # MatDiffBSDF placeholders (mi_plugin.py:1238-1241)
        self.a = torch.full((H, W, 3), 0.5, device=self.device)
        self.r = torch.full((H, W, 1), 0.5, device=self.device)
        self.m = torch.full((H, W, 1), 0.5, device=self.device)
        self.n = torch.full((H, W, 3), 0.5, device=self.device)
        self.use_mesh_normal = bool(use_mesh_normal)
        self.max_depth = int(max_depth)
        if rfilter not in ("gaussian", "box"):
            raise ValueError("rfilter must be 'gaussian' or 'box'")
        self.filter = _abi.FILTER_GAUSSIAN if rfilter == "gaussian" else _abi.FILTER_BOX
        if flags is None:   # reference-exact defaults; the row-stride quirk is only meaningful (and harmless) for square maps
            flags = _abi.FLAG_WO_WORLD_QUIRK | _abi.FLAG_ENV_HALF_TEXEL | (_abi.FLAG_ROW_STRIDE_H if H == W else 0)
        self.flags = int(flags)
        self.row0, self.rows = 0, H            # pixel shard (whole image by default)
        self.trans = None                      # TransSettings when the shape's BSDF is the TransBSDF plugin (set_bsdf)
        # forward formulation of mesh mode: "wavefront" (mb200_mesh_shade_fwd_wf: traversal and shading in separate kernels, path
        # state in a scratch buffer; the faster one) or "persistent" (mb200_mesh_shade_fwd: one kernel, no scratch memory)
        self.mesh_forward = os.environ.get("MB200_MESH_FORWARD", "wavefront")
        self.mesh_backward = os.environ.get("MB200_MESH_BACKWARD", "wavefront")    # mb200_mesh_shade_bwd_wf / "persistent": mb200_mesh_shade_bwd
        self._wf_scratch = None
        self._env = None
        if envmap is None:
            envmap = torch.ones(16, 32, 3)
        self.set_envmap(envmap, _abi.ENV_FILE)

    @classmethod
    def from_mesh(cls, verts, tris, camera, face_normals=False, device="cuda", trace=True, **kw):
        """Scene for a triangle mesh (the reference's depth-derived PLY).  The G-buffer of the fast path is the mesh's
        primary visibility through the pixel centres (mb200_mesh_primary); `trace=False` drops the mesh afterwards
        (G-buffer shading only)."""
        from .mesh import Mesh
        mesh = Mesh(verts, tris, face_normals=face_normals, device=device)
        c = _abi.Cfg()
        c.H, c.W = camera.height, camera.width
        c.view[:] = camera.view_matrix.reshape(-1).tolist()
        c.proj[:] = camera.proj_matrix.reshape(-1).tolist()
        c.cam_to_world[:] = camera.to_world.astype(np.float32).reshape(-1).tolist()
        c.tan_half_fov_x = camera.tan_half_fov_x
        # slightly off-centre: rays through the exact pixel centres pass through the mesh vertices (vertex k <-> pixel k) and
        # ~7 % of them slip between the triangles in float
        gpos, gnrm, _, _ = mesh.primary(c, 0.47, 0.53)
        return cls(gpos[..., :3], gnrm[..., :3], gpos[..., 3:], camera=camera, device=device, mesh=mesh if trace else None, **kw)

    # ---------------------------------------------------------------- BSDF plugin
    def set_bsdf(self, bsdf=None):
        """The `bsdf=` argument of render_final.load_estimated_mesh_w_env (render_final.py:19-97): {'name': 'matDiffBSDF'} (default)
        or {'name': 'TransBSDF', 'ior': ..., 'keep_albedo_color': ...} as trans_edit.py:18 passes it.  TransBSDF is forward only."""
        name = (bsdf or {}).get("name", "matDiffBSDF")
        if name in ("matDiffBSDF", "matbsdf"):
            self.trans = None
        elif name == "TransBSDF":
            self.trans = TransSettings(self.H, self.W, self.device, ior=bsdf.get("ior", 1.3), keep_albedo_color=bsdf.get("keep_albedo_color"))
        else:
            raise ValueError(f"unsupported bsdf {name!r}: this operator implements matDiffBSDF and TransBSDF")
        return self

    # ---------------------------------------------------------------- shard
    def set_shard(self, row0, rows):
        if row0 < 0 or rows <= 0 or row0 + rows > self.H:
            raise ValueError("shard rows out of range")
        self.row0, self.rows = int(row0), int(rows)

    def shard(self, row0, rows):
        """Context manager: render rows [row0, row0 + rows) inside the block, restore the previous shard on exit (the optimisers
        use it per step, so a Scene shared between phases is never left sharded)."""
        scene = self

        class _Shard:
            def __enter__(self_):
                self_.prev = (scene.row0, scene.rows)
                scene.set_shard(row0, rows)
                return scene

            def __exit__(self_, *exc):
                scene.row0, scene.rows = self_.prev
                return False
        return _Shard()

    # ---------------------------------------------------------------- envmap
    def set_envmap(self, env, mode):
        """env: (He, We, 3). mode ENV_FILE = bitmap loaded from file (column appended), ENV_ASSIGNED = tensor
        assigned through params['emitter.data'] (first/last column averaged)."""
        env = torch.as_tensor(env, dtype=torch.float32)
        if env.ndim != 3 or env.shape[-1] != 3 or env.shape[0] < 2 or env.shape[1] < 2:
            raise ValueError("envmap must be (He, We, 3) with He, We >= 2")
        self.env_user = env.detach().contiguous().to(self.device)
        self.env_mode = mode
        self._env = None                        # hierarchy rebuilt lazily on the next render

    def prepared_env(self, env_tensor=None, mode=None):
        """Returns (env4, hier, desc, He, We, mode) for the current (or the given) envmap; runs the ingest +
        Hierarchical2D build kernels on the current stream."""
        cache = env_tensor is None
        if cache:
            if self._env is not None:
                return self._env
            env_tensor, mode = self.env_user, self.env_mode
        env_tensor = env_tensor.detach().contiguous().float()
        He, We, _ = env_tensor.shape
        Wi = _abi.lib.mb200_env_internal_width(We, mode)
        desc = _abi.hier_describe(Wi, He)
        env4 = torch.empty(He, Wi, 4, device=self.device)
        hier = torch.empty(desc.total_floats, device=self.device)
        scratch = torch.empty(_abi.lib.mb200_env_scratch_bytes(Wi, He) // 8 + 1, dtype=torch.float64, device=self.device)
        _abi.check(_abi.lib.mb200_env_prepare(_abi.ptr(env_tensor), He, We, mode, _abi.ptr(env4), _abi.ptr(hier),
                                              C.byref(desc), _abi.ptr(scratch), _abi.stream_ptr()), "mb200_env_prepare")
        out = (env4, hier, desc, He, We, mode)
        if cache:
            self._env = out
        return out

    # ---------------------------------------------------------------- cfg
    def make_cfg(self, spp, seed, res_x, extra_flags=0, row0=None, rows=None):
        cam = self.camera
        c = _abi.Cfg()
        c.H, c.W, c.spp, c.max_depth = self.H, self.W, int(spp), self.max_depth
        c.seed = int(seed) & 0xFFFFFFFF
        c.filter, c.flags, c.use_mesh_normal = self.filter, self.flags | extra_flags, int(self.use_mesh_normal)
        c.row0 = self.row0 if row0 is None else row0
        c.rows = self.rows if rows is None else rows
        c.view[:] = cam.view_matrix.reshape(-1).tolist()
        c.proj[:] = cam.proj_matrix.reshape(-1).tolist()
        c.cam_to_world[:] = cam.to_world.astype(np.float32).reshape(-1).tolist()
        c.tan_half_fov_x = cam.tan_half_fov_x
        c.env_u_shift = (np.float32(0.5) / np.float32(res_x - 1)) if (c.flags & _abi.FLAG_ENV_HALF_TEXEL) else 0.0
        return c


class SceneParameters(dict):
    """`mi.traverse(scene)` stand-in with the keys the reference touches (inverse_img_w_mi.py:61-64, :72-78,
    :216-220, :334-342; render_final.py:182-192).  Assign, then call update()."""
    KEYS = ("shape.bsdf.a", "shape.bsdf.r", "shape.bsdf.m", "shape.bsdf.n", "shape.bsdf.use_mesh_normal",
            "emitter.data", "integrator.max_depth")
    TRANS_KEYS = ("shape.bsdf.bg", "shape.bsdf.mask", "shape.bsdf.specTrans", "shape.bsdf.ior")   # TransBSDF.traverse, mi_plugin.py:1763-1770

    def __init__(self, scene):
        super().__init__()
        self._scene = scene
        self._dirty = set()
        dict.__setitem__(self, "shape.bsdf.a", scene.a)
        dict.__setitem__(self, "shape.bsdf.r", scene.r)
        dict.__setitem__(self, "shape.bsdf.m", scene.m)
        dict.__setitem__(self, "shape.bsdf.n", scene.n)
        dict.__setitem__(self, "shape.bsdf.use_mesh_normal", scene.use_mesh_normal)
        dict.__setitem__(self, "emitter.data", scene.env_user)
        dict.__setitem__(self, "integrator.max_depth", scene.max_depth)
        if scene.trans is not None:
            dict.__setitem__(self, "shape.bsdf.bg", scene.trans.bg)
            dict.__setitem__(self, "shape.bsdf.mask", scene.trans.mask)
            dict.__setitem__(self, "shape.bsdf.specTrans", scene.trans.spec_trans)
            dict.__setitem__(self, "shape.bsdf.ior", scene.trans.ior)

    def __setitem__(self, key, value):
        if key in self.TRANS_KEYS and self._scene.trans is None:
            raise KeyError(f"{key!r} exists only when the shape's BSDF is TransBSDF (Scene.set_bsdf)")
        if key not in self.KEYS + self.TRANS_KEYS:
            raise KeyError(f"unknown scene parameter {key!r}; known: {self.KEYS + self.TRANS_KEYS}")
        dict.__setitem__(self, key, value)
        self._dirty.add(key)

    def update(self):
        s = self._scene
        H, W = s.H, s.W
        shapes = {"shape.bsdf.a": (H, W, 3), "shape.bsdf.r": (H, W, 1), "shape.bsdf.m": (H, W, 1), "shape.bsdf.n": (H, W, 3)}
        for key in sorted(self._dirty):
            v = self[key]
            if key in shapes:
                if not isinstance(v, torch.Tensor):
                    raise TypeError(f"{key} must be a torch.Tensor")
                if tuple(v.shape) != shapes[key]:
                    if key in ("shape.bsdf.r", "shape.bsdf.m") and tuple(v.shape) == (H, W):
                        v = v.unsqueeze(-1)
                    else:
                        raise ValueError(f"{key} must have shape {shapes[key]}, got {tuple(v.shape)}")
                if v.dtype != torch.float32:
                    raise TypeError(f"{key} must be float32")
                if not v.is_cuda:
                    raise ValueError(f"{key} must be a CUDA tensor (no CPU fallback)")
                setattr(s, key.rsplit(".", 1)[1], v.detach().contiguous())
            elif key == "shape.bsdf.use_mesh_normal":
                s.use_mesh_normal = bool(v)
            elif key == "emitter.data":
                s.set_envmap(v, _abi.ENV_ASSIGNED)
            elif key == "integrator.max_depth":
                s.max_depth = int(v)
            elif key == "shape.bsdf.bg":
                if not isinstance(v, torch.Tensor) or tuple(v.shape) != (H, W, 3) or v.dtype != torch.float32 or not v.is_cuda:
                    raise ValueError(f"shape.bsdf.bg must be a float32 CUDA tensor of shape {(H, W, 3)}")
                s.trans.bg = v.detach().contiguous()
            elif key == "shape.bsdf.mask":
                if not isinstance(v, torch.Tensor) or tuple(v.shape) != (H, W) or not v.is_cuda:
                    raise ValueError(f"shape.bsdf.mask must be a CUDA tensor of shape {(H, W)}")
                s.trans.mask = (v != 0).to(torch.uint8).contiguous()
            elif key == "shape.bsdf.specTrans":
                s.trans.spec_trans = float(v)
            elif key == "shape.bsdf.ior":
                s.trans.ior = float(v)
        self._dirty.clear()


def traverse(scene):
    return SceneParameters(scene)
