"""`mitsuba` stand-in (see materialist_b200.compat).  Pure host logic lives in SceneSpec / ParamsProxy (testable without a GPU);
the operator itself is reached through `_backend_render` / `SceneSpec.build`."""
import os
import types

import numpy as np
import torch

_DIFF_KEYS = {"shape.bsdf.a": "albedo", "shape.bsdf.r": "roughness", "shape.bsdf.m": "metallic", "shape.bsdf.n": "normal",
              "emitter.data": "envmap"}
_BSDF_NAMES = {"matdiffbsdf": "matDiffBSDF", "transbsdf": "TransBSDF"}


def look_at(origin, target, up):
    """mi.ScalarTransform4f.look_at: camera-to-world with +z = viewing direction, +x = left (Mitsuba's convention)."""
    o, t, u = (np.asarray(v, np.float64) for v in (origin, target, up))
    d = t - o; d /= np.linalg.norm(d)
    left = np.cross(u, d); left /= np.linalg.norm(left)
    nu = np.cross(d, left)
    m = np.eye(4)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = left, nu, d, o
    return m


class SensorSpec:
    def __init__(self, d):
        self.fov = float(d.get("fov", 35.0))
        tw = d.get("to_world")
        self.to_world = np.eye(4) if tw is None else np.asarray(getattr(tw, "matrix", tw), np.float64).reshape(4, 4)
        film = d.get("film", {})
        self.width, self.height = int(film.get("width", 512)), int(film.get("height", 512))


class SceneSpec:
    """What mi.load_dict({'type': 'scene', ...}) of the reference describes (inverse_img_w_mi.py:42-55, render_final.py:40-97)."""

    def __init__(self, d):
        shape = d["shape"]
        if shape.get("type") != "ply":
            raise ValueError("only a 'ply' shape is supported (the reference's depth-derived mesh)")
        self.mesh_path = shape["filename"]
        b = dict(shape.get("bsdf", {}))
        name = _BSDF_NAMES.get(str(b.get("type", "MatDiffBSDF")).lower())
        if name is None:
            raise ValueError(f"unsupported BSDF plugin {b.get('type')!r}: the operator provides MatDiffBSDF and TransBSDF")
        self.bsdf = {"name": name, **{k: v for k, v in b.items() if k in ("ior", "keep_albedo_color")}}
        self.cam_meta = b.get("cam_meta")
        self.use_mesh_normal = bool(b.get("use_mesh_normal", True))
        integ = d.get("integrator", {})
        if integ.get("type", "path") != "path":
            raise ValueError("only the 'path' integrator is supported")
        self.max_depth = int(integ.get("max_depth", 4))
        s = d.get("sensor", {})
        self.sensor = s if isinstance(s, SensorSpec) else SensorSpec(s)
        em = d.get("emitter", {})
        if em.get("type") != "envmap":
            raise ValueError("only an 'envmap' emitter is supported")
        self.envmap_file, self.envmap_bitmap = em.get("filename"), em.get("bitmap")
        self._scene = None

    def build(self, device="cuda"):
        """The materialist_b200.Scene behind this description (built once)."""
        if self._scene is None:
            import materialist_b200 as mb
            from ..gbuffer import read_image
            from ..mesh import read_ply_mesh
            cam = mb.Camera(to_world=self.sensor.to_world, x_fov=self.sensor.fov, width=self.sensor.width, height=self.sensor.height)
            if self.envmap_bitmap is not None:
                env = np.asarray(self.envmap_bitmap, np.float32)[..., :3]
            else:
                env = read_image(self.envmap_file)[..., :3]
            verts, tris = read_ply_mesh(self.mesh_path)
            sc = mb.Scene.from_mesh(verts, tris, cam, device=device, envmap=torch.from_numpy(np.ascontiguousarray(env)),
                                    use_mesh_normal=self.use_mesh_normal, max_depth=self.max_depth)
            sc.set_envmap(torch.from_numpy(np.ascontiguousarray(env)), mb._abi.ENV_FILE)
            sc.set_bsdf(self.bsdf)
            self._scene = sc
        return self._scene


class ParamsProxy(dict):
    """mi.traverse(scene): assignments are kept AS GIVEN (tensors keep their autograd history: they are the leaves mi.render
    differentiates) and forwarded to the scene on update()."""

    def __init__(self, spec):
        super().__init__()
        self._spec, self._dirty = spec, set()

    def __setitem__(self, k, v):
        dict.__setitem__(self, k, v); self._dirty.add(k)

    def update(self, *a, **k):
        if a or k:
            return dict.update(self, *a, **k)
        sc = self._spec.build() if isinstance(self._spec, SceneSpec) else self._spec
        real = _backend_traverse(sc)
        for key in sorted(self._dirty):
            v = self[key]
            if isinstance(v, Bitmap):
                v = torch.from_numpy(np.asarray(v))
            if isinstance(v, torch.Tensor):
                v = v.detach()
                if v.dtype.is_floating_point:
                    v = v.float()
                if key in ("shape.bsdf.r", "shape.bsdf.m") and v.ndim == 2:
                    v = v.unsqueeze(-1)
                v = v.to(sc.device).contiguous()
            real[key] = v
        real.update()
        self._dirty.clear()


def _backend_traverse(scene):
    from ..scene import traverse
    return traverse(scene)


def _backend_render(scene, spp, seed, seed_grad, leaves):
    """The one call into the operator (tests replace it to check the binding without a GPU)."""
    import materialist_b200 as mb
    return mb.render(scene, spp=spp, seed=seed, seed_grad=seed_grad, **leaves)


def render(scene, params=None, sensor=0, seed=0, seed_grad=0, spp=0, spp_grad=0):
    """mi.render(scene, params, spp=, seed=): forward with `seed`; the backward is the adjoint render with
    seed_grad = sample_tea_32(seed, 1)[0] (python/util.py).  Tensors assigned through `params` that require grad are attached."""
    if spp_grad not in (0, spp):
        raise ValueError("spp_grad != spp is not supported")
    leaves = {}
    p = params                      # (mi.render without `params` differentiates nothing, as in Mitsuba)
    if p is not None:
        for key, arg in _DIFF_KEYS.items():
            v = p.get(key) if isinstance(p, dict) else None
            if isinstance(v, torch.Tensor) and v.requires_grad:
                if key in ("shape.bsdf.r", "shape.bsdf.m") and v.ndim == 2:
                    v = v.unsqueeze(-1)
                leaves[arg] = v
    sc = scene.build() if isinstance(scene, SceneSpec) else scene
    return _backend_render(sc, int(spp) if spp else 64, int(seed), int(seed_grad), leaves)


class Bitmap:
    """mi.Bitmap(path | array): np.array(bitmap) gives the float32 pixels (EXR / HDR / PNG through the native readers)."""

    def __init__(self, src):
        if isinstance(src, (str, os.PathLike)):
            from ..gbuffer import read_image
            self.a = np.ascontiguousarray(read_image(str(src)))
        else:
            self.a = np.ascontiguousarray(np.asarray(src.detach().cpu() if isinstance(src, torch.Tensor) else src, np.float32))

    def __array__(self, dtype=None, copy=None):
        return self.a if dtype is None else self.a.astype(dtype)

    @property
    def shape(self):
        return self.a.shape


def write_bitmap(path, img, *a, **k):
    from ..imageio import write_bitmap as wb
    if isinstance(img, torch.Tensor):
        img = img.detach().cpu().numpy()
    wb(str(path), np.asarray(img, np.float32))


def TensorXf(data=0.0, shape=None):
    if shape is not None:
        return torch.full(tuple(shape), float(data))
    if isinstance(data, Bitmap):
        return torch.from_numpy(np.asarray(data))
    return data if isinstance(data, torch.Tensor) else torch.as_tensor(np.asarray(data, np.float32))


class _Any:
    """Placeholder for every other mitsuba name the reference's plugin classes mention at import / definition time."""

    def __init__(self, *a, **k): pass
    def __call__(self, *a, **k): return _Any()
    def __getattr__(self, k): return _Any()
    def __or__(self, o): return self
    __ror__ = __or__


class BSDF:
    def __init__(self, props=None):
        self.props = props


class OptixDenoiser:
    """render_final.py:163: the OptiX AI denoiser is out of scope; the stand-in returns its input unchanged."""

    def __init__(self, *a, **k): pass
    def __call__(self, img, *a, **k): return img


def module():
    m = types.ModuleType("mitsuba")
    m.__doc__ = __doc__
    state = {"variant": None, "bsdfs": {}}
    m._state = state
    m.set_variant = lambda name, *a: state.__setitem__("variant", name)
    m.variant = lambda: state["variant"]
    m.register_bsdf = lambda name, ctor: state["bsdfs"].__setitem__(name, ctor)
    m.BSDF, m.Bitmap, m.TensorXf, m.OptixDenoiser, m.render = BSDF, Bitmap, TensorXf, OptixDenoiser, render
    m.ScalarTransform4f = types.SimpleNamespace(look_at=look_at)
    m.util = types.SimpleNamespace(write_bitmap=write_bitmap)

    def load_dict(d):
        t = d.get("type")
        if t == "perspective":
            return SensorSpec(d)
        if t == "scene":
            return SceneSpec(d)
        raise ValueError(f"mi.load_dict: unsupported object type {t!r}")
    m.load_dict = load_dict
    m.traverse = lambda scene: ParamsProxy(scene)
    m.__getattr__ = lambda name: _Any()
    return m
