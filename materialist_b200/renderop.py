"""The render operator: `mi.render(scene, params, spp, seed)` and the two differentiable wrappers
`render_w_brdf` / `render_envmap` of inverse_img_w_mi.py:59-80, re-implemented as one
torch.autograd.Function over the fused sm_100a kernels (C-ABI, include/materialist_b200.h).

Forward renders with `seed`; backward is an independent adjoint render with
`seed_grad = sample_tea_32(seed, 1)[0]` and the re-evaluated-BSDF weights — it is
NOT the transpose of the forward sample set, exactly as in Mitsuba's `render_backward`.
"""
import ctypes as C

import numpy as np
import torch

from . import _abi
from .scene import Scene, traverse

_MASK32 = 0xFFFFFFFF

# Measurement hook (bench.py): when set to a list, every shade kernel launch is bracketed by CUDA events
# recorded on the launching stream and (name, start, end) is appended.
KERNEL_EVENTS = None


class _ktime:
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if KERNEL_EVENTS is not None:
            self.e0 = torch.cuda.Event(enable_timing=True); self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *exc):
        if KERNEL_EVENTS is not None:
            self.e1.record()
            KERNEL_EVENTS.append((self.name, self.e0, self.e1))
        return False


def tea32(v0, v1, rounds=4):
    """mitsuba.sample_tea_32 (host-side scalar; integer-exact)."""
    v0 &= _MASK32; v1 &= _MASK32; s = 0
    for _ in range(rounds):
        s = (s + 0x9E3779B9) & _MASK32
        v0 = (v0 + ((((v1 << 4) & _MASK32) + 0xA341316C) ^ ((v1 + s) & _MASK32) ^ ((v1 >> 5) + 0xC8013EA4))) & _MASK32
        v1 = (v1 + ((((v0 << 4) & _MASK32) + 0xAD90777D) ^ ((v0 + s) & _MASK32) ^ ((v0 >> 5) + 0x7E95761E))) & _MASK32
    return v0, v1


def default_seed_grad(seed):
    """mi.render: seed_grad = mi.sample_tea_32(seed, 1)[0] when left at 0."""
    return tea32(seed, 1)[0]


def _check_map(name, t, shape):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (no CPU fallback)")
    if tuple(t.shape) != tuple(shape):
        raise ValueError(f"{name} must have shape {tuple(shape)}, got {tuple(t.shape)}")
    return t.detach().contiguous()


def _primary_index(scene, cfg):
    """Primary-visibility index of (scene.mesh, scene camera): built on first use and kept ON THE MESH OBJECT (a Mesh is immutable once
    built, the camera is part of the key), so a new mesh — even one whose buffer lands on the old one's address — never sees a stale
    index.  MB200_PRIMARY_INDEX=0 disables it (every primary ray then walks the BVH)."""
    import os
    if os.environ.get("MB200_PRIMARY_INDEX", "1") == "0":
        return None
    mesh = scene.mesh
    key = (bytes(cfg.cam_to_world), float(cfg.tan_half_fov_x), int(cfg.H), int(cfg.W))
    cache = mesh.__dict__.setdefault("_primary_idx", {})
    idx = cache.get(key)
    if idx is None:
        nbytes = _abi.lib.mb200_mesh_primary_index_bytes(C.byref(cfg), C.byref(mesh.desc))
        if nbytes == 0:
            return None
        idx = torch.empty(nbytes // 8 + 1, dtype=torch.float64, device=scene.device)
        _abi.check(_abi.lib.mb200_mesh_primary_index_build(C.byref(cfg), C.byref(mesh.desc), _abi.ptr(mesh.buf), _abi.ptr(idx),
                                                           _abi.stream_ptr()), "mb200_mesh_primary_index_build")
        if len(cache) >= 4:                      # a handful of cameras per mesh at most (each index is ~22 ints per pixel)
            cache.pop(next(iter(cache)))
        cache[key] = idx
    return _abi.ptr(idx)


def _forward(scene, spp, seed, a, r, m, n, env_pack, extra_flags=0):
    env4, hier, desc, He, We, mode = env_pack
    cfg = scene.make_cfg(spp, seed, desc.res_x, extra_flags)
    first = C.c_int(0)
    prows = _abi.lib.mb200_fwd_partial_rows(C.byref(cfg), C.byref(first))
    stride = _abi.lib.mb200_partial_stride(cfg.filter)
    partials = torch.empty(prows, scene.W, stride, device=scene.device)
    st = _abi.stream_ptr()
    nmap = None if scene.use_mesh_normal else n
    if scene.mesh is not None and scene.mesh_forward == "wavefront":
        td = scene.trans.desc() if scene.trans is not None else None
        nbytes = _abi.lib.mb200_mesh_fwd_wf_scratch_bytes(C.byref(cfg))
        if nbytes == 0:
            raise ValueError("wavefront forward: unsupported configuration (spp too large)")
        if scene._wf_scratch is None or scene._wf_scratch.numel() * 8 < nbytes:
            scene._wf_scratch = torch.empty(nbytes // 8 + 1, dtype=torch.float64, device=scene.device)
        with _ktime("mesh_fwd_wf"):
            _abi.check(_abi.lib.mb200_mesh_shade_fwd_wf(C.byref(cfg), C.byref(td) if td is not None else None, C.byref(scene.mesh.desc),
                                                        _abi.ptr(scene.mesh.buf), _abi.fptr(a), _abi.fptr(r), _abi.fptr(m), _abi.fptr(nmap),
                                                        _abi.fptr(env4), _abi.fptr(hier), C.byref(desc), _abi.fptr(partials),
                                                        _abi.ptr(scene._wf_scratch), scene._wf_scratch.numel() * 8, _primary_index(scene, cfg), st),
                       "mb200_mesh_shade_fwd_wf")
    elif scene.trans is not None:               # TransBSDF plugin (trans_edit.py): forward only
        td = scene.trans.desc()
        if scene.mesh is not None:
            with _ktime("mesh_fwd_trans"):
                _abi.check(_abi.lib.mb200_trans_mesh_shade_fwd(C.byref(cfg), C.byref(td), C.byref(scene.mesh.desc), _abi.ptr(scene.mesh.buf),
                                                               _abi.fptr(a), _abi.fptr(r), _abi.fptr(m), _abi.fptr(nmap), _abi.fptr(env4), _abi.fptr(hier),
                                                               C.byref(desc), _abi.fptr(partials), st), "mb200_trans_mesh_shade_fwd")
        else:
            with _ktime("shade_fwd_trans"):
                _abi.check(_abi.lib.mb200_trans_shade_fwd(C.byref(cfg), C.byref(td), _abi.fptr(scene.gpos), _abi.fptr(scene.gnrm), _abi.fptr(a),
                                                          _abi.fptr(r), _abi.fptr(m), _abi.fptr(nmap), _abi.fptr(env4), _abi.fptr(hier), C.byref(desc),
                                                          _abi.fptr(partials), st), "mb200_trans_shade_fwd")
    elif scene.mesh is not None:
        with _ktime("mesh_fwd"):
            _abi.check(_abi.lib.mb200_mesh_shade_fwd(C.byref(cfg), C.byref(scene.mesh.desc), _abi.ptr(scene.mesh.buf), _abi.fptr(a), _abi.fptr(r),
                                                     _abi.fptr(m), _abi.fptr(nmap), _abi.fptr(env4), _abi.fptr(hier), C.byref(desc),
                                                     _abi.fptr(partials), st), "mb200_mesh_shade_fwd")
    else:
        with _ktime("shade_fwd"):
            _abi.check(_abi.lib.mb200_shade_fwd(C.byref(cfg), _abi.fptr(scene.gpos), _abi.fptr(scene.gnrm), _abi.fptr(a), _abi.fptr(r),
                                                _abi.fptr(m), _abi.fptr(nmap), _abi.fptr(env4), _abi.fptr(hier), C.byref(desc),
                                                _abi.fptr(partials), st), "mb200_shade_fwd")
    img = torch.empty(cfg.rows, scene.W, 3, device=scene.device)
    _abi.check(_abi.lib.mb200_film_develop(C.byref(cfg), _abi.fptr(partials), _abi.fptr(img), st), "mb200_film_develop")
    return img


def _film_weights(scene, spp, seed_grad, env_res_x, out=None):
    """Film-weight taps of the seed_grad render (needs nothing but the seed: callers may run it on a side stream while the loss
    of the forward image is still being reduced).  Returns the (wrows, W, 25) buffer, or None for the box filter."""
    cfg = scene.make_cfg(spp, seed_grad, env_res_x)
    if cfg.filter != _abi.FILTER_GAUSSIAN:
        return None
    first = C.c_int(0)
    wrows = _abi.lib.mb200_bwd_wpart_rows(C.byref(cfg), C.byref(first))
    wpart = out if out is not None else torch.empty(wrows, scene.W, _abi.FILM_TAPS, device=scene.device)
    if tuple(wpart.shape) != (wrows, scene.W, _abi.FILM_TAPS):
        raise ValueError("film-weight buffer has the wrong shape")
    with _ktime("film_weights"):
        _abi.check(_abi.lib.mb200_film_weights(C.byref(cfg), _abi.fptr(wpart), _abi.stream_ptr()), "mb200_film_weights")
    return wpart


def _backward(scene, spp, seed_grad, a, r, m, n, env_pack, grad_img_halo, want_a, want_r, want_m, want_n, want_env, out=None, wpart=None):
    """grad_img_halo: (gadj_rows, W, 3) — gradient w.r.t. the image for the shard rows plus the film halo.
    out: optional pre-zeroed (g_a, g_r, g_m) full-image buffers to accumulate into (fused optimiser path).
    wpart: film weights of this seed_grad render if the caller already computed them (_film_weights)."""
    env4, hier, desc, He, We, mode = env_pack
    cfg = scene.make_cfg(spp, seed_grad, desc.res_x)
    st = _abi.stream_ptr()
    dev = scene.device
    first = C.c_int(0)
    grows = _abi.lib.mb200_bwd_gadj_rows(C.byref(cfg), C.byref(first))
    if tuple(grad_img_halo.shape) != (grows, scene.W, 3):
        raise ValueError(f"grad image (with film halo) must be {(grows, scene.W, 3)}, got {tuple(grad_img_halo.shape)}")
    grad_img_halo = grad_img_halo.contiguous().float()
    if cfg.filter == _abi.FILTER_GAUSSIAN and wpart is None:
        wpart = _film_weights(scene, spp, seed_grad, desc.res_x)
    gadj = torch.empty(grows, scene.W, 4, device=dev)
    _abi.check(_abi.lib.mb200_film_adjoint(C.byref(cfg), _abi.fptr(wpart), _abi.fptr(grad_img_halo), _abi.fptr(gadj), st), "mb200_film_adjoint")
    H, W = scene.H, scene.W
    if out is not None:
        g_a, g_r, g_m = (o if w else None for o, w in zip(out, (want_a, want_r, want_m)))
    else:
        g_a = torch.zeros(H, W, 3, device=dev) if want_a else None
        g_r = torch.zeros(H, W, 1, device=dev) if want_r else None
        g_m = torch.zeros(H, W, 1, device=dev) if want_m else None
    g_n = torch.zeros(H, W, 3, device=dev) if (want_n and not scene.use_mesh_normal) else None
    n_slabs = _abi.lib.mb200_env_grad_slabs(He, We, mode) if want_env else 1
    g_env4 = torch.zeros((n_slabs,) + tuple(env4.shape), device=dev) if want_env else None
    nmap = None if scene.use_mesh_normal else n
    if scene.mesh is not None and scene.mesh_backward == "wavefront":
        nbytes = _abi.lib.mb200_mesh_bwd_wf_scratch_bytes(C.byref(cfg))
        if nbytes == 0:
            raise ValueError("wavefront adjoint: unsupported configuration (spp too large)")
        if scene._wf_scratch is None or scene._wf_scratch.numel() * 8 < nbytes:
            scene._wf_scratch = torch.empty(nbytes // 8 + 1, dtype=torch.float64, device=scene.device)
        with _ktime("mesh_bwd_wf"):
            _abi.check(_abi.lib.mb200_mesh_shade_bwd_wf(C.byref(cfg), C.byref(scene.mesh.desc), _abi.ptr(scene.mesh.buf), _abi.fptr(a), _abi.fptr(r),
                                                        _abi.fptr(m), _abi.fptr(nmap), _abi.fptr(env4), _abi.fptr(hier), C.byref(desc),
                                                        _abi.fptr(gadj), _abi.fptr(g_a), _abi.fptr(g_r), _abi.fptr(g_m), _abi.fptr(g_n),
                                                        _abi.fptr(g_env4), n_slabs, _abi.ptr(scene._wf_scratch), scene._wf_scratch.numel() * 8,
                                                        _primary_index(scene, cfg), st),
                       "mb200_mesh_shade_bwd_wf")
    elif scene.mesh is not None:
        with _ktime("mesh_bwd"):
            _abi.check(_abi.lib.mb200_mesh_shade_bwd(C.byref(cfg), C.byref(scene.mesh.desc), _abi.ptr(scene.mesh.buf), _abi.fptr(a), _abi.fptr(r),
                                                     _abi.fptr(m), _abi.fptr(nmap), _abi.fptr(env4), _abi.fptr(hier), C.byref(desc),
                                                     _abi.fptr(gadj), _abi.fptr(g_a), _abi.fptr(g_r), _abi.fptr(g_m), _abi.fptr(g_n),
                                                     _abi.fptr(g_env4), n_slabs, st), "mb200_mesh_shade_bwd")
    else:
        with _ktime("shade_bwd"):
            _abi.check(_abi.lib.mb200_shade_bwd(C.byref(cfg), _abi.fptr(scene.gpos), _abi.fptr(scene.gnrm), _abi.fptr(a), _abi.fptr(r),
                                                _abi.fptr(m), _abi.fptr(nmap), _abi.fptr(env4), _abi.fptr(hier), C.byref(desc),
                                                _abi.fptr(gadj), _abi.fptr(g_a), _abi.fptr(g_r), _abi.fptr(g_m), _abi.fptr(g_n),
                                                _abi.fptr(g_env4), n_slabs, st), "mb200_shade_bwd")
    g_env = None
    if want_env:
        g_env = torch.empty(He, We, 3, device=dev)
        _abi.check(_abi.lib.mb200_env_grad_finish(_abi.fptr(g_env4), n_slabs, He, We, mode, _abi.fptr(g_env), st), "mb200_env_grad_finish")
    if want_n and g_n is None:
        g_n = torch.zeros(H, W, 3, device=dev)
    return g_a, g_r, g_m, g_n, g_env


class _RenderOp(torch.autograd.Function):
    """Stand-in for mitsuba's `_RenderOp` + `dr.wrap_ad`."""

    @staticmethod
    def forward(ctx, scene, spp, seed, seed_grad, halo_exchange, a, r, m, n, env):
        H, W = scene.H, scene.W
        a_ = _check_map("albedo", a, (H, W, 3)) if a is not None else scene.a
        r_ = _check_map("roughness", r, (H, W, 1)) if r is not None else scene.r
        m_ = _check_map("metallic", m, (H, W, 1)) if m is not None else scene.m
        n_ = _check_map("normal", n, (H, W, 3)) if n is not None else scene.n
        if env is not None:
            if env.ndim != 3 or env.shape[-1] != 3:
                raise ValueError("envmap must be (He, We, 3)")
            if env.dtype != torch.float32:
                raise TypeError("envmap must be float32")
            if not env.is_cuda:
                raise ValueError("envmap must be a CUDA tensor (no CPU fallback)")
            env_pack = scene.prepared_env(env, _abi.ENV_ASSIGNED)
            # params['emitter.data'] = envmap; params.update() persists in the scene (inverse_img_w_mi.py:63-64)
            scene.env_user, scene.env_mode, scene._env = env.detach(), _abi.ENV_ASSIGNED, env_pack
        else:
            env_pack = scene.prepared_env()
        # params['shape.bsdf.*'] = ...; params.update() persists in the scene (inverse_img_w_mi.py:72-78)
        scene.a, scene.r, scene.m, scene.n = a_, r_, m_, n_
        ctx.scene, ctx.spp, ctx.seed_grad, ctx.halo_exchange = scene, spp, seed_grad, halo_exchange
        ctx.maps = (a_, r_, m_, n_)
        ctx.env_pack = env_pack
        ctx.has = (a is not None, r is not None, m is not None, n is not None, env is not None)
        return _forward(scene, spp, seed, a_, r_, m_, n_, env_pack)

    @staticmethod
    def backward(ctx, grad_img):
        scene = ctx.scene
        if scene.trans is not None:
            raise RuntimeError("TransBSDF is a forward-only editing plugin (the reference never differentiates it); render under torch.no_grad()")
        need = ctx.needs_input_grad[5:]
        want = [h and nd for h, nd in zip(ctx.has, need)]
        grad_img = grad_img.contiguous()
        if ctx.halo_exchange is not None:
            grad_img = ctx.halo_exchange(grad_img)          # adds the neighbours' 2-row film halo (multi-GPU)
        a_, r_, m_, n_ = ctx.maps
        g = _backward(scene, ctx.spp, ctx.seed_grad, a_, r_, m_, n_, ctx.env_pack, grad_img, *want)
        g = [gi if w else None for gi, w in zip(g, want)]
        return (None, None, None, None, None, *g)


def render(scene, params=None, spp=64, seed=0, seed_grad=0, albedo=None, roughness=None, metallic=None, normal=None,
           envmap=None, halo_exchange=None):
    """`mi.render(scene, params, spp=, seed=, seed_grad=)` → (rows, W, 3) float32 CUDA tensor.

    Tensors passed as albedo / roughness / metallic / normal / envmap are the differentiable leaves
    (what `dr.wrap_ad` attaches in the reference); everything else is read from the scene state set through
    `traverse(scene)[...] = ...; params.update()`.
    """
    if not isinstance(scene, Scene):
        raise TypeError("scene must be a materialist_b200.Scene")
    spp = int(spp)
    if spp <= 0:
        raise ValueError("spp must be positive")
    if scene.H * scene.W * spp >= 2 ** 32:
        raise ValueError("H*W*spp must be < 2^32 (one wavefront, as in the reference)")
    if not seed_grad:
        seed_grad = default_seed_grad(int(seed))
    return _RenderOp.apply(scene, spp, int(seed), int(seed_grad), halo_exchange, albedo, roughness, metallic, normal, envmap)


def render_envmap(scene, envmap, spp=64, seed=None):
    """inverse_img_w_mi.py:59-67 — sets emitter.data and renders; gradients flow to `envmap`.
    The reference draws `seed = np.random.randint(0, 1000)` inside; pass `seed` for reproducibility."""
    if seed is None:
        seed = int(np.random.randint(0, 1000))
    return render(scene, spp=spp, seed=seed, envmap=envmap)


def render_w_brdf(scene, albedo, roughness, metallic, normal=None, spp=64, seed=None):
    """inverse_img_w_mi.py:69-80 — sets shape.bsdf.{a,r,m[,n]} and renders; gradients flow to the maps."""
    if seed is None:
        seed = int(np.random.randint(0, 1000))
    return render(scene, spp=spp, seed=seed, albedo=albedo, roughness=roughness, metallic=metallic, normal=normal)


def sample_indices(scene, spp, seed):
    """Integer decisions of every lane of the shard: (S, 4) int32 = (hier off.x, off.y, texel index, lobe)."""
    env4, hier, desc, He, We, mode = scene.prepared_env()
    cfg = scene.make_cfg(spp, seed, desc.res_x)
    out = torch.empty(cfg.rows * scene.W * spp, 4, dtype=torch.int32, device=scene.device)
    _abi.check(_abi.lib.mb200_debug_sample_indices(C.byref(cfg), _abi.fptr(scene.gpos), _abi.ptr(scene.r), _abi.fptr(hier),
                                                   C.byref(desc), _abi.ptr(out), _abi.stream_ptr()), "mb200_debug_sample_indices")
    return out


def sample_record(scene, spp, seed, ad_weights=False, want_radiance=False):
    """Decision record of every lane of the shard through the kernels' own forward sample function: (S, 12) int32 =
    (hier off.x, off.y, texel index, lobe, emitter-sample envmap cell, BSDF-direction envmap cell, bits of the emitter
    direction x/y/z, bits of the BSDF-sampled direction x/y/z) [+ (S, 3) radiance per lane]."""
    env4, hier, desc, He, We, mode = scene.prepared_env()
    cfg = scene.make_cfg(spp, seed, desc.res_x, _abi.FLAG_AD_WEIGHTS if ad_weights else 0)
    S = cfg.rows * scene.W * spp
    out = torch.empty(S, 12, dtype=torch.int32, device=scene.device)
    rad = torch.empty(S, 3, device=scene.device) if want_radiance else None
    nmap = None if scene.use_mesh_normal else scene.n
    _abi.check(_abi.lib.mb200_debug_sample_record(C.byref(cfg), _abi.fptr(scene.gpos), _abi.fptr(scene.gnrm), _abi.ptr(scene.a), _abi.ptr(scene.r),
                                                  _abi.ptr(scene.m), _abi.fptr(nmap), _abi.fptr(env4), _abi.fptr(hier), C.byref(desc),
                                                  _abi.ptr(out), _abi.ptr(rad), _abi.stream_ptr()), "mb200_debug_sample_record")
    return (out, rad) if want_radiance else out
