"""ctypes front-end of the CPU ORACLE (oracle/mb_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module; nothing under materialist_b200/ does.  Arrays are numpy, fp32, C-contiguous.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
MAX_LEVELS = 24

FLAG_WO_WORLD_QUIRK = 1
FLAG_ROW_STRIDE_H = 2
FLAG_ENV_HALF_TEXEL = 4
FLAG_AD_WEIGHTS = 8
FILTER_BOX, FILTER_GAUSSIAN = 0, 1
ENV_ASSIGNED, ENV_FILE = 0, 1


class Cfg(C.Structure):
    _fields_ = [("H", C.c_int32), ("W", C.c_int32), ("spp", C.c_int32), ("max_depth", C.c_int32),
                ("seed", C.c_uint32), ("filter", C.c_int32), ("flags", C.c_int32), ("use_mesh_normal", C.c_int32),
                ("row0", C.c_int32), ("rows", C.c_int32),
                ("view", C.c_float * 16), ("proj", C.c_float * 16), ("cam_to_world", C.c_float * 16),
                ("tan_half_fov_x", C.c_float), ("env_u_shift", C.c_float)]


class HierDesc(C.Structure):
    _fields_ = [("res_x", C.c_int32), ("res_y", C.c_int32), ("n_levels", C.c_int32),
                ("lvl_off", C.c_int32 * MAX_LEVELS), ("lvl_w", C.c_int32 * MAX_LEVELS), ("lvl_h", C.c_int32 * MAX_LEVELS),
                ("total_floats", C.c_int32)]


class Trans(C.Structure):
    """mb200_trans (include/materialist_b200.h): TransBSDF parameters; bg / mask are HOST pointers for the oracle."""
    _fields_ = [("ior", C.c_float), ("spec_trans", C.c_float), ("refract_distance", C.c_float), ("reserved", C.c_int32),
                ("bg", C.c_void_p), ("mask", C.c_void_p)]


def build(force=False):
    """Compile the oracle with gcc (Makefile in this directory)."""
    so = os.path.join(_HERE, "libmb_oracle.so")
    src_m = max(os.path.getmtime(f) for f in [os.path.join(_HERE, f) for f in ("mb_oracle.c", "mb_oracle_mesh.c", "mb_oracle_aux.c", "Makefile")]
                + [os.path.join(_HERE, "..", "include", "mb200_exact_math.h")])
    if force or not os.path.exists(so) or os.path.getmtime(so) < src_m:
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return so


def _f32(x):
    return np.ascontiguousarray(x, dtype=np.float32)


def _p(x, t=C.c_float):
    return None if x is None else x.ctypes.data_as(C.POINTER(t))


class Oracle:
    def __init__(self, double=False):
        build()
        name = "libmb_oracle_f64.so" if double else "libmb_oracle.so"
        self.lib = C.CDLL(os.path.join(_HERE, name))
        self.lib.mbo_seed_grad.restype = C.c_uint32
        self.double = bool(self.lib.mbo_is_double())

    # ------------------------------------------------------------------ RNG
    def tea32(self, v0, v1, rounds=4):
        a, b = C.c_uint32(), C.c_uint32()
        self.lib.mbo_tea32(C.c_uint32(v0), C.c_uint32(v1), C.c_int(rounds), C.byref(a), C.byref(b))
        return a.value, b.value

    def pcg32_stream(self, initstate, initseq, n):
        out = np.zeros(n, np.uint32)
        self.lib.mbo_pcg32_stream(C.c_uint64(initstate), C.c_uint64(initseq), C.c_int(n), _p(out, C.c_uint32))
        return out

    def sampler_floats(self, seed, lane, n):
        out = np.zeros(n, np.float32)
        self.lib.mbo_sampler_floats(C.c_uint32(seed), C.c_uint32(lane), C.c_int(n), _p(out))
        return out

    def sampler_floats_n(self, seed, lane0, nlanes, n):
        out = np.zeros((nlanes, n), np.float32)
        self.lib.mbo_sampler_floats_n(C.c_uint32(seed), C.c_uint32(lane0), C.c_int(nlanes), C.c_int(n), _p(out))
        return out

    def seed_grad(self, seed):
        return int(self.lib.mbo_seed_grad(C.c_uint32(seed)))

    # ------------------------------------------------------------------ hierarchy / envmap
    def hier_describe(self, res_x, res_y):
        d = HierDesc()
        rc = self.lib.mbo_hier_describe(C.c_int(res_x), C.c_int(res_y), C.byref(d))
        if rc != 0:
            raise ValueError("hier_describe failed")
        return d

    def hier_build(self, data):
        data = _f32(data)
        d = self.hier_describe(data.shape[1], data.shape[0])
        hier = np.zeros(d.total_floats, np.float32)
        self.lib.mbo_hier_build(_p(data), C.byref(d), _p(hier))
        return hier, d

    def hier_sample(self, hier, d, s):
        s = _f32(s); n = s.shape[0]
        uv = np.zeros((n, 2), np.float32); pdf = np.zeros(n, np.float32); off = np.zeros((n, 2), np.int32)
        self.lib.mbo_hier_sample_n(_p(hier), C.byref(d), _p(s), C.c_int(n), _p(uv), _p(pdf), _p(off, C.c_int32))
        return uv, pdf, off

    def hier_eval(self, hier, d, uv):
        uv = _f32(uv); out = np.zeros(uv.shape[0], np.float32)
        self.lib.mbo_hier_eval_n(_p(hier), C.byref(d), _p(uv), C.c_int(uv.shape[0]), _p(out))
        return out

    def env_prepare(self, env, mode=ENV_ASSIGNED):
        env = _f32(env); He, We, _ = env.shape
        Wi = We + 1 if mode == ENV_FILE else We
        d = self.hier_describe(Wi, He)
        env_int = np.zeros((He, Wi, 3), np.float32); hier = np.zeros(d.total_floats, np.float32)
        self.lib.mbo_env_prepare(_p(env), C.c_int(He), C.c_int(We), C.c_int(mode), _p(env_int), _p(hier), C.byref(d))
        return env_int, hier, d

    def env_grad_finish(self, g_int, We, mode=ENV_ASSIGNED):
        g_int = _f32(g_int); He = g_int.shape[0]
        out = np.zeros((He, We, 3), np.float32)
        self.lib.mbo_env_grad_finish(_p(g_int), C.c_int(He), C.c_int(We), C.c_int(mode), _p(out))
        return out

    def env_eval(self, env_int, u_shift, dirs):
        dirs = _f32(dirs); out = np.zeros_like(dirs)
        self.lib.mbo_env_eval_n(_p(env_int), C.c_int(env_int.shape[0]), C.c_int(env_int.shape[1]), C.c_float(u_shift),
                                _p(dirs), C.c_int(dirs.shape[0]), _p(out))
        return out

    def env_sample(self, env_int, hier, d, u_shift, s):
        s = _f32(s); n = s.shape[0]
        dirs = np.zeros((n, 3), np.float32); pdf = np.zeros(n, np.float32); w = np.zeros((n, 3), np.float32)
        self.lib.mbo_env_sample_n(_p(env_int), _p(hier), C.byref(d), C.c_float(u_shift), _p(s), C.c_int(n), _p(dirs), _p(pdf), _p(w))
        return dirs, pdf, w

    def env_pdf(self, hier, d, u_shift, dirs):
        dirs = _f32(dirs); out = np.zeros(dirs.shape[0], np.float32)
        self.lib.mbo_env_pdf_n(_p(hier), C.byref(d), C.c_float(u_shift), _p(dirs), C.c_int(dirs.shape[0]), _p(out))
        return out

    # ------------------------------------------------------------------ BSDF lanes
    def bsdf_eval_pdf(self, cfg, p, n_geo, wi_w, wo_w, a, r, m, n_opt=None):
        p, n_geo, wi_w, wo_w = map(_f32, (p, n_geo, wi_w, wo_w)); L = p.shape[0]
        f = np.zeros((L, 3), np.float32); pdf = np.zeros(L, np.float32)
        self.lib.mbo_bsdf_eval_pdf(C.byref(cfg), C.c_int64(L), _p(p), _p(n_geo), _p(wi_w), _p(wo_w), _p(a), _p(r), _p(m), _p(n_opt), _p(f), _p(pdf))
        return f, pdf

    def bsdf_eval_grad(self, cfg, p, n_geo, wi_world, wo_world, a, r, m, w, n_opt=None):
        """Adjoint of eval_pdf's rgb value for cotangent w (L,3): (g_a (L,3), g_r (L), g_m (L), g_n (L,3)) at each lane's texel."""
        p, n_geo, wi_world, wo_world, w = map(_f32, (p, n_geo, wi_world, wo_world, w)); L = p.shape[0]
        ga = np.zeros((L, 3), np.float32); gr = np.zeros(L, np.float32); gm = np.zeros(L, np.float32); gn = np.zeros((L, 3), np.float32)
        self.lib.mbo_bsdf_eval_grad(C.byref(cfg), C.c_int64(L), _p(p), _p(n_geo), _p(wi_world), _p(wo_world), _p(_f32(a)), _p(_f32(r)), _p(_f32(m)),
                                    _p(None if n_opt is None else _f32(n_opt)), _p(w), _p(ga), _p(gr), _p(gm), _p(gn))
        return ga, gr, gm, gn

    def bsdf_sample(self, cfg, p, n_geo, wi_w, s1, s2, a, r, m, n_opt=None):
        p, n_geo, wi_w, s1, s2 = map(_f32, (p, n_geo, wi_w, s1, s2)); L = p.shape[0]
        wo = np.zeros((L, 3), np.float32); pdf = np.zeros(L, np.float32); w = np.zeros((L, 3), np.float32)
        self.lib.mbo_bsdf_sample(C.byref(cfg), C.c_int64(L), _p(p), _p(n_geo), _p(wi_w), _p(s1), _p(s2), _p(a), _p(r), _p(m), _p(n_opt), _p(wo), _p(pdf), _p(w))
        return wo, pdf, w

    # ------------------------------------------------------------------ TransBSDF mode (mi_plugin.py:1477-1770)
    def set_trans(self, ior=1.3, spec_trans=0.8, refract_distance=1.0, bg=None, mask=None):
        """Switches EVERY BSDF evaluation of this oracle instance's library to TransBSDF (bg (H,W,3) fp32, mask (H,W) bool);
        set_trans(bg=None) switches back to MatDiffBSDF."""
        if bg is None:
            self._trans_keep = None
            self.lib.mbo_set_trans(None)
            return
        bg = _f32(bg); mask = np.ascontiguousarray(mask, dtype=np.uint8)
        t = Trans(float(ior), float(spec_trans), float(refract_distance), 0, bg.ctypes.data, mask.ctypes.data)
        self._trans_keep = (bg, mask, t)
        self.lib.mbo_set_trans(C.byref(t))

    def trans_refracted_texel(self, cfg, p, n_geo, wi_w):
        p, n_geo, wi_w = map(_f32, (p, n_geo, wi_w)); L = p.shape[0]
        sc = np.zeros((L, 2), np.float32); flat = np.zeros(L, np.int64)
        self.lib.mbo_trans_refracted_texel(C.byref(cfg), C.c_int64(L), _p(p), _p(n_geo), _p(wi_w), _p(sc), _p(flat, C.c_int64))
        return sc, flat

    def terms(self, cos_h, NoV, NoL, VoH, rough, F0):
        arrs = list(map(_f32, (cos_h, NoV, NoL, VoH, rough, F0))); n = arrs[0].shape[0]
        D = np.zeros(n, np.float32); G = np.zeros(n, np.float32); F = np.zeros(n, np.float32)
        self.lib.mbo_terms(C.c_int(n), *[_p(x) for x in arrs], _p(D), _p(G), _p(F))
        return D, G, F

    def world_to_screen(self, cfg, p):
        p = _f32(p); n = p.shape[0]
        out = np.zeros((n, 2), np.float32); flat = np.zeros(n, np.int64)
        self.lib.mbo_world_to_screen(C.byref(cfg), C.c_int(n), _p(p), _p(out), _p(flat, C.c_int64))
        return out, flat

    # ------------------------------------------------------------------ render
    def render_fwd(self, cfg, gpos, gnrm, a, r, m, n_opt, env_int, hier, d, want_indices=False):
        img = np.zeros((cfg.rows, cfg.W, 3), np.float32)
        idx = np.zeros((cfg.rows * cfg.W * cfg.spp, 4), np.int32) if want_indices else None
        rc = self.lib.mbo_render_fwd(C.byref(cfg), _p(gpos), _p(gnrm), _p(a), _p(r), _p(m), _p(n_opt), _p(env_int), _p(hier), C.byref(d),
                                     _p(img), _p(idx, C.c_int32))
        if rc != 0:
            raise RuntimeError(f"mbo_render_fwd rc={rc}")
        return (img, idx) if want_indices else img

    def sample_record(self, cfg, gpos, gnrm, a, r, m, n_opt, env_int, hier, d, want_radiance=False):
        """(S, 12) int32 decision record per lane (see mbo_sample_record) [+ (S, 3) radiance]."""
        S = cfg.rows * cfg.W * cfg.spp
        out = np.zeros((S, 12), np.int32)
        rad = np.zeros((S, 3), np.float32) if want_radiance else None
        rc = self.lib.mbo_sample_record(C.byref(cfg), _p(gpos), _p(gnrm), _p(a), _p(r), _p(m), _p(n_opt), _p(env_int), _p(hier), C.byref(d),
                                        _p(out, C.c_int32), _p(rad))
        if rc != 0:
            raise RuntimeError(f"mbo_sample_record rc={rc}")
        return (out, rad) if want_radiance else out

    def render_bwd(self, cfg, gpos, gnrm, a, r, m, n_opt, env_int, hier, d, grad_img, want=("a", "r", "m", "env")):
        H, W = cfg.H, cfg.W
        grad_img = _f32(grad_img)
        assert grad_img.shape == (H, W, 3)
        g = {}
        if "a" in want: g["a"] = np.zeros((H, W, 3), np.float32)
        if "r" in want: g["r"] = np.zeros((H, W, 1), np.float32)
        if "m" in want: g["m"] = np.zeros((H, W, 1), np.float32)
        if "n" in want: g["n"] = np.zeros((H, W, 3), np.float32)
        if "env" in want: g["env_int"] = np.zeros_like(env_int)
        rc = self.lib.mbo_render_bwd(C.byref(cfg), _p(gpos), _p(gnrm), _p(a), _p(r), _p(m), _p(n_opt), _p(env_int), _p(hier), C.byref(d),
                                     _p(grad_img), _p(g.get("a")), _p(g.get("r")), _p(g.get("m")), _p(g.get("n")), _p(g.get("env_int")))
        if rc != 0:
            raise RuntimeError(f"mbo_render_bwd rc={rc}")
        return g

    # ------------------------------------------------------------------ mesh mode (oracle/mb_oracle_mesh.c)
    def mesh_create(self, verts, tris, face_normals=False):
        """verts (nv,3) float, tris (nt,3) int -> opaque handle (median-split BVH inside)."""
        verts = _f32(verts); tris = np.ascontiguousarray(tris, dtype=np.int32)
        self.lib.mbo_mesh_create.restype = C.c_void_p
        h = self.lib.mbo_mesh_create(_p(verts), C.c_int(verts.shape[0]), _p(tris, C.c_int32), C.c_int(tris.shape[0]), C.c_int(int(face_normals)))
        return C.c_void_p(h)

    def mesh_destroy(self, mesh):
        self.lib.mbo_mesh_destroy(mesh)

    def mesh_render_fwd(self, cfg, mesh, a, r, m, n_opt, env_int, hier, d, want_stats=False):
        img = np.zeros((cfg.rows, cfg.W, 3), np.float32)
        stats = np.zeros(5, np.int64)
        rc = self.lib.mbo_mesh_render_fwd(C.byref(cfg), mesh, _p(a), _p(r), _p(m), _p(n_opt), _p(env_int), _p(hier), C.byref(d),
                                          _p(img), _p(stats, C.c_int64))
        if rc != 0:
            raise RuntimeError(f"mbo_mesh_render_fwd rc={rc}")
        return (img, stats) if want_stats else img

    def mesh_path_records(self, cfg, mesh, a, r, m, n_opt, env_int, hier, d):
        """Debug: (rows*W*spp, 16) per-path records (L.rgb, nv, miss_k, 3 x (tri, visible, lobe), jx, jy)."""
        out = np.zeros((cfg.rows * cfg.W * cfg.spp, 16), np.float32)
        rc = self.lib.mbo_mesh_path_records(C.byref(cfg), mesh, _p(a), _p(r), _p(m), _p(n_opt), _p(env_int), _p(hier), C.byref(d), _p(out))
        if rc != 0:
            raise RuntimeError(f"mbo_mesh_path_records rc={rc}")
        return out

    def mesh_render_bwd(self, cfg, mesh, a, r, m, n_opt, env_int, hier, d, grad_img, want=("a", "r", "m", "env")):
        H, W = cfg.H, cfg.W
        grad_img = _f32(grad_img)
        assert grad_img.shape == (H, W, 3)
        g = {}
        if "a" in want: g["a"] = np.zeros((H, W, 3), np.float32)
        if "r" in want: g["r"] = np.zeros((H, W, 1), np.float32)
        if "m" in want: g["m"] = np.zeros((H, W, 1), np.float32)
        if "n" in want: g["n"] = np.zeros((H, W, 3), np.float32)
        if "env" in want: g["env_int"] = np.zeros_like(env_int)
        rc = self.lib.mbo_mesh_render_bwd(C.byref(cfg), mesh, _p(a), _p(r), _p(m), _p(n_opt), _p(env_int), _p(hier), C.byref(d),
                                          _p(grad_img), _p(g.get("a")), _p(g.get("r")), _p(g.get("m")), _p(g.get("n")), _p(g.get("env_int")))
        if rc != 0:
            raise RuntimeError(f"mbo_mesh_render_bwd rc={rc}")
        return g

    def mesh_primary(self, cfg, mesh, jx=0.5, jy=0.5):
        H, W = cfg.H, cfg.W
        pos = np.zeros((H, W, 3), np.float32); nrm = np.zeros((H, W, 3), np.float32); tri = np.zeros((H, W), np.int32)
        self.lib.mbo_mesh_primary(C.byref(cfg), mesh, C.c_float(jx), C.c_float(jy), _p(pos), _p(nrm), _p(tri, C.c_int32))
        return pos, nrm, tri

    def mesh_intersect(self, mesh, o, d, maxt=None, brute=False, any_hit=False):
        o, d = _f32(o), _f32(d); n = o.shape[0]
        maxt = None if maxt is None else _f32(maxt)
        tri = np.zeros(n, np.int32); tuv = np.zeros((n, 3), np.float32)
        self.lib.mbo_mesh_intersect_n(mesh, _p(o), _p(d), _p(maxt), C.c_int(n), C.c_int(int(brute)), C.c_int(int(any_hit)),
                                      _p(tri, C.c_int32), _p(tuv))
        return tri, tuv

    def mesh_vertex_normals(self, mesh, nv):
        out = np.zeros((nv, 3), np.float32)
        return out if self.lib.mbo_mesh_vertex_normals(mesh, _p(out)) else None
