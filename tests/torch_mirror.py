"""Independent float64 torch/numpy re-implementation of the restated render operator, vectorised over lanes,
with torch AUTOGRAD providing the gradients (detach() exactly where Mitsuba detaches).

Purpose: the C oracle's adjoint is hand-derived; this mirror derives it mechanically.
`mirror == float64 oracle` (tight) validates the hand-derived adjoints and gives a second implementation of
the BSDF / Hierarchical2D / envmap / film restatement.  Test infrastructure only; small images only.
"""
import math

import numpy as np
import torch

F64 = torch.float64


# ------------------------------------------------------------------ Hierarchical2D in numpy float64
def _lvl_index(x, y, width):
    return ((x & 1) | (((x & ~1) | (y & 1)) << 1)) + (y & ~1) * width


class Hier:
    def __init__(self, data):
        data = np.asarray(data, dtype=np.float32)            # stored values are fp32, as in the oracle
        ry, rx = data.shape
        self.rx, self.ry = rx, ry
        npx, npy = rx - 1, ry - 1
        avg = (np.float32(.25) * (((data[:-1, :-1] + data[:-1, 1:]) + data[1:, :-1]) + data[1:, 1:])).astype(np.float32)
        total = avg.astype(np.float64).sum(axis=1).sum()     # rows first, then row sums
        scale = np.float32(np.float32(float(npx) * float(npy)) / np.float32(total))
        self.level0 = (data * scale).astype(np.float32)
        lv = (avg * scale).astype(np.float32)
        self.levels = []                                      # un-swizzled 2D arrays, even-padded
        max_level = int(math.ceil(math.log2(max(npx, npy)))) if max(npx, npy) > 1 else 0
        for _ in range(max_level + 1):
            h, w = lv.shape
            pad = np.zeros((h + (h & 1), w + (w & 1)), np.float32)
            pad[:h, :w] = lv
            self.levels.append(pad)
            lv = (((pad[0::2, 0::2] + pad[0::2, 1::2]) + pad[1::2, 0::2]) + pad[1::2, 1::2]).astype(np.float32)

    def sample(self, s):
        s = np.asarray(s, dtype=np.float64)
        sx, sy = s[:, 0].copy(), s[:, 1].copy()
        ox = np.zeros(len(sx), np.int64); oy = np.zeros(len(sx), np.int64)
        for l in range(len(self.levels) - 2, -1, -1):       # top-most level (the total) is skipped
            L = self.levels[l].astype(np.float64)
            ox <<= 1; oy <<= 1
            v00, v10, v01, v11 = L[oy, ox], L[oy, ox + 1], L[oy + 1, ox], L[oy + 1, ox + 1]
            sx = np.clip(sx, 0, 1); sy = np.clip(sy, 0, 1)
            r0, r1 = v00 + v10, v01 + v11
            sy = sy * (r0 + r1)
            m = sy > r0
            oy = oy + m; sy = np.where(m, sy - r0, sy) / np.where(m, r1, r0)
            c0, c1 = np.where(m, v01, v00), np.where(m, v11, v10)
            sx = sx * (c0 + c1)
            m = sx > c0
            sx = np.where(m, sx - c0, sx) / np.where(m, c1, c0); ox = ox + m
        L0 = self.level0.astype(np.float64)
        v00, v10, v01, v11 = L0[oy, ox], L0[oy, ox + 1], L0[oy + 1, ox], L0[oy + 1, ox + 1]
        r0, r1 = v00 + v10, v01 + v11
        m = np.abs(r0 - r1) > 1e-4 * (r0 + r1)
        with np.errstate(all="ignore"):
            sy = np.where(m, (r0 - np.sqrt(np.maximum(r0 * r0 + sy * (r1 * r1 - r0 * r0), 0))) / (r0 - r1), sy)
            c0 = (1 - sy) * v00 + sy * v01; c1 = (1 - sy) * v10 + sy * v11
            m = np.abs(c0 - c1) > 1e-4 * (c0 + c1)
            sx = np.where(m, (c0 - np.sqrt(np.maximum(c0 * c0 + sx * (c1 * c1 - c0 * c0), 0))) / (c0 - c1), sx)
        pdf = (1 - sx) * c0 + sx * c1
        u = (ox + sx) * (1.0 / (self.rx - 1)); v = (oy + sy) * (1.0 / (self.ry - 1))
        return u, v, pdf, ox, oy

    def eval(self, u, v):
        npx, npy = self.rx - 1, self.ry - 1
        px, py = u * npx, v * npy
        ox = np.minimum(px.astype(np.int64), npx - 1); oy = np.minimum(py.astype(np.int64), npy - 1)
        w1x, w1y = px - ox, py - oy
        L0 = self.level0.astype(np.float64)
        v00, v10, v01, v11 = L0[oy, ox], L0[oy, ox + 1], L0[oy + 1, ox], L0[oy + 1, ox + 1]
        return (1 - w1y) * ((1 - w1x) * v00 + w1x * v10) + w1y * ((1 - w1x) * v01 + w1x * v11)


# ------------------------------------------------------------------ torch pieces
def t(x):
    return torch.as_tensor(np.asarray(x), dtype=F64)


def normalize(v):
    return v / torch.sqrt((v * v).sum(-1, keepdim=True))


def frame(n):
    sign = torch.where(n[..., 2] >= 0, torch.ones_like(n[..., 2]), -torch.ones_like(n[..., 2]))
    sign = torch.copysign(torch.ones_like(n[..., 2]), n[..., 2])
    a = -1.0 / (sign + n[..., 2]); b = n[..., 0] * n[..., 1] * a
    s = torch.stack([sign * n[..., 0] * n[..., 0] * a + 1, sign * b, -sign * n[..., 0]], -1)
    tt = torch.stack([b, n[..., 1] * n[..., 1] * a + sign, -n[..., 1]], -1)
    return s, tt, n


def to_world(fr, v):
    s, tt, n = fr
    return s * v[..., 0:1] + tt * v[..., 1:2] + n * v[..., 2:3]


def pow5(x):
    x2 = x * x
    return x * (x2 * x2)


def eval_brdf(wi, wo, n, a, r, m):
    """mi_plugin.py:1372-1427, disney branch. a (L,3), r/m (L,1)."""
    h = normalize(wi + wo)
    NoL = (n * wi).sum(-1, keepdim=True).clamp_min(0); NoV = (n * wo).sum(-1, keepdim=True).clamp_min(0)
    VoH = (wo * h).sum(-1, keepdim=True).clamp_min(0); NoH = (n * h).sum(-1, keepdim=True).clamp_min(0)
    alpha = r * r; alpha2 = alpha * alpha
    denom = (NoH * NoH * (alpha2 - 1.0) + 1.0) + 1e-6
    D = alpha2 / (math.pi * denom * denom)
    pdf = 0.5 * (D / (4 * VoH.clamp_min(1e-6)) * NoH) + 0.5 * (NoL / math.pi)
    base_d = a * (1 - m)
    FD90 = 0.5 + 2 * VoH ** 2 * r
    Fo = 1 + (FD90 - 1) * pow5(1 - NoV); Fi = 1 + (FD90 - 1) * pow5(1 - NoL)
    diff = base_d / math.pi * Fo * Fi * NoL
    k = (r + 1); k = k * k / 8
    G = (1 / (NoL * (1 - k) + k + 1e-6)) * (1 / (NoV * (1 - k) + k + 1e-6))
    C0 = (1 - m) * 0.04 + m * a
    Fm = C0 + (1 - C0) * pow5(1 - VoH)
    metal = D * G * Fm / 4 * NoL
    return diff + metal, pdf[..., 0]


def nan0(v):
    return torch.where(torch.isnan(v), torch.zeros_like(v), v)


def sample_brdf(s1, s2, wo, n, a, r, m):
    """mi_plugin.py:1296-1341 (+ samplers :217-281)."""
    fr = frame(n)
    theta = torch.asin(torch.sqrt(s2[:, 0].clamp_min(0))); phi = 2 * math.pi * s2[:, 1]
    wd = nan0(to_world(fr, torch.stack([torch.sin(theta) * torch.cos(phi), torch.sin(theta) * torch.sin(phi), torch.cos(theta)], -1)))
    alpha = (r * r)[:, 0]
    ct = torch.sqrt(((1 - s2[:, 0]) / (s2[:, 0] * (alpha * alpha - 1) + 1)).clamp_min(0))
    st = torch.sqrt((1 - ct * ct).clamp_min(0))
    wh = to_world(fr, torch.stack([st * torch.cos(phi), st * torch.sin(phi), ct], -1))
    ws = normalize(nan0(2 * (wo * wh).sum(-1, keepdim=True) * wh - wo))
    diffuse = (s1 > 0.5)[:, None]
    wi = torch.where(diffuse, wd, ws)
    f, p = eval_brdf(wi, wo, n, a, r, m)
    w = torch.where((p > 1e-6)[:, None], f / (p[:, None] + 1e-6), torch.zeros_like(f))
    return wi, torch.where(p > 0, p, torch.zeros_like(p)), w


def mis(a, b):
    a = a * a; b = b * b
    w = a / (a + b)
    return torch.where(torch.isfinite(w), w, torch.zeros_like(w)).detach()


def env_ingest(env, mode_file):
    if mode_file:
        return torch.cat([env, env[:, :1]], 1)
    avg = 0.5 * (env[:, :1] + env[:, -1:])
    return torch.cat([avg, env[:, 1:-1], avg], 1)


def env_lookup(env_int, u, v, u_shift):
    """eval_spectrum(uv): returns interpolated rgb (differentiable w.r.t. env_int); u, v numpy/tensor detached."""
    He, Wi, _ = env_int.shape
    u = u - u_shift
    u = u - torch.floor(u); v = v - torch.floor(v)
    u = u * (Wi - 1); v = v * (He - 1)
    px = torch.clamp(u.to(torch.int64), max=Wi - 2); py = torch.clamp(v.to(torch.int64), max=He - 2)
    w1x = (u - px)[:, None]; w1y = (v - py)[:, None]
    e = env_int
    v0 = (1 - w1x) * e[py, px] + w1x * e[py, px + 1]
    v1 = (1 - w1x) * e[py + 1, px] + w1x * e[py + 1, px + 1]
    return (1 - w1y) * v0 + w1y * v1


def dir_to_uv(d):
    u = torch.atan2(d[:, 0], -d[:, 2]) / (2 * math.pi)
    v = torch.acos(d[:, 1].clamp(-1, 1)) / math.pi
    return u, v


def inv_sin_theta(d):
    eps = 5.9604644775390625e-08
    return 1.0 / torch.sqrt((d[:, 0] ** 2 + d[:, 2] ** 2).clamp_min(eps * eps))


def gauss(x):
    return torch.clamp(torch.exp(-2.0 * x * x) - math.exp(-8.0), min=0.0)


def render(oracle, cam, gpos, gnrm, a, r, m, n_map, env, env_mode_file, spp, seed, *, use_mesh_normal=True, quirk=True,
           half_texel=True, ad_weights=False, gaussian=True, row_stride_h=True):
    """Full-image render. a (H,W,3), r,m (H,W,1), n_map (H,W,3)|None, env (He,We,3): float64 torch (may require grad).
    gpos/gnrm numpy (H,W,4). Returns image (H,W,3) float64."""
    H, W = gpos.shape[:2]
    S = H * W * spp
    draws = t(oracle.sampler_floats_n(seed, 0, S, 8))
    pix = torch.arange(S) // spp
    py, px = pix // W, pix % W
    valid = t(gpos[..., 3].reshape(-1))[pix] != 0
    p = t(gpos[..., :3].reshape(-1, 3))[pix]; ng = t(gnrm[..., :3].reshape(-1, 3))[pix]
    # texel index through the plugin's projection (mi_plugin.py:645-671, :1378-1381)
    V, P = t(cam.view_matrix), t(cam.proj_matrix)
    ph = torch.cat([p, torch.ones(S, 1, dtype=F64)], -1)
    clip = (ph @ V.T) @ P.T
    sx = (clip[:, 0] / clip[:, 3] + 1) * 0.5 * W; sy = (clip[:, 1] / clip[:, 3] + 1) * 0.5 * H
    flat = (torch.floor(sx).long() + torch.floor(sy).long() * (H if row_stride_h else W)).clamp(0, H * W - 1)
    A = a.reshape(-1, 3)[flat]; Rr = r.reshape(-1, 1)[flat]; M = m.reshape(-1, 1)[flat]
    n_sh = ng if (use_mesh_normal or n_map is None) else n_map.reshape(-1, 3)[flat]
    view = normalize(t(cam.to_world[:3, 3]) - p)

    env_int = env_ingest(env, env_mode_file)
    He, Wi, _ = env_int.shape
    u_shift = float(np.float32(0.5) / np.float32(Wi - 1)) if half_texel else 0.0
    ed = env_int.detach().numpy().astype(np.float32)
    theta_scale = np.float32(np.float32(1.0) / np.float32(He - 1)) * np.float32(math.pi)
    sin_t = np.sin((np.arange(He, dtype=np.float32) * theta_scale).astype(np.float64)).astype(np.float32)
    lum = ((ed[..., 0] * np.float32(0.212671) + ed[..., 1] * np.float32(0.715160)) + ed[..., 2] * np.float32(0.072169)).astype(np.float32)
    hier = Hier((lum * sin_t[:, None]).astype(np.float32))

    # ---- emitter sampling
    hu, hv, hpdf, _, _ = hier.sample(draws[:, 2:4].numpy())
    u_em = t(hu) + u_shift; v_em = t(hv)
    th, ph_ = v_em * math.pi, u_em * 2 * math.pi
    d_em = torch.stack([torch.sin(th) * torch.sin(ph_), torch.cos(th), -torch.sin(th) * torch.cos(ph_)], -1)
    pdf_em = t(hpdf) * inv_sin_theta(d_em) / (2 * math.pi ** 2)
    le_em = env_lookup(env_int, u_em, v_em, u_shift)
    f_em, p_em = eval_brdf(d_em, view, n_sh, A, Rr, M)
    act_em = (pdf_em != 0)
    L = torch.where(act_em[:, None], f_em * le_em / pdf_em[:, None] * mis(pdf_em, p_em)[:, None], torch.zeros_like(f_em))
    # ---- BSDF sampling
    with torch.no_grad():
        wi, pdf_bs, w_primal = sample_brdf(draws[:, 4], draws[:, 5:7], view, n_sh, A, Rr, M)
        d_bs = to_world(frame(ng), wi) if quirk else wi
    w_bs = w_primal
    if ad_weights:
        f2, p2 = eval_brdf(d_bs, view, n_sh, A, Rr, M)
        w_bs = torch.where((p2 > 0)[:, None], f2 / p2.detach()[:, None].clamp_min(1e-300), w_primal)
    u_b, v_b = dir_to_uv(d_bs)
    ub = u_b - u_shift; ub = ub - torch.floor(ub); vb = v_b - torch.floor(v_b)
    em_pdf = t(hier.eval(ub.numpy(), vb.numpy())) * inv_sin_theta(d_bs) / (2 * math.pi ** 2)
    le_bs = env_lookup(env_int, u_b, v_b, u_shift)
    act_bs = (w_bs.detach().max(-1).values != 0) & (pdf_bs > 0)
    L = L + torch.where(act_bs[:, None], w_bs * le_bs * mis(pdf_bs, em_pdf)[:, None], torch.zeros_like(L))
    # ---- primary misses see the envmap directly
    jx, jy = draws[:, 0], draws[:, 1]
    dmiss = t(cam.pixel_ray_dirs((px + jx).numpy(), (py + jy).numpy()))
    um, vm = dir_to_uv(dmiss)
    L = torch.where(valid[:, None], L, env_lookup(env_int, um, vm, u_shift))
    # ---- film
    acc = torch.zeros(H * W, 4, dtype=F64)
    L4 = torch.cat([L, torch.ones(S, 1, dtype=F64)], -1)
    if gaussian:
        for oj in range(-2, 3):
            for oi in range(-2, 3):
                relx, rely = (oi + 0.5) - jx, (oj + 0.5) - jy
                w = torch.where(relx.abs() <= 2, gauss(relx), torch.zeros_like(relx)) * torch.where(rely.abs() <= 2, gauss(rely), torch.zeros_like(rely))
                qx, qy = px + oi, py + oj
                ok = (qx >= 0) & (qx < W) & (qy >= 0) & (qy < H)
                acc = acc.index_add(0, (qy * W + qx)[ok], (L4 * w[:, None])[ok])
    else:
        acc = acc.index_add(0, pix, L4)
    wsum = torch.where(acc[:, 3:] == 0, torch.ones_like(acc[:, 3:]), acc[:, 3:])
    return (acc[:, :3] / wsum).reshape(H, W, 3)
