#!/usr/bin/env python
"""Generates the golden fixtures under tests/golden/ by IMPORTING THE REFERENCE (read-only, /root/reference) in
this container and running its own torch / numpy functions on seeded inputs.  The reference cannot travel to the
GPU box, so the vectors are committed; this script is what made them.

mitsuba / drjit / open3d / lovely_tensors / matplotlib are not installed here, so they are replaced by inert stub
modules: only the pure torch / numpy code paths of the reference are exercised:
  * myutils/mi_plugin.py   D_GGX, G_Smith, fresnelSchlick, diffuse_sampler, specular_sampler, eval_brdf (pdf)
  * mymodels/mlps.py       PosMLP (brdf_net 'arm' and envmap_net 'envmap' instantiations of inverse_img_w_mi.py:117,163)
  * myutils/envmap_utils.py build_envmap, sample_envmap (+ its searchsorted indices), lookup_envmap
  * myutils/computeSH.py   computeK, computeSHFromImage, reconstImageFromSH

Run:  python tests/golden/make_golden.py      (writes *.npz next to this file)
"""
import math
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def install_stubs():
    class _Any:
        def __init__(self, *a, **k): pass
        def __call__(self, *a, **k): return _Any()
        def __getattr__(self, k): return _Any()
        def __or__(self, o): return self
        def __pos__(self): return 0

    mi = types.ModuleType("mitsuba")
    mi.set_variant = lambda *a, **k: None
    mi.BSDF = type("BSDF", (), {"__init__": lambda self, props=None: None})
    mi.register_bsdf = lambda *a, **k: None
    mi.__getattr__ = lambda name: _Any()
    dr = types.ModuleType("drjit")
    dr.__getattr__ = lambda name: _Any()
    dr.wrap_ad = lambda **k: (lambda f: f)
    for name, mod in (("mitsuba", mi), ("drjit", dr)):
        sys.modules[name] = mod
    for name in ("open3d", "lovely_tensors", "matplotlib", "matplotlib.pyplot", "cv2_stub"):
        m = types.ModuleType(name)
        m.__getattr__ = lambda n: _Any()
        m.monkey_patch = lambda *a, **k: None
        sys.modules[name] = m
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if not hasattr(np, "math"):
        np.math = math                      # computeSH.py:63 uses np.math.factorial (numpy < 2)
    sys.path.insert(0, REF)


def g(seed):
    return torch.Generator(device="cpu").manual_seed(seed)


def bsdf_terms(mp):
    n = 4096
    cos_h, NoV, NoL, VoH = (torch.rand(n, generator=g(10 + i)) for i in range(4))
    rough = torch.rand(n, generator=g(20)) * 0.93 + 0.07
    F0 = torch.rand(n, generator=g(21))
    D = mp.D_GGX(cos_h, rough); G = mp.G_Smith(NoV, NoL, rough); F = mp.fresnelSchlick(VoH, F0)
    # samplers: only the polar angle is convention-free (the reference's torch frame differs from mi.Frame3f)
    s2 = torch.rand(n, 2, generator=g(22))
    normal = torch.nn.functional.normalize(torch.randn(n, 3, generator=g(23)), dim=-1)
    wo = torch.nn.functional.normalize(normal + 0.8 * torch.randn(n, 3, generator=g(24)), dim=-1)
    wi_d = mp.diffuse_sampler(s2, normal)
    wi_s = mp.specular_sampler(s2, rough[:, None], wo, normal)
    h_s = torch.nn.functional.normalize(wi_s + wo, dim=-1)
    # pdf of the torch eval_brdf == pdf of MatDiffBSDF.eval_brdf (mi_plugin.py:365-369 vs :1397-1401)
    mat = {"albedo": torch.rand(n, 3, generator=g(25)), "roughness": rough[:, None], "metallic": torch.rand(n, 1, generator=g(26)),
           "normal": normal}
    wi = torch.nn.functional.normalize(normal + 0.9 * torch.randn(n, 3, generator=g(27)), dim=-1)
    _, pdf = mp.eval_brdf(wi, wo, normal, mat, True)
    np.savez_compressed(os.path.join(HERE, "bsdf_terms.npz"), cos_h=cos_h.numpy(), NoV=NoV.numpy(), NoL=NoL.numpy(), VoH=VoH.numpy(),
                        rough=rough.numpy(), F0=F0.numpy(), D=D.numpy(), G=G.numpy(), F=F.numpy(),
                        s2=s2.numpy(), normal=normal.numpy(), wo=wo.numpy(),
                        diffuse_cos=(wi_d * normal).sum(-1).numpy(), specular_cos_h=(h_s * normal).sum(-1).numpy(),
                        albedo=mat["albedo"].numpy(), metallic=mat["metallic"].numpy(), wi=wi.numpy(), pdf=pdf.reshape(-1).numpy())


def posmlp(mlps):
    out = {}
    for tag, kw, n_in, N in (("arm", dict(in_dims=7, out_dims=5, color_ch=5, output_type="arm"), 5, 32 * 32 * 4),
                             ("envmap", dict(in_dims=5, out_dims=3, color_ch=3, output_type="envmap"), 3, 512)):
        torch.manual_seed(100)
        net = mlps.PosMLP(dims=[256] * 4, skip_connection=[1, 3], weight_norm=False, multires_view=2, **kw)
        with torch.no_grad():                                   # the last layer is zero-initialised: perturb it so gradients are informative
            net.lin4.weight.normal_(0, 0.05, generator=g(101)); net.lin4.bias.normal_(0, 0.05, generator=g(102))
        x = torch.rand(N, n_in, generator=g(103)) if tag == "arm" else torch.ones(N, n_in)
        x.requires_grad_(True)
        y = net(x)
        gy = torch.randn(y.shape, generator=g(104))
        y.backward(gy)
        out[tag + "_x"] = x.detach().numpy(); out[tag + "_y"] = y.detach().numpy(); out[tag + "_gy"] = gy.numpy()
        out[tag + "_gx"] = x.grad.numpy()
        for l in range(5):
            lin = getattr(net, f"lin{l}")
            lin = lin.linear if hasattr(lin, "linear") else lin
            out[f"{tag}_W{l}"] = lin.weight.detach().numpy(); out[f"{tag}_b{l}"] = lin.bias.detach().numpy()
            out[f"{tag}_gW{l}"] = lin.weight.grad.numpy(); out[f"{tag}_gb{l}"] = lin.bias.grad.numpy()
        out[tag + "_nparams"] = np.array(sum(p.numel() for p in net.parameters()))
    np.savez_compressed(os.path.join(HERE, "posmlp.npz"), **out)


def envmap_utils(eu):
    import cv2
    os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
    out = {}
    hdr = cv2.imread(os.path.join(REF, "envmaps", "0.hdr"), cv2.IMREAD_UNCHANGED)[..., ::-1].astype(np.float32).copy()   # BGR -> RGB, (16,32,3)
    cases = {"rand16x32": (0.2 + torch.exp(torch.randn(16, 32, 3, generator=g(200)))), "hdr0": torch.from_numpy(hdr)}
    for tag, env in cases.items():
        d = eu.build_envmap(env)
        n = 2000
        s2 = torch.rand(2, n, generator=g(201))
        dirs, pdf = eu.sample_envmap(d, s2)
        v_idx = torch.searchsorted(d["m_cdf"], s2[0].reshape(-1, 1))
        u_idx = torch.searchsorted(d["c_cdf"][v_idx.flatten(), :], s2[1].reshape(-1, 1))
        w = torch.nn.functional.normalize(torch.randn(500, 3, generator=g(202)), dim=-1)
        look = eu.lookup_envmap(env, w)
        out.update({f"{tag}_env": env.numpy(), f"{tag}_c_cdf": d["c_cdf"].numpy(), f"{tag}_m_cdf": d["m_cdf"].numpy(), f"{tag}_s2": s2.numpy(),
                    f"{tag}_dirs": dirs.numpy(), f"{tag}_pdf": pdf.numpy(), f"{tag}_v_idx": v_idx.numpy(), f"{tag}_u_idx": u_idx.numpy(),
                    f"{tag}_w": w.numpy(), f"{tag}_lookup": look.numpy()})
    np.savez_compressed(os.path.join(HERE, "envmap_utils.npz"), **out)


def compute_sh(sh):
    rs = np.random.RandomState(300)
    im = rs.rand(8, 16, 3)
    np.random.seed(301)                                          # computeSHFromImage jitters with the global numpy RNG
    coef = sh.computeSHFromImage(im)
    np.random.seed(301)
    jit = np.random.random(8 * 16 * 2).reshape(8 * 16, 2)        # the (y, x) jitters it drew, in order
    larr = np.array([0, 1, 1, 1, 2, 2, 2, 2, 2, 3, 3, 3, 3, 3, 3, 3, 4, 4, 4, 4, 4, 4, 4, 4, 4], dtype=np.int32)
    marr = np.array([0, -1, 0, 1, -2, -1, 0, 1, 2, -3, -2, -1, 0, 1, 2, 3, -4, -3, -2, -1, 0, 1, 2, 3, 4], dtype=np.int32)
    K = sh.computeK(larr.copy(), marr.copy())
    rec = sh.reconstImageFromSH(coef, 16, 32, isClip=False)
    rec_clip = sh.reconstImageFromSH(coef, 16, 32, isClip=True)
    np.savez_compressed(os.path.join(HERE, "compute_sh.npz"), im=im, jitter=jit, coef=coef, K=K, rec=rec, rec_clip=rec_clip)


def helpers_and_rotate(sh, eu):
    """S4 (computeSH 'AfterRotate' variants) and the small CDF helpers of envmap_utils, from the reference's own code."""
    rs = np.random.RandomState(400)
    out = {}
    env = rs.rand(12, 24, 3).astype(np.float32)
    cams = [((0.3, -0.5, 0.8), (0.1, 1.0, 0.05), False), ((1.0, 0.2, -0.4), (0.0, 0.3, 1.0), True)]
    for k, (loc, up, inv) in enumerate(cams):
        np.random.seed(500 + k)
        out[f"rot{k}_coef"] = sh.computeSHFromImageAfterRotate(env, loc, up, isInv=inv)
        np.random.seed(500 + k)
        out[f"rot{k}_jitter"] = np.random.random(12 * 24 * 2).reshape(12 * 24, 2)
        out[f"rot{k}_rec"] = sh.reconstImageFromSHAfterRotate(out[f"rot{k}_coef"], loc, up, nrows=10, ncols=20, isClip=False, isInv=inv)
        out[f"rot{k}_cam"] = np.array([loc, up], np.float64); out[f"rot{k}_inv"] = np.array(inv)
    out["env"] = env
    u, v = rs.rand(50), rs.rand(50)
    out["uv"] = np.stack([u, v]); out["uv_color"] = np.stack([sh.uvToEnvmap(env, a, b) for a, b in zip(u, v)])
    # CDF helpers on a (16, 32) map: torch CPU
    e = torch.rand(16, 32, 3, generator=g(401))
    d = eu.build_envmap(e)
    x = torch.rand(200, generator=g(402))
    vi = eu.cdf_search_1d(d["m_cdf"], x)
    out["cdf_env"] = e.numpy(); out["cdf_x"] = x.numpy(); out["cdf_vi"] = vi.numpy()
    out["pdf1d"] = eu.get_pdf_from_cdf_1d(d["m_cdf"], vi.clamp_max(15)).numpy()
    out["interp1d"] = eu.interp_1d(d["m_cdf"], x, vi.clamp_max(15)).numpy()
    ui = eu.cdf_search_2d(d["c_cdf"], x, 5)
    out["cdf_ui"] = ui.numpy()
    out["pdf2d"] = eu.get_pdf_from_cdf_2d(d["c_cdf"], ui.clone(), 5).numpy()
    out["interp2d"] = eu.interp_2d(d["c_cdf"], x, ui.clone(), 5).numpy()          # NaN where ui == 0 (reference behaviour)
    np.savez_compressed(os.path.join(HERE, "helpers_rotate.npz"), **out)


def e5_and_s5(eu, sh):
    """E5: sample_env1 / sample_brdf1 (envmap_utils.py:7-28) with the random numbers they draw recorded (torch.rand is patched to
    replay them on the CUDA side).  S5: the torch SH variants are broken as shipped; the fixture is their INTENDED
    maths in float64 — a Riemann sum 4 pi / (W H) * sum L Y sin(theta) on the (phi, theta) = linspace grids of :410-430, with the
    reference's own working numpy building blocks for Y: computeK (:58-68) and the associated Legendre functions P_l_m (:13-56)."""
    out = {}
    n = 1500
    env = 0.2 + torch.exp(torch.randn(16, 32, 3, generator=g(600)))
    d = eu.build_envmap(env)
    normals = torch.nn.functional.normalize(torch.randn(n, 3, generator=g(601)), dim=-1)
    wo = torch.nn.functional.normalize(normals + 0.7 * torch.randn(n, 3, generator=g(602)), dim=-1)
    mat = {"albedo": torch.rand(n, 3, generator=g(603)), "roughness": torch.rand(n, 1, generator=g(604)) * 0.93 + 0.07,
           "metallic": torch.rand(n, 1, generator=g(605)), "normal": normals}
    draws = []
    real_rand = torch.rand

    def rec_rand(*shape, **kw):
        kw.pop("device", None)
        t = real_rand(*shape, generator=g(700 + len(draws)))
        draws.append(t)
        return t
    torch.rand = rec_rand
    try:
        wi, pdf, w = eu.sample_env1(wo, normals, mat, True, "cpu", d)
        wi2, pdf2, w2 = eu.sample_brdf1(wo, normals, mat, True, "cpu")
    finally:
        torch.rand = real_rand
    out.update(e5_env=env.numpy(), e5_normals=normals.numpy(), e5_wo=wo.numpy(), e5_albedo=mat["albedo"].numpy(), e5_rough=mat["roughness"].numpy(),
               e5_metal=mat["metallic"].numpy(), e5_draw0=draws[0].numpy(), e5_draw1=draws[1].numpy(), e5_draw2=draws[2].numpy(),
               e5_env_wi=wi.numpy(), e5_env_pdf=pdf.numpy(), e5_env_w=w.numpy(), e5_brdf_wi=wi2.numpy(), e5_brdf_pdf=pdf2.numpy(), e5_brdf_w=w2.numpy())
    # ---- S5
    rs = np.random.RandomState(800)
    H, W, l_max = 12, 24, 2
    img = rs.rand(H, W, 3)
    phis, thetas = np.meshgrid(np.linspace(0, 2 * np.pi, W), np.linspace(0, np.pi, H), indexing="xy")
    P = {(0, 0): sh.P_0_0, (1, 0): sh.P_1_0, (1, 1): sh.P_1_1, (2, 0): sh.P_2_0, (2, 1): sh.P_2_1, (2, 2): sh.P_2_2}
    coeffs = np.zeros((l_max + 1, 2 * l_max + 1, 3))
    Y = {}
    for l in range(l_max + 1):
        for m in range(-l, l + 1):
            K = float(sh.computeK(np.array([l]), np.array([m]))[0])
            plm = P[(l, abs(m))](thetas.astype(np.float64))
            y = K * plm * (np.sqrt(2) * np.cos(m * phis) if m > 0 else (np.sqrt(2) * np.sin(-m * phis) if m < 0 else 1.0))
            Y[(l, m)] = y
            coeffs[l, m + l] = (img * (y * np.sin(thetas))[..., None]).sum((0, 1))
    coeffs *= 4 * np.pi / (W * H)
    rec = np.zeros((H, W, 3))
    for (l, m), y in Y.items():
        rec += coeffs[l, m + l][None, None, :] * y[..., None]
    out.update(s5_img=img, s5_coeffs=coeffs, s5_rec=rec)
    np.savez_compressed(os.path.join(HERE, "e5_s5.npz"), **out)


def main():
    install_stubs()
    import myutils.mi_plugin as mp
    import mymodels.mlps as mlps
    import myutils.envmap_utils as eu
    import myutils.computeSH as sh
    bsdf_terms(mp); print("bsdf_terms.npz")
    posmlp(mlps); print("posmlp.npz")
    envmap_utils(eu); print("envmap_utils.npz")
    compute_sh(sh); print("compute_sh.npz")
    helpers_and_rotate(sh, eu); print("helpers_rotate.npz")
    e5_and_s5(eu, sh); print("e5_s5.npz")


if __name__ == "__main__":
    main()
