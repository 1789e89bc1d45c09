#!/usr/bin/env python
"""Golden vectors for the Dr.Jit-typed BSDF plugins of the reference: MatDiffBSDF (mi_plugin.py:1229-1475) and
TransBSDF (:1477-1770), produced by EXECUTING THE REFERENCE'S OWN SOURCE from /root/reference on top of the numpy
float32 stand-ins of drjit / mitsuba in drjit_np_shim.py (mitsuba / drjit themselves are not installable here).

What runs from the reference, unmodified: mi_world_to_screen, mi_diffuse_sampler, mi_specular_sampler, D_GGX, G_Smith,
MatDiffBSDF.{__init__ (camera matrices), sample, sample_brdf, eval_brdf, eval_pdf},
TransBSDF.{__init__, calculate_refraction, calculate_refracted_screen_coor, sample, sample_brdf, eval_brdf, eval_pdf}.

Run:  python tests/golden/make_bsdf_golden.py   -> matdiff_bsdf.npz, trans_bsdf.npz next to this file.
"""
import math
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
sys.path.insert(0, HERE)
import drjit_np_shim as shim  # noqa: E402


def load_reference_plugin():
    import torch  # noqa: F401  (before the stubs go in)
    dr, mi = shim.install()

    class _Any:
        def __init__(self, *a, **k): pass
        def __call__(self, *a, **k): return _Any()
        def __getattr__(self, k): return _Any()

    for name in ("open3d", "lovely_tensors", "matplotlib", "matplotlib.pyplot"):
        m = types.ModuleType(name)
        def _ga(n):
            if n.startswith("__"):
                raise AttributeError(n)
            return _Any()
        m.__getattr__ = _ga
        m.monkey_patch = lambda *a, **k: None
        sys.modules[name] = m
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.path.insert(0, REF)
    import myutils.mi_plugin as mp
    return mp, dr, mi


def lanes(L, H, W, seed, cam_json):
    """Seeded lanes: surface points on pixel-centre rays (so texel indices are far from rounding boundaries), normals
    facing the camera, a view direction and a light direction in the upper hemisphere (some below, to exercise clamps)."""
    import json
    import torch
    rng = np.random.RandomState(seed)
    meta = json.load(open(cam_json))
    to_world = np.array(meta["to_world"][0], np.float64)
    fov = math.radians(meta["x_fov"][0])
    px = rng.randint(0, W, L); py = rng.randint(0, H, L)
    jx = rng.uniform(0.2, 0.8, L); jy = rng.uniform(0.2, 0.8, L)
    t = math.tan(0.5 * fov)
    # camera looks down +z of its local frame (Mitsuba); NDC x,y in [-1,1] as mi_world_to_screen inverts it
    ndc_x = (px + jx) / W * 2 - 1; ndc_y = (py + jy) / H * 2 - 1
    depth = rng.uniform(16, 33, L)
    view = np.linalg.inv(to_world)
    # invert clip = P @ (view @ p):  x_ndc = (f/aspect * xc) / (-zc), y_ndc = f*yc / (-zc)
    f = 1 / t; aspect = W / H
    zc = -depth
    xc = ndc_x * (-zc) / (f / aspect); yc = ndc_y * (-zc) / f
    pc = np.stack([xc, yc, zc, np.ones(L)], -1)
    p = (to_world @ pc.T).T[:, :3]
    cam_pos = to_world[:3, 3]
    wi = cam_pos[None] - p; wi /= np.linalg.norm(wi, axis=-1, keepdims=True)
    n = wi + 0.6 * rng.randn(L, 3); n /= np.linalg.norm(n, axis=-1, keepdims=True)
    wo = n + 0.9 * rng.randn(L, 3); wo /= np.linalg.norm(wo, axis=-1, keepdims=True)
    g = torch.Generator(device="cpu").manual_seed(seed)
    # 64 x 64 random blocks tiled over the image: the fixture stores the blocks only (tests rebuild the maps with np.tile)
    B = 64
    blocks = dict(a=torch.rand(B, B, 3, generator=g).numpy(), r=(torch.rand(B, B, 1, generator=g) * 0.93 + 0.07).numpy(),
                  m=torch.rand(B, B, 1, generator=g).numpy(), bg=torch.rand(B, B, 3, generator=g).numpy())
    maps = {k: np.ascontiguousarray(np.tile(v, (H // B, W // B, 1))) for k, v in blocks.items()}
    maps["blocks"] = blocks
    yy, xx = np.mgrid[0:H, 0:W]
    maps["mask"] = (((xx // 16) + (yy // 16)) % 2 == 0)
    return dict(p=p.astype(np.float32), n=n.astype(np.float32), wi=wi.astype(np.float32), wo=wo.astype(np.float32),
                s1=rng.rand(L).astype(np.float32), s2=rng.rand(L, 2).astype(np.float32)), maps


def main():
    mp, dr, mi = load_reference_plugin()
    cam_json = os.path.join(REF, "myutils", "default_cam.json")
    H = W = 512
    L = 8192
    ln, maps = lanes(L, H, W, 1234, cam_json)
    si = shim.FakeSI(ln["p"], ln["n"], ln["wi"])
    wo_local = si.to_local(mi.Vector3f(ln["wo"]))

    # ---------------------------------------------------------------- MatDiffBSDF
    b = mp.MatDiffBSDF(mi.Properties(cam_meta=cam_json))
    b.a = mi.TensorXf(maps["a"]); b.r = mi.TensorXf(maps["r"]); b.m = mi.TensorXf(maps["m"])
    sc = mp.mi_world_to_screen(si.p, b.view_matrix, b.persp_proj_matx, b.width, b.height)
    f, pdf = b.eval_pdf(None, si, wo_local)
    bs, w = b.sample(None, si, mi.Float(ln["s1"]), mi.Vector2f(ln["s2"]), True)
    blk = {k + "_block": v for k, v in maps["blocks"].items()}
    np.savez_compressed(os.path.join(HERE, "matdiff_bsdf.npz"), H=H, W=W, **ln, a_block=blk["a_block"], r_block=blk["r_block"], m_block=blk["m_block"],
                        view=b.view_matrix.m, proj=b.persp_proj_matx.m, screen=sc.numpy(),
                        wo_world_used=si.to_world(wo_local).numpy(), wi_world_used=si.to_world(si.wi).numpy(),
                        eval_f=f.numpy(), eval_pdf=pdf.numpy(), sample_wo=bs.wo.numpy(), sample_pdf=bs.pdf.numpy(), sample_weight=w.numpy())

    # ---------------------------------------------------------------- TransBSDF (both refract_distance settings)
    out = dict(H=H, W=W, **ln, **blk, mask_cell=16, view=b.view_matrix.m, proj=b.persp_proj_matx.m)   # mask = 16-px checkerboard
    for tag, props, ior, st in (("k", dict(ior=1.2, keep_albedo_color=True), 1.2, 0.4), ("d", dict(), 1.3, 0.8)):
        t = mp.TransBSDF(mi.Properties(cam_meta=cam_json, **props))
        t.a = mi.TensorXf(maps["a"]); t.r = mi.TensorXf(maps["r"]); t.m = mi.TensorXf(maps["m"])
        t.bg = mi.TensorXf(maps["bg"]); t.mask = mi.TensorXf(maps["mask"].astype(np.float32)) >= 1
        t.specTrans = st if tag == "k" else t.specTrans; t.ior = ior
        wi_w = si.to_world(si.wi)
        rsc = t.calculate_refracted_screen_coor(wi_w, si.n, 1.0 / t.ior, si.p, sc)
        f, pdf = t.eval_pdf(None, si, wo_local)
        bs, w = t.sample(None, si, mi.Float(ln["s1"]), mi.Vector2f(ln["s2"]), True)
        out.update({f"{tag}_ior": np.float32(t.ior), f"{tag}_specTrans": np.float32(shim._raw(t.specTrans)),
                    f"{tag}_refract_distance": np.float32(t.refract_distance), f"{tag}_refr_screen": rsc.numpy(),
                    f"{tag}_eval_f": f.numpy(), f"{tag}_eval_pdf": pdf.numpy(), f"{tag}_sample_wo": bs.wo.numpy(),
                    f"{tag}_sample_pdf": bs.pdf.numpy(), f"{tag}_sample_weight": w.numpy(), f"{tag}_eta": np.float32(bs.eta)})
    np.savez_compressed(os.path.join(HERE, "trans_bsdf.npz"), **out)
    print("wrote matdiff_bsdf.npz, trans_bsdf.npz")


if __name__ == "__main__":
    main()
