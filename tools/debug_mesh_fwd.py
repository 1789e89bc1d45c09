"""Diagnostic: where does the CUDA-vs-oracle forward difference of the mesh mode live (a few flipped samples or everywhere)?"""
import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as orc
from test_gpu_mesh_parity import _scene, _cuda_scene
from test_reference_render_pin import REF_FLAGS, pin_cfg, rel_l2
from materialist_b200 import renderop

O = orc.Oracle()
for (H, spp, max_depth, flags, gaussian, fn) in ((40, 32, 4, REF_FLAGS, False, False), (40, 32, 4, REF_FLAGS, True, False), (40, 32, 3, REF_FLAGS | 8, False, True),
                                                 (64, 64, 4, REF_FLAGS, False, False)):
    W = H
    cam, verts, tris, a, r, m, env = _scene(H, W)
    om = O.mesh_create(verts, tris, face_normals=fn)
    env_int, hier, d = O.env_prepare(env, orc.ENV_ASSIGNED)
    cfg = pin_cfg(d, 5, 0, H, spp=spp, H=H, W=W, max_depth=max_depth, flags=flags)
    cfg.filter = orc.FILTER_GAUSSIAN if gaussian else orc.FILTER_BOX
    ref = O.mesh_render_fwd(cfg, om, a, r, m, None, env_int, hier, d)
    s = _cuda_scene(cam, verts, tris, env, flags & ~8, max_depth, gaussian, fn)
    ta, tr, tm = (torch.from_numpy(x).cuda() for x in (a, r, m))
    img = renderop._forward(s, spp, 5, ta, tr, tm, None, s.prepared_env(), extra_flags=flags & 8).cpu().numpy()
    err = np.abs(img - ref).sum(-1).reshape(-1)
    order = np.argsort(-err)
    tot = np.linalg.norm(ref)
    print(f"H={H} spp={spp} depth={max_depth} flags={flags} gauss={gaussian} face={fn}: rel-L2 {rel_l2(img, ref):.3e}; pixels with |d|>1e-4*mean: {(err > 1e-4 * ref.mean()).sum()} of {H*W}")
    for k in order[:6]:
        print("   px", k % W, k // W, "err", err[k], "ref", ref.reshape(-1, 3)[k], "x spp =", err[k] * spp)
    keep = np.ones(H * W, bool); keep[order[:max(1, H * W // 200)]] = False
    print("   rel-L2 without the worst 0.5% pixels:", np.linalg.norm((img - ref).reshape(-1, 3)[keep]) / np.linalg.norm(ref.reshape(-1, 3)[keep]))
