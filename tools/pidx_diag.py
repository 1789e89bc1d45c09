"""Diagnostic: primary-visibility index path vs BVH path, per pixel (box filter, 1 spp => one path per pixel)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import materialist_b200 as mb
from test_gpu_mesh_parity import _scene, _cuda_scene, REF_FLAGS

for (H, W) in ((48, 48), (96, 64)):
    cam, verts, tris, a, r, m, env = _scene(H, W)
    for spp, gauss, md in ((1, False, 2), (1, False, 4), (16, True, 4)):
        out = {}
        for on in ("1", "0"):
            os.environ["MB200_PRIMARY_INDEX"] = on
            s = _cuda_scene(cam, verts, tris, env, REF_FLAGS, max_depth=md, gaussian=gauss)
            ta, tr, tm = (torch.from_numpy(np.ascontiguousarray(x)).cuda() for x in (a, r, m))
            out[on] = mb.render(s, spp=spp, seed=21, albedo=ta, roughness=tr, metallic=tm).detach().cpu().numpy()
        d = np.abs(out["1"] - out["0"]).max(-1)
        ys, xs = np.nonzero(d)
        print(f"{H}x{W} spp {spp} gauss {gauss} depth {md}: {len(ys)} differing pixels, max abs {d.max():.3e}")
        for y, x in list(zip(ys, xs))[:12]:
            print("   ", y, x, out["1"][y, x], out["0"][y, x])
