import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import Case, rel_l2
from oracle import oracle as orc
import materialist_b200 as mb

O = orc.Oracle()
for (H, W, spp, He, We) in ((2160, 3840, 256, 1024, 2048), (2160, 3840, 256, 128, 256), (512, 512, 256, 1024, 2048)):
    c = Case(H=H, W=W, spp=spp, He=He, We=We)
    row0, rows = H // 2 - 2, 4
    s = c.scene()
    a, r, m, n = c.torch_maps()
    s.set_shard(row0, rows)
    img = mb.render(s, spp=c.spp, seed=3, albedo=a, roughness=r, metallic=m).cpu().numpy()
    ref, idx_ref = c.oracle_fwd(O, 3, want_indices=True, row0=row0, rows=rows)
    s.r = r
    idx = mb.sample_indices(s, c.spp, 3).cpu().numpy()
    bad = (idx != idx_ref).any(axis=1)
    d = np.abs(img - ref)
    rel = d / np.maximum(np.abs(ref), 1e-3)
    print(dict(H=H, W=W, He=He, rel_l2=rel_l2(img, ref), lanes_differ=int(bad.sum()), cols=[int((idx != idx_ref)[:, k].sum()) for k in range(4)],
               max_abs=float(d.max()), n_rel_gt_1e3=int((rel > 1e-3).sum()), median_rel=float(np.median(rel))), flush=True)
    env_int, hier, dd = O.env_prepare(c.env, c.env_mode)
    env4, hier_g, desc, *_ = s.prepared_env()
    hg = hier_g.cpu().numpy()[:hier.size]
    print("  hier differs:", int((hg.view(np.uint32) != hier.view(np.uint32)).sum()), "of", hier.size, flush=True)
