"""Diagnostic (GPU, tuning build with -DMB200_TRAV_STATS): traversal steps per ray of the wavefront mesh kernels on the C2m scene.
MB200_LIB=materialist_b200/tuning/lib_travstats.so python tools/trav_stats.py"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import materialist_b200 as mb
from materialist_b200 import synthetic, renderop

H = W = 512; spp = 16
cam = mb.Camera(width=W, height=H)
verts, tris = synthetic.grid_mesh(synthetic.bumpy_positions(H, W, cam))
env = synthetic.envmap(128, 256)
for depth in (2, 4):
    s = mb.Scene.from_mesh(verts, tris, cam, device="cuda", envmap=env, use_mesh_normal=True, max_depth=depth)
    a, r, m = (t.cuda() for t in synthetic.materials(H, W, seed_base=1))
    renderop._forward(s, spp, 3, a, r, m, None, s.prepared_env())
    torch.cuda.synchronize()
    st = s._wf_scratch.view(torch.int64)[20:29].cpu().numpy().reshape(3, 3)
    nrays = H * W * spp
    print(f"max_depth {depth}: primary rays {nrays}")
    for mode, name in ((0, "closest"), (1, "shadow")):
        node, leaf, it = st[mode]
        print(f"  {name}: box steps {node} ({node / nrays:.1f} per primary ray), triangle tests {leaf} ({leaf / nrays:.1f}), warp iterations {it} "
              f"-> lanes active per iteration {(node + leaf) / max(it, 1):.1f} of 32")
