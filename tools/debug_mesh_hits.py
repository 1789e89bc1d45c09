"""Diagnostic: GPU BVH vs oracle BVH vs oracle brute force on the pixel-centre-ish rays of the shipped indoor mesh."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as orc
from materialist_b200.mesh import Mesh
from materialist_b200.scene import Camera

g = np.load(os.path.join(ROOT, "tests/golden/indoor_pin.npz"))
O = orc.Oracle()
om = O.mesh_create(g["verts"], g["tris"])
gm = Mesh(g["verts"], g["tris"])
cam = Camera(width=512, height=512)
for jx, jy in ((0.5, 0.5), (0.013, 0.977)):
    sx, sy = np.meshgrid(np.arange(512) + jx, np.arange(512) + jy)
    d = cam.pixel_ray_dirs(sx.reshape(-1), sy.reshape(-1)).astype(np.float32)
    o = np.zeros_like(d)
    t_o, tuv_o = O.mesh_intersect(om, o, d)
    t_g, tuv_g = gm.intersect(o, d)
    t_g = t_g.cpu().numpy(); tuv_g = tuv_g.cpu().numpy()
    bad = np.nonzero(t_o != t_g)[0]
    print("jitter", jx, jy, "mismatches", len(bad), "of", len(t_o), "tuv mismatches", int((tuv_o != tuv_g).any(-1).sum()))
    if len(bad):
        sel = bad[:40]
        t_b, tuv_b = O.mesh_intersect(om, o[sel], d[sel], brute=True)
        print("oracle-BVH == brute:", int((t_o[sel] == t_b).sum()), " GPU == brute:", int((t_g[sel] == t_b).sum()), "of", len(sel))
        for i, k in enumerate(sel[:10]):
            print(k, "orc", t_o[k], tuv_o[k], "gpu", t_g[k], tuv_g[k], "brute", t_b[i], tuv_b[i])
