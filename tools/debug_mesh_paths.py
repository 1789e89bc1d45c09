"""Diagnostic (GPU): per-PATH radiance of the wavefront mesh forward (read from its scratch buffer: float4 L[path]) against the
oracle's per-path records -> which samples differ, and what the oracle's path looked like there."""
import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as orc
from test_gpu_mesh_parity import _scene, _cuda_scene
from test_reference_render_pin import REF_FLAGS, pin_cfg, rel_l2
from materialist_b200 import renderop

O = orc.Oracle()
cases = ((40, 32, 4, REF_FLAGS, False, False), (40, 32, 4, REF_FLAGS, True, False))
for (H, spp, max_depth, flags, gaussian, fn) in cases:
    W = H
    cam, verts, tris, a, r, m, env = _scene(H, W)
    om = O.mesh_create(verts, tris, face_normals=fn)
    env_int, hier, d = O.env_prepare(env, orc.ENV_ASSIGNED)
    cfg = pin_cfg(d, 5, 0, H, spp=spp, H=H, W=W, max_depth=max_depth, flags=flags)
    cfg.filter = orc.FILTER_GAUSSIAN if gaussian else orc.FILTER_BOX
    rec = O.mesh_path_records(cfg, om, a, r, m, None, env_int, hier, d)
    s = _cuda_scene(cam, verts, tris, env, flags & ~8, max_depth, gaussian, fn)
    ta, tr, tm = (torch.from_numpy(x).cuda() for x in (a, r, m))
    img = renderop._forward(s, spp, 5, ta, tr, tm, None, s.prepared_env(), extra_flags=flags & 8)
    torch.cuda.synchronize()
    prows = img.shape[0] + (4 if gaussian else 0)
    prows = min(prows, H)
    nb = prows * W * spp
    raw = s._wf_scratch.view(torch.uint8)[256 + 6 * nb * 16: 256 + 7 * nb * 16].view(torch.float32).view(nb, 4).cpu().numpy()
    L = raw[:, :3]
    assert nb == rec.shape[0], (nb, rec.shape)
    err = np.abs(L - rec[:, :3]).sum(-1)
    scale = np.abs(rec[:, :3]).sum(-1) + 1e-6
    bad = np.where(err > 1e-4 * scale)[0]
    print(f"H={H} gauss={gaussian}: paths {nb}, differing (>1e-4 rel) {len(bad)}; max rel {np.max(err / scale):.3e}")
    for i in bad[:12]:
        pix, sidx = divmod(int(i), spp)
        print(f"   path {i} px ({pix % W},{pix // W}) s {sidx}: cuda {L[i]} oracle {rec[i, :3]} nv {rec[i, 3]:.0f} miss_k {rec[i, 4]:.0f} "
              f"v0 (tri {rec[i, 5]:.0f} vis {rec[i, 6]:.0f} lobe {rec[i, 7]:.0f}) v1 (tri {rec[i, 8]:.0f} vis {rec[i, 9]:.0f} lobe {rec[i, 10]:.0f}) "
              f"v2 (tri {rec[i, 11]:.0f} vis {rec[i, 12]:.0f} lobe {rec[i, 13]:.0f})")
