"""GPU: the TransBSDF path (csrc: trans_* device functions fused into shade_fwd_kernel<.., TRANS> / mesh_fwd_kernel<.., TRANS>,
lane kernels in mb200_lanes.cu) through the C-ABI, against
  * golden vectors produced by executing the reference's own TransBSDF source (tests/golden/trans_bsdf.npz), and
  * the CPU oracle in TransBSDF mode (oracle.set_trans) for whole renders in G-buffer and mesh mode.
Bar: refracted background texel indices bit-exact vs the oracle; radiance <= 1e-4 rel-L2."""
import os

import numpy as np
import pytest
import torch

from helpers import Case, REF_FLAGS, rel_l2
from oracle import oracle as orc
from test_bsdf_plugin_golden import check_dirs, check_lanes, cfg512, load_golden, sampled_pdf_ok

pytestmark = pytest.mark.gpu


def _si(g, m0):
    from materialist_b200.myutils.mi_plugin import SurfaceInteraction
    p, n = (torch.from_numpy(g[k]).cuda() for k in ("p", "n"))
    si = SurfaceInteraction(p, n, torch.zeros_like(p))
    si.wi = si.to_local(torch.from_numpy(m0["wi_world_used"]).cuda())
    return si


@pytest.mark.parametrize("tag", ["k", "d"])
def test_transbsdf_lanes_match_reference_source_and_oracle(oracle32, tag):
    from materialist_b200.myutils.mi_plugin import TransBSDF
    g = load_golden("trans_bsdf.npz"); m0 = load_golden("matdiff_bsdf.npz")
    props = {"ior": float(g[tag + "_ior"])}
    if float(g[tag + "_refract_distance"]) == 100.0:
        props["keep_albedo_color"] = True
    b = TransBSDF(props)
    assert b.refract_distance == float(g[tag + "_refract_distance"])
    b.a, b.r, b.m, b.bg = (torch.from_numpy(g[k]).cuda() for k in ("a", "r", "m", "bg"))
    b.mask = torch.from_numpy(g["mask"]).cuda(); b.specTrans = float(g[tag + "_specTrans"])
    si = _si(g, m0)
    wi_w = si.to_world(si.wi)
    sc = b.calculate_refracted_screen_coor(wi_w, si.n, 1.0 / b.ior, si.p).cpu().numpy()
    assert np.abs(sc - g[tag + "_refr_screen"]).max() < 2e-3
    O = oracle32; cfg = cfg512(O)
    try:
        O.set_trans(b.ior, b.specTrans, b.refract_distance, g["bg"], g["mask"])
        sc_o, flat_o = O.trans_refracted_texel(cfg, g["p"], g["n"], wi_w.cpu().numpy())
    finally:
        O.set_trans(bg=None)
    assert np.array_equal(sc, sc_o)                                           # integer-deciding arithmetic: bit-exact vs the oracle
    wo_local = si.to_local(torch.from_numpy(m0["wo_world_used"]).cuda())
    f, pdf = b.eval_pdf(None, si, wo_local)
    check_lanes(f.cpu().numpy(), g[tag + "_eval_f"], "eval f", p99=5e-5); check_lanes(pdf.cpu().numpy(), g[tag + "_eval_pdf"], "eval pdf", p99=5e-5)
    bs, w = b.sample(None, si, torch.from_numpy(g["s1"]).cuda(), torch.from_numpy(g["s2"]).cuda())
    assert bs.eta == b.ior
    check_dirs(bs.wo.cpu().numpy(), g[tag + "_sample_wo"], "wo")
    check_lanes(w.cpu().numpy(), g[tag + "_sample_weight"], "weight", p99=1e-4, worst=5e-3)
    sampled_pdf_ok(bs.pdf.cpu().numpy(), g[tag + "_sample_pdf"], g[tag + "_sample_weight"])


def test_matdiffbsdf_lanes_match_reference_source():
    """The CUDA lane kernels against the reference's own MatDiffBSDF source (executed on the numpy Dr.Jit stand-ins)."""
    from materialist_b200.myutils.mi_plugin import MatDiffBSDF
    g = load_golden("matdiff_bsdf.npz")
    b = MatDiffBSDF({})
    b.a, b.r, b.m = (torch.from_numpy(g[k]).cuda() for k in ("a", "r", "m"))
    si = _si(g, g)
    f, pdf = b.eval_pdf(None, si, si.to_local(torch.from_numpy(g["wo_world_used"]).cuda()))
    check_lanes(f.cpu().numpy(), g["eval_f"], "eval f", p99=5e-5); check_lanes(pdf.cpu().numpy(), g["eval_pdf"], "eval pdf", p99=5e-5)
    bs, w = b.sample(None, si, torch.from_numpy(g["s1"]).cuda(), torch.from_numpy(g["s2"]).cuda())
    check_dirs(bs.wo.cpu().numpy(), g["sample_wo"], "wo")
    check_lanes(w.cpu().numpy(), g["sample_weight"], "weight", p99=1e-4, worst=5e-3)
    sampled_pdf_ok(bs.pdf.cpu().numpy(), g["sample_pdf"], g["sample_weight"])


def _edit_inputs(H, W, seed=21):
    rs = np.random.RandomState(seed)
    bg = rs.rand(H, W, 3).astype(np.float32)
    yy, xx = np.mgrid[0:H, 0:W]
    mask = ((xx - W * 0.45) ** 2 + (yy - H * 0.55) ** 2) < (0.3 * H) ** 2          # a disc: edited and unedited pixels, an edge
    return bg, mask


@pytest.mark.parametrize("gaussian,keep,max_depth", [(True, True, 4), (False, False, 2)])
def test_trans_gbuffer_render_matches_oracle(oracle32, gaussian, keep, max_depth):
    import materialist_b200 as mb
    O = oracle32
    c = Case(H=48, W=48, spp=32, He=8, We=16, gaussian=gaussian, max_depth=max_depth, sun=50.0)
    bg, mask = _edit_inputs(c.H, c.W)
    ior, st, dist = (1.2, 0.4, 100.0) if keep else (1.3, 0.8, 1.0)
    try:
        O.set_trans(ior, st, dist, bg, mask)
        ref = c.oracle_fwd(O, 11)
    finally:
        O.set_trans(bg=None)
    plain = c.oracle_fwd(O, 11)
    s = c.scene().set_bsdf({"name": "TransBSDF", "ior": ior, **({"keep_albedo_color": True} if keep else {})})
    p = mb.traverse(s)
    p["shape.bsdf.bg"] = torch.from_numpy(bg).cuda(); p["shape.bsdf.mask"] = torch.from_numpy(mask).cuda()
    p["shape.bsdf.specTrans"] = st; p["shape.bsdf.ior"] = ior
    p.update()
    a, r, m, n = c.torch_maps()
    with torch.no_grad():
        img = mb.render(s, spp=c.spp, seed=11, albedo=a, roughness=r, metallic=m).cpu().numpy()
    assert rel_l2(img, ref) <= 1e-4, rel_l2(img, ref)
    assert rel_l2(ref, plain) > 1e-2                                           # the edit is visible ...
    yy, xx = np.mgrid[0:c.H, 0:c.W]
    outside = ((xx - c.W * 0.45) ** 2 + (yy - c.H * 0.55) ** 2) > (0.3 * c.H + 4) ** 2
    # ... and confined to the mask (+ film footprint): outside only the TransBSDF's 1e-4 epsilons (pdf clamp, weight) differ from MatDiffBSDF's 1e-6
    assert rel_l2(ref[outside], plain[outside]) < 1e-3
    # forward only: asking for gradients raises
    a.requires_grad_(True)
    with pytest.raises(RuntimeError):
        mb.render(s, spp=4, seed=1, albedo=a, roughness=r, metallic=m).sum().backward()
    # without the TransBSDF the trans parameters do not exist
    with pytest.raises(KeyError):
        mb.traverse(c.scene())["shape.bsdf.bg"] = torch.from_numpy(bg).cuda()


def test_trans_mesh_render_matches_oracle(oracle32):
    import materialist_b200 as mb
    from test_gpu_mesh_parity import _scene, _cuda_scene, assert_radiance_parity
    from test_reference_render_pin import pin_cfg
    O = oracle32
    H = W = 40
    cam, verts, tris, a, r, m, env = _scene(H, W)
    bg, mask = _edit_inputs(H, W)
    om = O.mesh_create(verts, tris)
    env_int, hier, d = O.env_prepare(env, orc.ENV_ASSIGNED)
    cfg = pin_cfg(d, 5, 0, H, spp=32, H=H, W=W, max_depth=4, flags=REF_FLAGS)
    try:
        O.set_trans(1.2, 0.4, 1.0, bg, mask)
        ref, st = O.mesh_render_fwd(cfg, om, a, r, m, None, env_int, hier, d, want_stats=True)
    finally:
        O.set_trans(bg=None)
    plain = O.mesh_render_fwd(cfg, om, a, r, m, None, env_int, hier, d)
    O.mesh_destroy(om)
    s = _cuda_scene(cam, verts, tris, env, REF_FLAGS).set_bsdf({"name": "TransBSDF", "ior": 1.2})
    p = mb.traverse(s)
    p["shape.bsdf.bg"] = torch.from_numpy(bg).cuda(); p["shape.bsdf.mask"] = torch.from_numpy(mask).cuda(); p["shape.bsdf.specTrans"] = 0.4
    p.update()
    ta, tr, tm = (torch.from_numpy(x).cuda() for x in (a, r, m))
    with torch.no_grad():
        img = mb.render(s, spp=32, seed=5, albedo=ta, roughness=tr, metallic=tm).cpu().numpy()
    assert_radiance_parity(img, ref, 32, shape=(H, W))
    assert rel_l2(ref, plain) > 1e-2


def test_transprancy_edit_entry_point():
    """trans_edit.transprancy_edit: material overrides inside the mask + average over seeds 0..n-1."""
    import materialist_b200 as mb
    from materialist_b200 import trans_edit
    c = Case(H=32, W=32, spp=16, He=8, We=16)
    bg, mask = _edit_inputs(32, 32)
    a, r, m, _ = c.torch_maps()
    mat = {"albedo": a, "roughness": r, "metallic": m, "mask": torch.from_numpy(mask).cuda(), "bg": torch.from_numpy(bg).cuda(),
           "envmap": torch.from_numpy(c.env).cuda()}
    s = c.scene().set_bsdf({"name": "TransBSDF", "ior": 1.2, "keep_albedo_color": False})
    img = trans_edit.transprancy_edit(s, mat, 1.2, False, 0.4, n_iter=3, spp=16)
    ea, er, em = trans_edit.edit_materials(mat, False)
    tm_ = torch.from_numpy(mask).cuda()
    assert torch.all(ea[tm_] == 0.7) and torch.allclose(er[tm_], torch.tensor(0.3, device="cuda")) and torch.all(em[tm_] == 0) and torch.equal(ea[~tm_], a[~tm_])
    with torch.no_grad():
        one = sum(mb.render(s, spp=16, seed=i) for i in range(3)) / 3
    assert torch.equal(img, one) and img.shape == (32, 32, 3) and torch.isfinite(img).all()
    with pytest.raises(ValueError):
        trans_edit.transprancy_edit(c.scene(), mat, 1.2, False, 0.4)
