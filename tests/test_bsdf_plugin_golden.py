"""CPU (-m "not gpu"): the oracle's MatDiffBSDF / TransBSDF against golden vectors produced by EXECUTING THE REFERENCE'S OWN
Dr.Jit-typed source (myutils/mi_plugin.py:1229-1770) on numpy stand-ins for the drjit / mitsuba array types
(tests/golden/make_bsdf_golden.py + drjit_np_shim.py; mitsuba itself is not installable here).

This pins what tests/golden/bsdf_terms.npz could not: the Disney-diffuse value, Fresnel blend, the sampling directions in
Mitsuba's Frame3f, the brdf/(pdf+eps) weights, mi_world_to_screen texel indices, and every TransBSDF formula including
the twice-refracted background lookup."""
import os

import numpy as np
import pytest

from helpers import Case

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    """npz -> dict; the (512,512,.) maps are stored as 64 x 64 blocks (tiled) and the mask as its checkerboard cell size."""
    g = dict(np.load(os.path.join(GOLD, name)))
    H, W = int(g["H"]), int(g["W"])
    for k in ("a", "r", "m", "bg"):
        if k + "_block" in g:
            b = g[k + "_block"]
            g[k] = np.ascontiguousarray(np.tile(b, (H // b.shape[0], W // b.shape[1], 1)))
    if "mask_cell" in g:
        yy, xx = np.mgrid[0:H, 0:W]; c = int(g["mask_cell"])
        g["mask"] = ((xx // c) + (yy // c)) % 2 == 0
    return g


def lane_stats(got, ref):
    e = np.abs(np.asarray(got, np.float64) - ref) / np.maximum(np.abs(ref), 1e-4)
    return float(np.median(e)), float(np.percentile(e, 99)), float(e.max())


def check_lanes(got, ref, name, p99=2e-5, worst=1e-3):
    """per-lane relative error: median, 99th percentile, and the worst lane (a glossy lane ON the GGX peak at r = 0.07
    amplifies 1 ulp of the half vector by ~1e4: `worst` bounds the 99.9th percentile, the single worst lane only loosely)"""
    e = np.abs(np.asarray(got, np.float64) - ref) / np.maximum(np.abs(ref), 1e-4)
    med, q99, q999, mx = float(np.median(e)), float(np.percentile(e, 99)), float(np.percentile(e, 99.9)), float(e.max())
    assert med <= 1e-6 and q99 <= p99 and q999 <= worst and mx <= 0.5, (name, med, q99, q999, mx)


def check_dirs(got, ref, name, tol=1e-3):
    """unit vectors: absolute error (a relative bar is meaningless on a component near 0).  A glossy reflection about a half
    vector at grazing incidence amplifies 1 ulp of the half vector ~1e3 times: bulk bar 1e-4 (99.9 %), worst lane `tol`."""
    e = np.abs(np.asarray(got, np.float64) - ref).max(-1)
    assert np.percentile(e, 99.9) <= 1e-4 and e.max() <= tol, (name, np.percentile(e, 99.9), e.max())


def cfg512(O):
    c = Case(H=512, W=512, spp=1, He=8, We=16)                 # default_cam.json is 512 x 512
    _, _, d = O.env_prepare(c.env)
    return c.cfg(d, 0)


def sampled_pdf_ok(got, ref, weight_ref):
    """The pdf of a SAMPLED glossy direction sits on the GGX peak (D up to 1e4 at r = 0.07): 1 ulp in the half vector moves it
    by ~1e-3, and for a direction sampled below the horizon (weight 0, path dead) it is not meaningful at all."""
    alive = weight_ref.max(-1) > 0
    e = np.abs(got - ref) / np.maximum(np.abs(ref), 1e-4)
    assert np.median(e) <= 1e-6 and np.percentile(e[alive], 99) <= 2e-3 and np.percentile(e[alive], 99.9) <= 5e-2 and e[alive].max() <= 0.5, \
        (np.median(e), np.percentile(e[alive], 99.9), e[alive].max())


def test_matdiffbsdf_matches_reference_source(oracle32):
    g = load_golden("matdiff_bsdf.npz")
    O = oracle32; cfg = cfg512(O)
    assert np.array_equal(np.array(cfg.view[:], np.float32).reshape(4, 4), g["view"])        # MatDiffBSDF.__init__ camera matrices
    assert np.array_equal(np.array(cfg.proj[:], np.float32).reshape(4, 4), g["proj"])
    sc, flat = O.world_to_screen(cfg, g["p"])
    assert np.abs(sc - g["screen"]).max() < 2e-3
    assert np.array_equal(flat, np.floor(g["screen"][:, 0]).astype(np.int64) + np.floor(g["screen"][:, 1]).astype(np.int64) * 512)
    f, pdf = O.bsdf_eval_pdf(cfg, g["p"], g["n"], g["wi_world_used"], g["wo_world_used"], g["a"], g["r"], g["m"])
    check_lanes(f, g["eval_f"], "eval f"); check_lanes(pdf, g["eval_pdf"], "eval pdf")
    assert (g["eval_f"].max(-1) == 0).mean() > 0.05 and (g["eval_f"].max(-1) > 0).mean() > 0.5     # both sides of the horizon exercised
    wo, pdf_s, w = O.bsdf_sample(cfg, g["p"], g["n"], g["wi_world_used"], g["s1"], g["s2"], g["a"], g["r"], g["m"])
    check_dirs(wo, g["sample_wo"], "sample wo"); check_lanes(w, g["sample_weight"], "sample weight")
    sampled_pdf_ok(pdf_s, g["sample_pdf"], g["sample_weight"])


@pytest.mark.parametrize("tag", ["k", "d"])       # k: ior 1.2, keep_albedo_color (refract_distance 100), specTrans 0.4; d: plugin defaults
def test_transbsdf_matches_reference_source(oracle32, tag):
    g = load_golden("trans_bsdf.npz"); m0 = load_golden("matdiff_bsdf.npz")
    O = oracle32; cfg = cfg512(O)
    wi_w, wo_w = m0["wi_world_used"], m0["wo_world_used"]
    assert float(g[tag + "_eta"]) == float(g[tag + "_ior"])                                        # bs.eta = self.ior (:1542)
    try:
        O.set_trans(g[tag + "_ior"], g[tag + "_specTrans"], g[tag + "_refract_distance"], g["bg"], g["mask"])
        sc, flat = O.trans_refracted_texel(cfg, g["p"], g["n"], wi_w)
        ref_sc = g[tag + "_refr_screen"]
        assert np.abs(sc - ref_sc).max() < 2e-3
        ref_flat = np.floor(ref_sc[:, 0]).astype(np.int64) + np.floor(ref_sc[:, 1]).astype(np.int64) * 512
        assert (flat != ref_flat).mean() <= 1e-3                                                      # a lane ON a texel boundary may round either way
        assert (ref_sc.min() >= 0) and (ref_sc.max() <= 511) and len(np.unique(ref_flat)) > 1000
        f, pdf = O.bsdf_eval_pdf(cfg, g["p"], g["n"], wi_w, wo_w, g["a"], g["r"], g["m"])
        check_lanes(f, g[tag + "_eval_f"], "eval f"); check_lanes(pdf, g[tag + "_eval_pdf"], "eval pdf")
        # the edit really changes the value on masked lanes and only there
        edited = g["mask"].reshape(-1)[np.floor(m0["screen"][:, 0]).astype(np.int64) + np.floor(m0["screen"][:, 1]).astype(np.int64) * 512]
        differs = np.abs(g[tag + "_eval_f"] - m0["eval_f"]).max(-1) > 1e-6
        assert differs[edited].mean() > 0.5 and not differs[~edited].any()
        wo, pdf_s, w = O.bsdf_sample(cfg, g["p"], g["n"], wi_w, g["s1"], g["s2"], g["a"], g["r"], g["m"])
        check_dirs(wo, g[tag + "_sample_wo"], "sample wo"); check_lanes(w, g[tag + "_sample_weight"], "sample weight")
        sampled_pdf_ok(pdf_s, g[tag + "_sample_pdf"], g[tag + "_sample_weight"])
    finally:
        O.set_trans(bg=None)
    f, _ = O.bsdf_eval_pdf(cfg, g["p"], g["n"], wi_w, wo_w, g["a"], g["r"], g["m"])                # mode really switched back
    check_lanes(f, m0["eval_f"], "matdiff after trans")
