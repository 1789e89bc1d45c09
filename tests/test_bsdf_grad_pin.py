"""Adjoint pin (CPU, -m "not gpu"): the oracle's hand-derived eval_brdf_grad against the Jacobian of the REFERENCE'S OWN
MatDiffBSDF.eval_pdf source (myutils/mi_plugin.py:1372-1427, :1449-1460), taken by float64 central differences through the
numpy Dr.Jit stand-ins (tests/golden/make_bsdf_grad_golden.py -> matdiff_bsdf_grad.npz).  For a cotangent w on the rgb value,
Mitsuba's dr.backward delivers J^T w to (albedo, roughness, metallic[, normal]) at the lane's texel; that is what the adjoint
render accumulates.  The CUDA twin is tests/test_gpu_aux_parity.py::test_cuda_bsdf_grad_matches_reference_source."""
import numpy as np
import pytest

from test_bsdf_plugin_golden import cfg512, load_golden


def grad_case(g, tag):
    """Inputs + reference J^T w for the lanes kept by the generator (one lane per texel)."""
    keep = g[f"{tag}_keep"]
    J = g[f"{tag}_J"].astype(np.float64)                         # (L, 5 or 8, 3)
    w = np.random.RandomState(5).randn(len(keep), 3).astype(np.float32)
    ref = np.einsum("lkc,lc->lk", J, w.astype(np.float64))      # (L, 5 or 8)
    n_map = None
    if tag == "nmap":
        H, W = int(g["H"]), int(g["W"])
        n_map = np.zeros((H, W, 3), np.float32); n_map[..., 2] = 1.0
        n_map.reshape(-1, 3)[g["nmap_flat"]] = g["nmap_normals"]
    lanes = {k: g[k][keep] for k in ("p", "n", "wi_world_used", "wo_world_used")}
    return lanes, w, ref, J, n_map


def check_grad(got, ref, J, w, name):
    """`got`, `ref`: (L, K).  Error measured against the size of the terms being summed (|J|^T |w|): the cotangent is random, so a
    lane's J^T w can cancel to ~0 while its terms are O(1)."""
    scale = np.einsum("lkc,lc->lk", np.abs(J), np.abs(w.astype(np.float64))) + 1e-12
    e = np.abs(got.astype(np.float64) - ref) / scale
    live = scale > 1e-9
    assert live.mean() > 0.5, name
    assert np.median(e[live]) <= 2e-6 and np.percentile(e[live], 99) <= 2e-4 and np.percentile(e[live], 99.9) <= 5e-3, \
        (name, np.median(e[live]), np.percentile(e[live], 99), np.percentile(e[live], 99.9), e[live].max())
    tot = np.linalg.norm(got - ref) / np.linalg.norm(ref)
    assert tot <= 1e-3, (name, tot)


@pytest.mark.parametrize("tag", ["mesh", "nmap"])
def test_oracle_bsdf_grad_matches_reference_source(oracle32, tag):
    g = load_golden("matdiff_bsdf_grad.npz")
    lanes, w, ref, J, n_map = grad_case(g, tag)
    cfg = cfg512(oracle32)
    cfg.use_mesh_normal = int(tag == "mesh")
    ga, gr, gm, gn = oracle32.bsdf_eval_grad(cfg, lanes["p"], lanes["n"], lanes["wi_world_used"], lanes["wo_world_used"],
                                             g["a"], g["r"], g["m"], w, n_opt=n_map)
    got = np.concatenate([ga, gr[:, None], gm[:, None]] + ([gn] if tag == "nmap" else []), -1)
    check_grad(got, ref, J, w, tag)
    # the value the Jacobian was taken around is the one the forward pin already covers
    f, _ = oracle32.bsdf_eval_pdf(cfg, lanes["p"], lanes["n"], lanes["wi_world_used"], lanes["wo_world_used"], g["a"], g["r"], g["m"], n_opt=n_map)
    e = np.abs(f - g[f"{tag}_f"]) / np.maximum(np.abs(g[f"{tag}_f"]), 1e-4)
    assert np.percentile(e, 99) <= 2e-5
