"""CPU: the oracle's HAND-DERIVED adjoint against torch autograd on an independent float64 re-implementation
(tests/torch_mirror.py), and fp32-oracle vs fp64-oracle as the measure of fp32 round-off.  This is what lets the
GPU parity tests trust oracle gradients."""
import numpy as np
import pytest
import torch

import torch_mirror as tm
from helpers import Case, rel_l2
from oracle import oracle as orc


@pytest.mark.parametrize("mode", [orc.ENV_ASSIGNED, orc.ENV_FILE])
@pytest.mark.parametrize("gaussian", [True, False])
@pytest.mark.parametrize("use_mesh_normal", [True, False])
def test_oracle_forward_and_adjoint_match_autograd_mirror(oracle32, oracle64, mode, gaussian, use_mesh_normal):
    c = Case(H=12, W=12, spp=8, He=8, We=16, sun=30.0, invalid_border=1, env_mode=mode, gaussian=gaussian,
             use_mesh_normal=use_mesh_normal)
    seed = 7
    img64 = c.oracle_fwd(oracle64, seed)
    img32 = c.oracle_fwd(oracle32, seed)
    ta, tr, tm_, tn, te = [torch.tensor(x, dtype=torch.float64, requires_grad=True) for x in (c.a, c.r, c.m, c.n, c.env)]
    kw = dict(use_mesh_normal=use_mesh_normal, gaussian=gaussian)
    img_m = tm.render(oracle64, c.cam, c.gpos, c.gnrm, ta, tr, tm_, tn, te, mode == orc.ENV_FILE, c.spp, seed, **kw)
    assert rel_l2(img_m.detach().numpy(), img64) < 1e-6          # two independent implementations agree
    assert rel_l2(img32, img64) < 1e-5                           # fp32 round-off of the restatement
    # adjoint render: seed_grad, AD weights
    sg = oracle64.seed_grad(seed)
    G = np.random.RandomState(0).randn(c.H, c.W, 3).astype(np.float32)
    want = ("a", "r", "m", "n", "env")
    g64 = c.oracle_bwd(oracle64, sg, G, want=want)
    g32 = c.oracle_bwd(oracle32, sg, G, want=want)
    img_a = tm.render(oracle64, c.cam, c.gpos, c.gnrm, ta, tr, tm_, tn, te, mode == orc.ENV_FILE, c.spp, sg, ad_weights=True, **kw)
    (img_a * torch.tensor(G, dtype=torch.float64)).sum().backward()
    for k, t in (("a", ta), ("r", tr), ("m", tm_), ("n", tn), ("env", te)):
        if k == "n" and use_mesh_normal:
            assert np.abs(g64["n"]).max() == 0
            continue
        assert rel_l2(t.grad.numpy(), g64[k]) < 1e-5, k          # hand-derived adjoint == autograd
        assert rel_l2(g32[k], g64[k]) < 2e-4, k                  # fp32 oracle within the 1e-3 gradient bar


def test_forward_primal_of_ad_pass_differs_only_by_weights(oracle64):
    """P6: the AD pass re-evaluates the BSDF weight f2/p2 instead of f/(p+1e-6): images agree to ~1e-6*."""
    c = Case(H=10, W=10, spp=16, He=8, We=16, sun=0.0, gaussian=False, flags=orc.FLAG_ENV_HALF_TEXEL)   # no wo-quirk: f2/p2 == f/p
    a = c.oracle_fwd(oracle64, 3)
    b = c.oracle_fwd(oracle64, 3, extra_flags=orc.FLAG_AD_WEIGHTS)
    assert rel_l2(a, b) < 1e-4 and not np.array_equal(a, b)


def test_hierarchy_sampling_is_a_density(oracle32):
    """Hierarchical2D: E[1/pdf] = 1 over the unit square, sampled pdf == bilinear eval, patch masses match."""
    rs = np.random.RandomState(1)
    data = (0.2 + np.exp(rs.randn(17, 33))).astype(np.float32)
    hier, d = oracle32.hier_build(data)
    s = rs.rand(400000, 2).astype(np.float32)
    uv, pdf, off = oracle32.hier_sample(hier, d, s)
    assert abs((1.0 / pdf).mean() - 1.0) < 5e-3
    np.testing.assert_allclose(oracle32.hier_eval(hier, d, uv), pdf, rtol=2e-4, atol=1e-5)
    assert off[:, 0].max() <= 31 and off[:, 1].max() <= 15 and off.min() >= 0
    # patch histogram ~ patch integrals (level 1)
    hist = np.zeros((16, 32)); np.add.at(hist, (off[:, 1], off[:, 0]), 1.0); hist /= hist.sum()
    mass = 0.25 * (data[:-1, :-1] + data[:-1, 1:] + data[1:, :-1] + data[1:, 1:]); mass /= mass.sum()
    assert np.abs(hist - mass).max() < 6e-4


def test_shard_rows_reproduce_full_image(oracle32):
    c = Case(H=20, W=16, spp=8, He=8, We=16, gaussian=True)
    full = c.oracle_fwd(oracle32, 5)
    parts = [c.oracle_fwd(oracle32, 5, row0=r0, rows=n) for r0, n in ((0, 7), (7, 6), (13, 7))]
    assert np.array_equal(np.concatenate(parts, 0), full)
