"""-m gpu: the CUDA path (through the C-ABI) against the CPU oracle on identical inputs and sample sequences.
Bars (BASELINE.json north_star): radiance <= 1e-4 rel-L2, gradients <= 1e-3 rel-L2, sample / CDF / texel indices bit-exact."""
import numpy as np
import pytest
import torch

from helpers import Case, rel_l2
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

TOL_RADIANCE = 1e-4
TOL_GRAD = 1e-3

CASES = {
    "gauss_assigned": dict(H=48, W=48, spp=64, He=16, We=32, gaussian=True, env_mode=orc.ENV_ASSIGNED),
    "gauss_file_sun": dict(H=40, W=40, spp=32, He=32, We=64, gaussian=True, env_mode=orc.ENV_FILE),
    "box_ragged_spp": dict(H=33, W=33, spp=19, He=17, We=31, gaussian=False, env_mode=orc.ENV_ASSIGNED),
    "gauss_invalid_border_nmap": dict(H=36, W=36, spp=40, He=16, We=32, gaussian=True, invalid_border=3, use_mesh_normal=False),
    "gauss_nonsquare": dict(H=24, W=40, spp=32, He=16, We=32, gaussian=True),
    "spp_below_warp": dict(H=20, W=20, spp=5, He=8, We=16, gaussian=True),
    # 8 lanes per pixel (four pixels per warp): pixel count not a multiple of 4, spp = the group width / a single sample
    "lpp_ragged_pixels_nmap": dict(H=21, W=21, spp=8, He=16, We=32, gaussian=True, use_mesh_normal=False, invalid_border=2),
    "spp_one_box": dict(H=18, W=18, spp=1, He=8, We=16, gaussian=False),
    "spp_above_128": dict(H=16, W=16, spp=160, He=16, We=32, gaussian=True),          # adjoint back on a warp per pixel
}


@pytest.mark.parametrize("name", list(CASES))
def test_forward_radiance(name, oracle32):
    import materialist_b200 as mb
    c = Case(**CASES[name])
    ref = c.oracle_fwd(oracle32, seed=7)
    s = c.scene()
    a, r, m, n = c.torch_maps()
    img = mb.render(s, spp=c.spp, seed=7, albedo=a, roughness=r, metallic=m, normal=n).cpu().numpy()
    assert np.isfinite(img).all()
    assert rel_l2(img, ref) <= TOL_RADIANCE, rel_l2(img, ref)


@pytest.mark.parametrize("name", list(CASES))
def test_indices_bit_exact(name, oracle32):
    import materialist_b200 as mb
    c = Case(**CASES[name])
    _, idx_ref = c.oracle_fwd(oracle32, seed=11, want_indices=True)
    s = c.scene()
    a, r, m, n = c.torch_maps()
    s.r = r
    idx = mb.sample_indices(s, c.spp, 11).cpu().numpy()
    assert idx.shape == idx_ref.shape
    assert (idx != idx_ref).sum() == 0, f"{(idx != idx_ref).any(axis=1).sum()} lanes differ"


_REC_NAMES = ("hier off.x", "hier off.y", "texel", "lobe", "emitter cell", "bsdf-direction cell",
              "d_em.x", "d_em.y", "d_em.z", "d_bs.x", "d_bs.y", "d_bs.z")


def _record_both(c, O, seed, ad=False, row0=0, rows=None):
    import materialist_b200 as mb
    env_int, hier, d = O.env_prepare(c.env, c.env_mode)
    cfg = c.cfg(d, seed, row0, rows, extra_flags=orc.FLAG_AD_WEIGHTS if ad else 0)
    ref, ref_L = O.sample_record(cfg, c.gpos, c.gnrm, c.a, c.r, c.m, None if c.use_mesh_normal else c.n, env_int, hier, d, want_radiance=True)
    s = c.scene()
    a, r, m, n = c.torch_maps()
    s.a, s.r, s.m, s.n = a, r, m, n
    if rows is not None:
        s.set_shard(row0, rows)
    got, got_L = mb.renderop.sample_record(s, c.spp, seed, ad_weights=ad, want_radiance=True)
    return ref, ref_L, got.cpu().numpy(), got_L.cpu().numpy()


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("ad", [False, True])
def test_sample_record_bit_exact(name, ad, oracle32):
    """The PRODUCTION forward sample function (shade_sample, instrumented instantiation) against the oracle, lane by lane: every
    integer decision — hierarchy cell, texel, lobe, the envmap cell of the emitter sample AND of the BSDF-sampled direction — and
    the IEEE bits of both sampled directions are equal; the lane's radiance agrees to float rounding."""
    c = Case(**CASES[name])
    ref, ref_L, got, got_L = _record_both(c, oracle32, 11, ad)
    for k, nm in enumerate(_REC_NAMES):
        bad = np.flatnonzero(ref[:, k] != got[:, k])
        assert bad.size == 0, (nm, bad.size, bad[:5], ref[bad[:5], k], got[bad[:5], k])
    err = np.abs(got_L - ref_L).max(-1) / np.maximum(np.abs(ref_L).max(-1), 1e-3)
    assert np.percentile(err, 99.9) < 2e-5 and err.max() < 1e-3, (np.percentile(err, 99.9), err.max())


@pytest.mark.parametrize("name", list(CASES))
def test_backward_gradients(name, oracle32):
    import materialist_b200 as mb
    c = Case(**CASES[name])
    s = c.scene()
    a, r, m, n = c.torch_maps(requires_grad=True)
    env = torch.from_numpy(c.env).cuda().requires_grad_(True)
    assigned = c.env_mode == orc.ENV_ASSIGNED
    img = mb.render(s, spp=c.spp, seed=7, albedo=a, roughness=r, metallic=m, normal=n, envmap=env if assigned else None)
    G = np.random.RandomState(0).randn(c.H, c.W, 3).astype(np.float32)
    img.backward(torch.from_numpy(G).cuda())
    want = ("a", "r", "m", "env") + (() if c.use_mesh_normal else ("n",))
    ref = c.oracle_bwd(oracle32, mb.default_seed_grad(7), G, want=want)
    assert rel_l2(a.grad.cpu().numpy(), ref["a"]) <= TOL_GRAD
    assert rel_l2(r.grad.cpu().numpy(), ref["r"]) <= TOL_GRAD
    assert rel_l2(m.grad.cpu().numpy(), ref["m"]) <= TOL_GRAD
    if not c.use_mesh_normal:
        assert rel_l2(n.grad.cpu().numpy(), ref["n"]) <= TOL_GRAD
    if assigned:
        assert rel_l2(env.grad.cpu().numpy(), ref["env"]) <= TOL_GRAD


def test_env_hierarchy_bit_exact(oracle32):
    """The GPU-built Hierarchical2D pyramid equals the oracle's bit for bit (incl. a 'sun' stress envmap)."""
    import materialist_b200 as mb
    from materialist_b200 import synthetic
    for (He, We, mode) in ((16, 32, orc.ENV_ASSIGNED), (16, 32, orc.ENV_FILE), (37, 53, orc.ENV_ASSIGNED), (128, 256, orc.ENV_FILE)):
        env = synthetic.envmap(He, We).numpy()
        env_int, hier, d = oracle32.env_prepare(env, mode)
        p, nrm, v = synthetic.gbuffer(8, 8)
        s = mb.Scene(p, nrm, v, envmap=torch.from_numpy(env))
        s.set_envmap(torch.from_numpy(env), mode)
        env4, hier_g, desc, *_ = s.prepared_env()
        assert desc.total_floats == d.total_floats and list(desc.lvl_off) == list(d.lvl_off)
        assert np.array_equal(env4.cpu().numpy()[..., :3], env_int)
        hg = hier_g.cpu().numpy()
        assert np.array_equal(hg.view(np.uint32), hier.view(np.uint32)), f"{(hg != hier).sum()} of {hier.size} differ"


def test_shards_equal_full_image(oracle32):
    """N-shard result == 1-shard result: bitwise for the image (fixed-order film gather)."""
    import materialist_b200 as mb
    c = Case(H=40, W=40, spp=32, He=16, We=32, gaussian=True)
    a, r, m, n = c.torch_maps()
    s = c.scene()
    full = mb.render(s, spp=c.spp, seed=3, albedo=a, roughness=r, metallic=m).cpu().numpy()
    parts = []
    for row0, rows in ((0, 13), (13, 14), (27, 13)):
        s.set_shard(row0, rows)
        parts.append(mb.render(s, spp=c.spp, seed=3, albedo=a, roughness=r, metallic=m).cpu().numpy())
    assert np.array_equal(np.concatenate(parts, 0), full)


def test_host_pipelined_front_end_equals_direct_calls():
    """materialist_b200.hostpipe: copies on side streams, double-buffered inputs — same images and gradients as direct calls."""
    import materialist_b200 as mb
    from materialist_b200.hostpipe import HostPipelinedRenderWBRDF
    c = Case(H=40, W=40, spp=32, He=16, We=32, gaussian=True)
    s = c.scene()
    sets = []
    for k in range(3):                                                        # three different input sets through two buffer sets
        a, r, m = (t.numpy() for t in __import__("materialist_b200").synthetic.materials(c.H, c.W, seed_base=20 + k))
        g = np.random.RandomState(k).randn(c.H, c.W, 3).astype(np.float32)
        sets.append(tuple(torch.from_numpy(x).pin_memory() for x in (a, r, m, g)))
    outs = [tuple(torch.empty(*sh).pin_memory() for sh in ((c.H, c.W, 3), (c.H, c.W, 3), (c.H, c.W, 1), (c.H, c.W, 1))) for _ in range(3)]
    pipe = HostPipelinedRenderWBRDF(s, c.spp)
    pipe.stage(0, *sets[0])
    for k in range(3):
        pipe.step(50 + k, k % 2, *outs[k], next_inputs=sets[k + 1] if k < 2 else None)
    pipe.synchronize()
    for k in range(3):
        a, r, m, g = (t.cuda() for t in sets[k])
        a.requires_grad_(True); r.requires_grad_(True); m.requires_grad_(True)
        img = mb.render(s, spp=c.spp, seed=50 + k, albedo=a, roughness=r, metallic=m)
        img.backward(g)
        assert torch.equal(outs[k][0], img.detach().cpu())
        for got, ref in zip(outs[k][1:], (a.grad, r.grad, m.grad)):
            assert rel_l2(got.numpy(), ref.cpu().numpy()) < 1e-5          # float atomics: order-dependent in the last bits


def test_host_pipelined_front_end_row_shard():
    """With a row shard the pipe moves only the shard's rows (+ the 2-row film halo of the maps): same image rows and gradients as
    the direct sharded call with whole maps."""
    import materialist_b200 as mb
    from materialist_b200.hostpipe import HostPipelinedRenderWBRDF
    c = Case(H=40, W=40, spp=32, He=16, We=32, gaussian=True)
    s = c.scene()
    row0, rows = 13, 14
    a, r, m = (torch.from_numpy(x) for x in (c.a, c.r, c.m))
    g = torch.from_numpy(np.random.RandomState(3).randn(rows, c.W, 3).astype(np.float32))
    # stand-in for ShardContext.halo_exchange in one process: the neighbours' image-gradient rows are zero
    halo = lambda gr: torch.cat([torch.zeros_like(gr[:2]), gr, torch.zeros_like(gr[:2])], 0)
    with s.shard(row0, rows):
        pipe = HostPipelinedRenderWBRDF(s, c.spp, halo_exchange=halo)
        (m0, m1), (o0, o1) = pipe.map_rows, pipe.out_rows
        assert (m0, m1) == (row0 - 2, row0 + rows + 2) and (o0, o1) == (row0, row0 + rows)
        ins = tuple(t[m0:m1].contiguous().pin_memory() for t in (a, r, m)) + (g.pin_memory(),)
        outs = tuple(torch.empty(*sh).pin_memory() for sh in ((rows, c.W, 3), (rows, c.W, 3), (rows, c.W, 1), (rows, c.W, 1)))
        pipe.stage(0, *ins)
        pipe.step(77, 0, *outs)
        pipe.synchronize()
        da, dr, dm = (t.cuda().requires_grad_(True) for t in (a, r, m))
        img = mb.render(s, spp=c.spp, seed=77, albedo=da, roughness=dr, metallic=dm, halo_exchange=halo)
        img.backward(g.cuda())
    assert (s.row0, s.rows) == (0, c.H)
    assert torch.equal(outs[0], img.detach().cpu())
    for got, ref in zip(outs[1:], (da.grad, dr.grad, dm.grad)):
        assert rel_l2(got.numpy(), ref[row0:row0 + rows].cpu().numpy()) < 1e-5
        assert float(ref[:row0].abs().sum() + ref[row0 + rows:].abs().sum()) == 0.0      # G-buffer mode: gradients stay in the shard's rows


def test_no_cpu_fallback():
    import materialist_b200 as mb
    c = Case(H=8, W=8, spp=4, He=8, We=16)
    s = c.scene()
    a, r, m, n = c.torch_maps(device="cpu")
    with pytest.raises(ValueError):
        mb.render(s, spp=4, seed=1, albedo=a, roughness=r, metallic=m)
