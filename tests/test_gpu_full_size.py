"""-m gpu: parity and size-independent properties at BASELINE.json's FULL sizes (C2: 512x512, 64 spp, env 256x128;
C5: 3840x2160, 256 spp, env 2048x1024).  The oracle renders a few rows of the same image (global lane ids make any
row block reproduce the full-image samples); the properties need no oracle at all."""
import numpy as np
import pytest
import torch

from helpers import Case, rel_l2
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def _render_with_grads(c, seed, G, want_env=True):
    import materialist_b200 as mb
    s = c.scene()
    a, r, m, n = c.torch_maps(requires_grad=True)
    env = torch.from_numpy(c.env).cuda().requires_grad_(want_env)
    img = mb.render(s, spp=c.spp, seed=seed, albedo=a, roughness=r, metallic=m, envmap=env)
    img.backward(G)
    return s, img.detach(), a.grad, r.grad, m.grad, env.grad


def test_c2_full_size_rows_vs_oracle(oracle32):
    """C2 at full size on the GPU; rows [248, 264) re-rendered by the oracle, forward and adjoint (material gradients are
    per pixel, so the row block of the GPU gradient must equal the oracle's gradient of the same rows)."""
    import materialist_b200 as mb
    c = Case(H=512, W=512, spp=64, He=128, We=256)
    row0, rows = 248, 16
    Gn = np.zeros((c.H, c.W, 3), np.float32)
    Gn[row0 - 2:row0 + rows + 2] = np.random.RandomState(1).randn(rows + 4, c.W, 3).astype(np.float32)
    s, img, ga, gr, gm, genv = _render_with_grads(c, 7, torch.from_numpy(Gn).cuda(), want_env=False)
    ref = c.oracle_fwd(oracle32, 7, row0=row0, rows=rows)
    assert rel_l2(img[row0:row0 + rows].cpu().numpy(), ref) <= 1e-4
    gref = c.oracle_bwd(oracle32, mb.default_seed_grad(7), Gn, want=("a", "r", "m"), row0=row0, rows=rows)
    sl = slice(row0, row0 + rows)
    for got, key in ((ga, "a"), (gr, "r"), (gm, "m")):
        assert rel_l2(got[sl].cpu().numpy(), gref[key][sl]) <= 1e-3, key


def test_c2_full_size_properties():
    """Determinism of the image (bitwise), shard invariance (bitwise), linearity of the adjoint in the image gradient, and the
    adjoint identity  <g_env, env> = <G, primal of the AD pass>  (radiance is linear in the texels once the sampling
    densities are detached, as Mitsuba detaches them)."""
    import materialist_b200 as mb
    from materialist_b200 import _abi, renderop as mbr
    c = Case(H=512, W=512, spp=64, He=128, We=256)
    G = torch.from_numpy(np.random.RandomState(2).rand(c.H, c.W, 3).astype(np.float32)).cuda()
    s, img, ga, gr, gm, genv = _render_with_grads(c, 5, G)
    s2, img2, ga2, *_rest = _render_with_grads(c, 5, 2.0 * G)
    assert torch.equal(img, img2)                                                       # bitwise deterministic forward
    assert float((ga2 - 2 * ga).norm() / (2 * ga).norm()) < 1e-5                        # adjoint is linear in G (atomics reorder only)
    a, r, m, n = c.torch_maps()
    parts = []
    for row0, rows in ((0, 200), (200, 112), (312, 200)):
        s.set_shard(row0, rows)
        parts.append(mb.render(s, spp=c.spp, seed=5, albedo=a, roughness=r, metallic=m))
    assert torch.equal(torch.cat(parts, 0), img)                                        # shards == full image, bitwise
    s.set_shard(0, c.H)
    # adjoint identity with the AD-pass primal (seed_grad, AD weights)
    env_pack = s.prepared_env()
    img_ad = mbr._forward(s, c.spp, mb.default_seed_grad(5), a, r, m, None, env_pack, extra_flags=_abi.FLAG_AD_WEIGHTS)
    lhs = float((genv.double() * torch.from_numpy(c.env).cuda().double()).sum())
    rhs = float((G.double() * img_ad.double()).sum())
    assert abs(lhs - rhs) <= 2e-4 * abs(rhs), (lhs, rhs)


def test_c2_full_size_sample_record_bit_exact(oracle32):
    """C2 at full size: 32 rows (1 M lanes) of the decision record of the production forward sample function == the oracle's, word
    for word (hierarchy cell, texel, lobe, both envmap cells, the bits of both sampled directions)."""
    from test_gpu_render_parity import _record_both, _REC_NAMES
    c = Case(H=512, W=512, spp=64, He=128, We=256)
    ref, ref_L, got, got_L = _record_both(c, oracle32, 7, row0=240, rows=32)
    for k, nm in enumerate(_REC_NAMES):
        assert np.array_equal(ref[:, k], got[:, k]), (nm, int((ref[:, k] != got[:, k]).sum()))


def test_c5_full_size_rows_vs_oracle_and_adjoint_identity(oracle32):
    """C5 (4K, 256 spp, 2048x1024 per-texel-noise envmap with a sun, the SURVEY §8d generator): 4 rows against the oracle at the
    north-star tolerances — radiance 1e-4, material gradients 1e-3 — the decision record of those rows bit-exact, and the adjoint
    identity over the whole image.  (Round 1 met only 3e-4 here: glibc's and CUDA's atan2f / sincospif differ by ulps, which the
    2047-texel-wide white-noise map and the GGX peak amplify; the direction chain is now bit-identical to the oracle's.)"""
    import materialist_b200 as mb
    from materialist_b200 import _abi, renderop as mbr
    from test_gpu_render_parity import _record_both, _REC_NAMES
    c = Case(H=2160, W=3840, spp=256, He=1024, We=2048)
    row0, rows = 1078, 4
    s = c.scene()
    a, r, m, n = c.torch_maps(requires_grad=True)
    env = torch.from_numpy(c.env).cuda().requires_grad_(True)
    img = mb.render(s, spp=c.spp, seed=3, albedo=a, roughness=r, metallic=m, envmap=env)
    ref = c.oracle_fwd(oracle32, 3, row0=row0, rows=rows)
    err = rel_l2(img[row0:row0 + rows].detach().cpu().numpy(), ref)
    assert err <= 1e-4, err
    g = torch.Generator(device="cuda").manual_seed(0)
    G = torch.rand(c.H, c.W, 3, device="cuda", generator=g)
    img.backward(G)
    with torch.no_grad():
        img_ad = mbr._forward(s, c.spp, mb.default_seed_grad(3), a.detach(), r.detach(), m.detach(), None, s.prepared_env(), extra_flags=_abi.FLAG_AD_WEIGHTS)
    lhs = float((env.grad.double() * env.detach().double()).sum())
    rhs = float((G.double() * img_ad.double()).sum())
    assert abs(lhs - rhs) <= 5e-4 * abs(rhs), (lhs, rhs)
    # material gradients of the sampled rows against the oracle (per-pixel quantities: the row block of the GPU gradient must equal
    # the oracle's gradient of the same rows, given the same image gradient on the rows + film halo)
    Gn = G.cpu().numpy()
    gref = c.oracle_bwd(oracle32, mb.default_seed_grad(3), Gn, want=("a", "r", "m"), row0=row0, rows=rows)
    sl = slice(row0, row0 + rows)
    for got, key in ((a.grad, "a"), (r.grad, "r"), (m.grad, "m")):
        e = rel_l2(got[sl].cpu().numpy(), gref[key][sl])
        assert e <= 1e-3, (key, e)
    # decisions of those rows, bit-exact (1 row = 983 040 lanes)
    rref, _, rgot, _ = _record_both(c, oracle32, 3, row0=row0, rows=1)
    for k, nm in enumerate(_REC_NAMES):
        assert np.array_equal(rref[:, k], rgot[:, k]), (nm, int((rref[:, k] != rgot[:, k]).sum()))


def test_mesh_adjoint_rows_vs_oracle_at_c2m_size(oracle32):
    """C2m (512x512 height-field mesh, 521 k triangles, 64 spp, max_depth 4, 256x128 envmap, gaussian film): the forward image and
    the adjoint of a 2-row sample of the image against the oracle.  In mesh mode a path scatters material gradients to whatever
    texels its secondary vertices hit, so the WHOLE gradient maps of that row sample are compared (radiance 1e-4, gradients 1e-3)."""
    import materialist_b200 as mb
    from materialist_b200 import renderop as mbr
    from oracle import oracle as orc
    from test_gpu_mesh_parity import _scene, _cuda_scene
    from test_reference_render_pin import REF_FLAGS, pin_cfg
    H = W = 512; spp = 64; row0, rows = 255, 2
    cam, verts, tris, a, r, m, env = _scene(H, W, env_hw=(128, 256))
    assert tris.shape[0] > 520_000
    om = oracle32.mesh_create(verts, tris)
    env_int, hier, d = oracle32.env_prepare(env, orc.ENV_ASSIGNED)
    seed = 17; sg = mb.default_seed_grad(seed)
    s = _cuda_scene(cam, verts, tris, env, REF_FLAGS)
    ta, tr, tm = (torch.from_numpy(x).cuda() for x in (a, r, m))
    ref = oracle32.mesh_render_fwd(pin_cfg(d, seed, row0, rows, spp=spp, H=H, W=W), om, a, r, m, None, env_int, hier, d)
    G = np.random.RandomState(5).randn(H, W, 3).astype(np.float32)
    gref = oracle32.mesh_render_bwd(pin_cfg(d, sg, row0, rows, spp=spp, H=H, W=W), om, a, r, m, None, env_int, hier, d, G, want=("a", "r", "m", "env"))
    gref["env"] = oracle32.env_grad_finish(gref.pop("env_int"), env.shape[1], orc.ENV_ASSIGNED)
    oracle32.mesh_destroy(om)
    with s.shard(row0, rows):
        img = mbr._forward(s, spp, seed, ta, tr, tm, None, s.prepared_env())
        e = rel_l2(img.cpu().numpy(), ref)
        assert e <= 1e-4, e
        Gh = torch.from_numpy(G[row0 - 2:row0 + rows + 2]).cuda().contiguous()
        g_a, g_r, g_m, _, g_env = mbr._backward(s, spp, sg, ta, tr, tm, None, s.prepared_env(), Gh, True, True, True, False, True)
    for key, got in (("a", g_a), ("r", g_r), ("m", g_m), ("env", g_env)):
        e = rel_l2(got.cpu().numpy().reshape(gref[key].shape), gref[key])
        assert e <= 1e-3, (key, e)
    assert float(g_a[:row0 - 8].abs().max()) > 0          # gradients did reach texels far outside the sampled rows (secondary vertices)


def test_mesh_mode_at_1080p_scale():
    """Mesh mode on a 1920x1080 height-field mesh (4.1 M triangles, 12 BVH levels): GPU-built BVH == brute force (float64-free
    restatement of the same Moeller-Trumbore in numpy, smallest t wins) on random rays; a forward render is finite, deterministic
    and invariant under row sharding."""
    import materialist_b200 as mb
    from materialist_b200 import synthetic
    from materialist_b200.mesh import Mesh
    H, W = 1080, 1920
    cam = mb.Camera(width=W, height=H)
    verts, tris = synthetic.grid_mesh(synthetic.bumpy_positions(H, W, cam))
    assert tris.shape[0] > 4_000_000
    gm = Mesh(verts, tris)
    assert gm.desc.n_levels >= 11
    rs = np.random.RandomState(0)
    n = 48
    tgt = verts[rs.randint(0, len(verts), n)] + rs.randn(n, 3).astype(np.float32) * 0.05
    o = np.zeros((n, 3), np.float32); o[n // 2:] = verts[rs.randint(0, len(verts), n - n // 2)] + np.float32([0, 0, 3.0])
    d = tgt - o; d /= np.linalg.norm(d, axis=-1, keepdims=True); d = d.astype(np.float32)
    tri_gpu, tuv = gm.intersect(o, d)
    tri_gpu = tri_gpu.cpu().numpy(); t_gpu = tuv.cpu().numpy()[:, 0]
    # brute force in float32 numpy, same operation order as the kernels (non-contracted: numpy never fuses)
    p0, p1, p2 = (verts[tris[:, k]] for k in range(3))
    e1, e2 = p1 - p0, p2 - p0
    for i in range(n):
        pvec = np.cross(d[i], e2).astype(np.float32)
        det = (e1 * pvec).sum(-1, dtype=np.float32)
        with np.errstate(all="ignore"):
            inv = np.float32(1) / det
            tvec = o[i] - p0
            u = (tvec * pvec).sum(-1, dtype=np.float32) * inv
            qvec = np.cross(tvec, e1).astype(np.float32)
            v = (d[i] * qvec).sum(-1, dtype=np.float32) * inv
            t = (e2 * qvec).sum(-1, dtype=np.float32) * inv
        ok = (u >= 0) & (u <= 1) & (v >= 0) & (u + v <= 1) & (t >= 0)
        if not ok.any():
            assert tri_gpu[i] < 0, i
            continue
        tmin = t[ok].min()
        # numpy's row sums may associate differently from the kernel's ((x+y)+z): compare t to a few ulp, ids when unambiguous
        assert tri_gpu[i] >= 0 and abs(t_gpu[i] - tmin) <= 4e-6 * max(1.0, abs(tmin)), (i, t_gpu[i], tmin)
        near = np.flatnonzero(ok & (np.abs(t - tmin) <= 4e-6 * max(1.0, abs(tmin))))
        assert tri_gpu[i] in near, (i, tri_gpu[i], near[:4])
    assert (tri_gpu >= 0).mean() > 0.5
    a, r, m = (t.cuda() for t in synthetic.materials(H, W, seed_base=1))
    s = mb.Scene.from_mesh(verts, tris, cam, envmap=synthetic.envmap(32, 64, seed=4))
    with torch.no_grad():
        full = mb.render(s, spp=2, seed=3, albedo=a, roughness=r, metallic=m)
        again = mb.render(s, spp=2, seed=3, albedo=a, roughness=r, metallic=m)
        s.set_shard(500, 80)
        part = mb.render(s, spp=2, seed=3, albedo=a, roughness=r, metallic=m)
    assert torch.isfinite(full).all() and float(full.mean()) > 0
    assert torch.equal(full, again) and torch.equal(full[500:580], part)


def test_film_weight_kernels_agree_at_c2_size():
    """The film weights of the adjoint render have two kernels: one thread per pixel for images that fill the GPU that way (the whole
    C2 image), one warp per pixel below that (a 16-row shard of it).  Same samples, same taps, different summation order: the rows
    both produce must agree to float rounding — and the full-image buffer must be the sum the weights are defined as (every sample
    spreads a total weight of (sum_i wx_i) (sum_j wy_j) over its 25 taps; spot-checked against a direct evaluation on the host)."""
    import ctypes as C
    from materialist_b200 import _abi, renderop
    c = Case(H=512, W=512, spp=64, He=16, We=32)
    s = c.scene()
    res_x = s.prepared_env()[2].res_x
    seed_grad = 12345
    full = renderop._film_weights(s, c.spp, seed_grad, res_x).clone()
    first = C.c_int(0)
    row0, rows = 200, 16
    with s.shard(row0, rows):
        cfg = s.make_cfg(c.spp, seed_grad, res_x)
        wrows = _abi.lib.mb200_bwd_wpart_rows(C.byref(cfg), C.byref(first))
        part = renderop._film_weights(s, c.spp, seed_grad, res_x).clone()
    assert part.shape[0] == wrows and wrows * c.W < 148 * 2 * 256          # the shard takes the warp-per-pixel kernel
    ref = full[first.value:first.value + wrows]
    assert float((part - ref).abs().max()) <= 2e-6 * float(ref.abs().max())
    assert float(ref.sum()) > 0
    # total weight per pixel against the analytic value spp * (integral of the filter)^2 within Monte-Carlo noise
    tot = full.sum(-1)
    g = lambda x: np.maximum(0.0, np.exp(-2.0 * x * x) - np.exp(-8.0))
    xs = (np.arange(200000) + 0.5) / 200000 * 4.0 - 2.0
    integral = float(g(xs).mean() * 4.0)
    assert abs(float(tot.mean()) / (c.spp * integral * integral) - 1.0) < 2e-3
