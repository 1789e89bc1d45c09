"""Mesh-mode oracle (oracle/mb_oracle_mesh.c): BVH against brute force, adjoint against float64 finite differences."""
import numpy as np
import pytest

from oracle import oracle as orc
from materialist_b200 import synthetic
from materialist_b200.camera import Camera
from test_reference_render_pin import pin_cfg, rel_l2


def small_scene(H=20, W=20, seed=0):
    cam = Camera(width=W, height=H)
    pos = synthetic.bumpy_positions(H, W, cam)
    verts, tris = synthetic.grid_mesh(pos)
    a, r, m = (x.numpy() for x in synthetic.materials(H, W, seed_base=11))
    env = synthetic.envmap(8, 16, sun=50.0).numpy()
    return cam, verts, tris, a, r, m, env


def test_grid_mesh_faces_camera():
    cam, verts, tris, *_ = small_scene()
    n = np.cross(verts[tris[:, 1]] - verts[tris[:, 0]], verts[tris[:, 2]] - verts[tris[:, 0]])
    c = verts[tris].mean(1)
    assert ((n * (cam.to_world[:3, 3] - c)).sum(-1) > 0).all()


def test_bvh_matches_brute_force(oracle32):
    cam, verts, tris, *_ = small_scene(24, 24)
    mesh = oracle32.mesh_create(verts, tris)
    rng = np.random.RandomState(1)
    n = 4000
    o = np.concatenate([np.zeros((n // 2, 3)), verts[rng.randint(0, len(verts), n - n // 2)] + rng.randn(n - n // 2, 3) * 0.5]).astype(np.float32)
    tgt = verts[rng.randint(0, len(verts), n)] + rng.randn(n, 3).astype(np.float32) * 0.3
    d = tgt - o; d /= np.linalg.norm(d, axis=-1, keepdims=True)
    t0, tuv0 = oracle32.mesh_intersect(mesh, o, d, brute=True)
    t1, tuv1 = oracle32.mesh_intersect(mesh, o, d)
    assert (t0 >= 0).mean() > 0.5
    assert (t0 == t1).all() and (tuv0 == tuv1).all()
    a0, _ = oracle32.mesh_intersect(mesh, o, d, brute=True, any_hit=True)
    a1, _ = oracle32.mesh_intersect(mesh, o, d, any_hit=True)
    assert (a0 == a1).all() and ((a0 > 0) == (t0 >= 0)).all()
    oracle32.mesh_destroy(mesh)


def test_mesh_paths_have_occlusion_and_interreflection(oracle32):
    cam, verts, tris, a, r, m, env = small_scene()
    mesh = oracle32.mesh_create(verts, tris)
    env_int, hier, d = oracle32.env_prepare(env, orc.ENV_ASSIGNED)
    img, st = oracle32.mesh_render_fwd(pin_cfg(d, 3, 0, 20, spp=16, H=20, W=20), mesh, a, r, m, None, env_int, hier, d, want_stats=True)
    assert np.isfinite(img).all() and img.mean() > 0
    paths, verts_n, occluded, escaped, _ = st.tolist()
    assert verts_n > 1.05 * paths * 0.9 and occluded > 0.02 * paths and escaped < paths
    img2 = oracle32.mesh_render_fwd(pin_cfg(d, 3, 0, 20, spp=16, H=20, W=20, max_depth=2), mesh, a, r, m, None, env_int, hier, d)
    assert img.mean() > img2.mean()          # interreflection adds energy
    oracle32.mesh_destroy(mesh)


@pytest.mark.parametrize("max_depth", [2, 4])
def test_mesh_adjoint_matches_finite_differences(oracle64, max_depth):
    """d<G, img>/d(a, m, env) of the seed_grad render: albedo and metallic do not move any sampled direction and the
    image is linear in the envmap texels for a frozen hierarchy, so central differences of the float64 oracle are exact
    up to O(eps^2).  (Roughness moves the sampled directions, which Mitsuba detaches; its chain rule is the same code.)
    Run with the bs.wo quirk OFF: with it, ~2 % of the vertices send the ray below the surface where p2 == 0 and the AD
    pass keeps the primal weight f/(p+1e-6), which this build treats as a constant (same convention as
    the G-buffer path and tests/torch_mirror.py) while a finite difference sees it."""
    H = W = 20
    cam, verts, tris, a, r, m, env = small_scene(H, W)
    O = oracle64
    mesh = O.mesh_create(verts, tris)
    env_int, hier, d = O.env_prepare(env, orc.ENV_ASSIGNED)
    cfg = pin_cfg(d, 77, 0, H, spp=8, H=H, W=W, max_depth=max_depth,
                  flags=orc.FLAG_ROW_STRIDE_H | orc.FLAG_ENV_HALF_TEXEL | orc.FLAG_AD_WEIGHTS)
    rng = np.random.RandomState(5)
    G = rng.randn(H, W, 3).astype(np.float32)
    g = O.mesh_render_bwd(cfg, mesh, a, r, m, None, env_int, hier, d, G, want=("a", "r", "m", "env"))

    def f(a_, m_, e_):
        return float((O.mesh_render_fwd(cfg, mesh, a_, r, m_, None, e_, hier, d).astype(np.float64) * G).sum())

    for name, grad in (("a", g["a"]), ("m", g["m"]), ("env", g["env_int"])):
        dlt = rng.randn(*grad.shape).astype(np.float32)
        eps = 0.25 if name == "env" else 1e-3        # linear in env: any step is exact; the float32 image output limits small steps
        args = {"a": (a, m, env_int), "m": (a, m, env_int), "env": (a, m, env_int)}[name]
        ap, mp, ep = args; am, mm, em = args
        if name == "a": ap, am = a + eps * dlt, a - eps * dlt
        if name == "m": mp, mm = m + eps * dlt, m - eps * dlt
        if name == "env": ep, em = env_int + eps * dlt, env_int - eps * dlt
        fd = (f(ap, mp, ep) - f(am, mm, em)) / (2 * eps)
        an = float((grad.astype(np.float64) * dlt).sum())
        assert abs(fd - an) <= 2e-3 * max(abs(fd), abs(an)) + 1e-6, (name, fd, an)
    assert np.abs(g["r"]).sum() > 0
    O.mesh_destroy(mesh)


def test_bvh_matches_brute_force_through_shared_vertices(oracle32):
    """Pixel-centre rays of the shipped indoor mesh pass (to rounding) THROUGH mesh vertices: Moeller-Trumbore accepts several
    triangles with u or v rounded to exactly 0 and the closest-hit rule (smallest t, then smallest id) must not depend on the
    acceleration structure.  (Unpadded boxes failed this for 3 % of the pixel centres.)"""
    import os
    fix = os.path.join(os.path.dirname(__file__), "golden", "indoor_pin.npz")
    g = np.load(fix)
    mesh = oracle32.mesh_create(g["verts"], g["tris"], face_normals=True)
    cam = Camera(width=512, height=512)
    rng = np.random.RandomState(0)
    px = rng.randint(0, 512, 1500) + 0.5; py = rng.randint(0, 512, 1500) + 0.5
    d = cam.pixel_ray_dirs(px, py).astype(np.float32); o = np.zeros_like(d)
    t0, tuv0 = oracle32.mesh_intersect(mesh, o, d, brute=True)
    t1, tuv1 = oracle32.mesh_intersect(mesh, o, d)
    assert (t0 == t1).all() and (tuv0 == tuv1).all()
    assert ((tuv0[:, 1] == 0) | (tuv0[:, 2] == 0)).mean() > 0.5      # the stress case is really exercised
    oracle32.mesh_destroy(mesh)
