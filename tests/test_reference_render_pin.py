"""PIN of the restated render operator (P-rows) against the REFERENCE'S OWN OUTPUT.

tests/golden/indoor_pin.npz (made by tests/golden/make_pin_fixture.py) holds the exact inputs of the render the reference
saved for its shipped scene output_imgs/indoor — mesh, material maps, envmap — and 32 rows of that render
(best_results/rendered_img.exr: Mitsuba 3.5 cuda_ad_rgb, `path` max_depth 4, 64 spp, linear radiance).  The sampler seed
of that run (993) was recovered by tools/ref_render_pin.py; with it the mesh-mode oracle reproduces the reference's
per-pixel Monte Carlo NOISE PATTERN, which only happens if the sampler seeding, the draw order of the path loop, the
primary rays, the mesh hits, the envmap hierarchy / emitter sampling, the BSDF sampling with the `bs.wo` world-space quirk
through the interpolated shading frame, MIS and the gaussian film are all restated correctly.  Each convention is also
checked to be the better one of its alternatives (switching it off makes the match measurably worse).
"""
import os

import numpy as np
import pytest

from oracle import oracle as orc
from materialist_b200.camera import Camera

FIX = os.path.join(os.path.dirname(__file__), "golden", "indoor_pin.npz")
REF_FLAGS = orc.FLAG_WO_WORLD_QUIRK | orc.FLAG_ROW_STRIDE_H | orc.FLAG_ENV_HALF_TEXEL


def pin_cfg(d, seed, row0, rows, flags=REF_FLAGS, max_depth=4, spp=64, H=512, W=512):
    cam = Camera(width=W, height=H)
    c = orc.Cfg()
    c.H, c.W, c.spp, c.max_depth = H, W, spp, max_depth
    c.seed = seed
    c.filter = orc.FILTER_GAUSSIAN
    c.flags = flags
    c.use_mesh_normal = 1
    c.row0, c.rows = row0, rows
    c.view[:] = cam.view_matrix.reshape(-1).tolist()
    c.proj[:] = cam.proj_matrix.reshape(-1).tolist()
    c.cam_to_world[:] = cam.to_world.astype(np.float32).reshape(-1).tolist()
    c.tan_half_fov_x = cam.tan_half_fov_x
    c.env_u_shift = float(np.float32(0.5) / np.float32(d.res_x - 1)) if flags & orc.FLAG_ENV_HALF_TEXEL else 0.0
    return c


def rel_l2(x, y):
    return float(np.linalg.norm(np.asarray(x, np.float64) - y) / np.linalg.norm(np.asarray(y, np.float64)))


@pytest.fixture(scope="module")
def pin():
    g = dict(np.load(FIX))
    O = orc.Oracle()
    g["O"] = O
    g["mesh"] = O.mesh_create(g["verts"], g["tris"])
    g["mesh_face"] = O.mesh_create(g["verts"], g["tris"], face_normals=True)
    g["envp"] = O.env_prepare(g["env"], orc.ENV_ASSIGNED)
    yield g
    O.mesh_destroy(g["mesh"]); O.mesh_destroy(g["mesh_face"])


def render_rows(g, seed, r0, rows, mesh="mesh", **kw):
    env_int, hier, d = g["envp"]
    return g["O"].mesh_render_fwd(pin_cfg(d, seed, r0, rows, **kw), g[mesh], g["a"], g["r"], g["m"], None, env_int, hier, d)


def test_oracle_reproduces_reference_render_noise_pattern(pin):
    row0 = int(pin["row0"]); seed = int(pin["seed"])
    r0, rows = row0 + 8, 16
    ref = pin["ref"][8:8 + rows]
    img = render_rows(pin, seed, r0, rows)
    e_all = rel_l2(img, ref); e_g = rel_l2(img[..., 1], ref[..., 1])
    # absolute radiance, NO fitted scale: 64-spp Monte Carlo noise alone is 6-7 % (next assert)
    assert e_g < 0.008, e_g          # green: the channel without the albedo last-vs-best-iterate mismatch (see make_pin_fixture.py)
    assert e_all < 0.02, e_all
    assert abs(img.mean() / ref.mean() - 1) < 0.01
    wrong = render_rows(pin, seed - 1, r0, rows)
    assert rel_l2(wrong, ref) > 0.05 and rel_l2(wrong[..., 1], ref[..., 1]) > 0.05


def test_pin_discriminates_conventions(pin):
    """Every restated convention beats its alternative on the reference's own image (green channel, rows 248..264)."""
    row0 = int(pin["row0"]); seed = int(pin["seed"])
    r0, rows = row0 + 8, 16
    ref = pin["ref"][8:8 + rows][..., 1]
    base = rel_l2(render_rows(pin, seed, r0, rows)[..., 1], ref)
    alt = {
        "bs.wo world-space quirk OFF (mi_plugin.py:1444)": render_rows(pin, seed, r0, rows, flags=REF_FLAGS & ~orc.FLAG_WO_WORLD_QUIRK),
        "envmap half-texel u shift OFF": render_rows(pin, seed, r0, rows, flags=REF_FLAGS & ~orc.FLAG_ENV_HALF_TEXEL),
        "max_depth 3": render_rows(pin, seed, r0, rows, max_depth=3),
        "max_depth 2 (direct lighting only)": render_rows(pin, seed, r0, rows, max_depth=2),
        "face normals as shading frame (no computed vertex normals)": render_rows(pin, seed, r0, rows, mesh="mesh_face"),
    }
    for name, img in alt.items():
        e = rel_l2(img[..., 1], ref)
        assert e > 1.3 * base, (name, e, base)


# ---------------------------------------------------------------- second shipped scene: output_imgs/jinjya (seed 705)
FIX2 = os.path.join(os.path.dirname(__file__), "golden", "jinjya_pin.npz")


def _srgb_scaled_err(img, ref):
    """rel-L2 on x^(1/2.2) with one fitted scale: the saved image is linear_to_srgb(render * gt.mean()/pred.mean())."""
    sr = np.maximum(img, 0) ** (1 / 2.2)
    k = (sr * ref).sum() / (sr * sr).sum()
    return rel_l2(sr * k, ref)


def test_oracle_reproduces_second_reference_render_jinjya():
    """The same oracle, unchanged, singles out ONE seed of the 1000 on the reference's second shipped scene as well
    (tests/golden/pin_search_jinjya.txt: 705 at 0.011 against >= 0.036 for all others): an independent confirmation of the seeding,
    draw order, mesh hits, emitter sampling, bs.wo quirk, MIS and film restated for the P-rows — on a render that was saved by the
    BRDF phase (render_w_brdf), i.e. through the other of the two operator entry points."""
    g = dict(np.load(FIX2))
    O = orc.Oracle()
    mesh = O.mesh_create(g["verts"], g["tris"])
    try:
        env_int, hier, d = O.env_prepare(g["env"], orc.ENV_ASSIGNED)
        row0, seed = int(g["row0"]), int(g["seed"])
        r0, rows = row0 + 6, 12
        ref = g["ref"][6:6 + rows]
        render = lambda sd: O.mesh_render_fwd(pin_cfg(d, sd, r0, rows), mesh, g["a"], g["r"], g["m"], None, env_int, hier, d)
        e = _srgb_scaled_err(render(seed), ref)
        assert e < 0.02, e           # 0.011 - 0.017 depending on the rows (any other seed: 0.036 - 0.039, pure Monte Carlo noise)
        for wrong in (seed - 1, seed + 1, 993):
            ew = _srgb_scaled_err(render(wrong), ref)
            assert ew > 0.03 and ew > 2.0 * e, (wrong, ew, e)
    finally:
        O.mesh_destroy(mesh)
