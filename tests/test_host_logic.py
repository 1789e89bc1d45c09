"""CPU: host-side logic that needs no GPU."""
import torch


def test_brdf_phase_lr_matches_guarded_steplr():
    from materialist_b200.inverse import brdf_phase_lr
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.Adam([p], lr=3e-4)
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=100, gamma=0.8)
    for k in range(650):
        lr = opt.param_groups[0]["lr"]
        assert abs(lr - brdf_phase_lr(k)) < 1e-12, k
        opt.step()
        if lr > 1.5e-4:
            sched.step()
    assert abs(brdf_phase_lr(10_000) - 3e-4 * 0.8 ** 4) < 1e-12


# ---------------------------------------------------------------- small helpers of envmap_utils / computeSH vs the reference
def _golden(name):
    import os
    import numpy as np
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name))


def test_uv_to_envmap_and_cdf_helpers_match_reference():
    """uvToEnvmap (computeSH.py:75-85) and the CDF helpers (envmap_utils.py:92-136) against vectors produced by the
    reference's own functions (tests/golden/make_golden.py: helpers_and_rotate); pure host code, no GPU."""
    import numpy as np
    from materialist_b200.myutils import computeSH as sh, envmap_utils as eu
    from oracle import aux_oracle as aux
    g = _golden("helpers_rotate.npz")
    col = sh.uvToEnvmap(g["env"], g["uv"][0], g["uv"][1])
    assert np.allclose(col, g["uv_color"], rtol=1e-6, atol=1e-7)
    d = aux.build_envmap(g["cdf_env"])
    c_cdf, m_cdf = torch.from_numpy(d["c_cdf"]), torch.from_numpy(d["m_cdf"])
    x = torch.from_numpy(g["cdf_x"])
    vi = eu.cdf_search_1d(m_cdf, x)
    assert np.array_equal(vi.numpy(), g["cdf_vi"])
    assert np.allclose(eu.get_pdf_from_cdf_1d(m_cdf, vi.clamp_max(15)).numpy(), g["pdf1d"], rtol=1e-5, atol=1e-7)
    assert np.allclose(eu.interp_1d(m_cdf, x, vi.clamp_max(15)).numpy(), g["interp1d"], rtol=1e-4, atol=1e-5)
    ui = eu.cdf_search_2d(c_cdf, x, 5)
    assert np.array_equal(ui.numpy(), g["cdf_ui"])
    assert np.allclose(eu.get_pdf_from_cdf_2d(c_cdf, ui.clone(), 5).numpy(), g["pdf2d"], rtol=1e-5, atol=1e-7)
    ref = g["interp2d"]
    got_nan = eu.interp_2d(c_cdf, x, ui.clone(), 5, ref_exact_nan=True).numpy()
    assert np.array_equal(np.isnan(got_nan), np.isnan(ref))                      # the reference's NaN at index 0
    ok = ~np.isnan(ref)
    assert np.allclose(got_nan[ok], ref[ok], rtol=1e-4, atol=1e-5)
    assert np.isfinite(eu.interp_2d(c_cdf, x, ui.clone(), 5).numpy()).all()      # default: the valid branch is selected
