"""Drop-in for the reference's myutils/envmap_utils.py (marginal / conditional CDF envmap sampling in torch),
running on the CUDA kernels mb200_cdf_build / mb200_cdf_sample.  Same names, argument meaning and quirks:

  * build_envmap    :43-66   marginal built from the sum of the CUMULATIVE row (:53-55), Rec.601 luma, +1e-6
  * sample_envmap   :172-201 `u = u_idx` (no fractional du), z-up angle2xyz, pdf with +1e-6
  * importance_sample :139-169 adds `du` and a nearest-texel lookup; the reference's version yields NaN whenever
                      u_idx == 0 (interp_2d divides 0/0, :105-107); here those lanes use the `interp_else` branch,
                      which is what the masked blend intends.  Pass ref_exact_nan=True to get the NaNs back.
  * lookup_envmap   :29-36   nearest texel, y-up convention (inconsistent with sample_envmap's z-up — kept)
  * sample_env1 / sample_brdf1 :7-28  thin wrappers over the torch BRDF of mi_plugin
"""
import math

import torch

from .. import _abi


def lookup_envmap(envmap, w):
    height, width = envmap.shape[0], envmap.shape[1]
    phi = torch.atan2(w[..., 0], -w[..., 2]) / (2.0 * math.pi)
    u = torch.clamp((phi * width + width) % width, 0, width - 1).int()
    theta = torch.acos(w[..., 1]) / (math.pi)
    v = torch.clamp(theta * height, 0, height - 1).int()
    return envmap[v.long(), u.long()]


def luminance(x):
    return 0.299 * x[0] + 0.587 * x[1] + 0.114 * x[2]


def build_envmap(envmap):
    if envmap.ndim != 3 or envmap.shape[-1] != 3:
        raise ValueError("envmap must be (h, w, 3)")
    env = envmap.detach().contiguous().float()
    h, w, _ = env.shape
    c_cdf = torch.empty(h, w, device=env.device); m_cdf = torch.empty(h, device=env.device)
    _abi.check(_abi.lib.mb200_cdf_build(_abi.ptr(env), h, w, _abi.ptr(c_cdf), _abi.ptr(m_cdf), _abi.stream_ptr()), "mb200_cdf_build")
    return {"envmap": envmap, "c_cdf": c_cdf, "m_cdf": m_cdf}


def _sample(envmap_dict, sample2):
    c_cdf, m_cdf = envmap_dict["c_cdf"].contiguous(), envmap_dict["m_cdf"].contiguous()
    h, w = c_cdf.shape
    if sample2.ndim != 2 or sample2.shape[0] != 2:
        raise ValueError("sample2 must be (2, n)")
    s = sample2.detach().contiguous().float()
    n = s.shape[1]
    dirs = torch.empty(n, 3, device=s.device); pdf = torch.empty(n, 1, device=s.device)
    v_idx = torch.empty(n, 1, dtype=torch.int64, device=s.device); u_idx = torch.empty(n, 1, dtype=torch.int64, device=s.device)
    _abi.check(_abi.lib.mb200_cdf_sample(_abi.ptr(c_cdf), _abi.ptr(m_cdf), h, w, _abi.ptr(s), n, _abi.ptr(dirs), _abi.ptr(pdf),
                                         _abi.ptr(v_idx), _abi.ptr(u_idx), _abi.stream_ptr()), "mb200_cdf_sample")
    return dirs, pdf, v_idx, u_idx


def sample_envmap(envmap_dict, sample2):
    dirs, pdf, _, _ = _sample(envmap_dict, sample2)
    return dirs, pdf


def sample_envmap_indices(envmap_dict, sample2):
    """(v_idx, u_idx) int64 — the searchsorted indices behind sample_envmap (the in-repo 'CDF indices')."""
    return _sample(envmap_dict, sample2)[2:]


def compute_direction(theta, phi):
    from .mi_plugin import angle2xyz
    return angle2xyz(theta, phi)


def importance_sample(envmap_dict, sample2, ref_exact_nan=False):
    envmap, marg_cdf, cond_cdf = envmap_dict["envmap"], envmap_dict["m_cdf"], envmap_dict["c_cdf"]
    h, w, _ = envmap.shape
    _, _, v_idx, u_idx = _sample(envmap_dict, sample2)
    x0 = sample2[0, :].reshape(-1, 1); x1 = sample2[1, :].reshape(-1, 1)
    vi = v_idx.clamp(max=h - 1)
    prev = torch.where(v_idx > 0, marg_cdf[(vi - 1).clamp(min=0)], torch.zeros_like(x0))
    dv = torch.where(v_idx > 0, (x0 - prev) / (marg_cdf[vi] - prev), x0 / marg_cdf[vi])
    pdf_m = torch.where(v_idx > 0, marg_cdf[vi] - prev, marg_cdf[vi])
    v = v_idx + dv
    ui = u_idx.clone(); ui1 = (ui - 1).clamp(min=0); ui[ui == 32] = 31; ui = ui.clamp(max=w - 1)
    row = vi.reshape(-1)
    c_hi, c_lo = cond_cdf[row, ui.reshape(-1)].reshape(-1, 1), cond_cdf[row, ui1.reshape(-1)].reshape(-1, 1)
    interp_if = (x1 - c_lo) / (c_hi - c_lo)
    interp_else = x1 / c_hi
    mask = (u_idx > 0).float()
    du = mask * interp_if + (1 - mask) * interp_else if ref_exact_nan else torch.where(u_idx > 0, interp_if, interp_else)
    pdf_c = torch.where(u_idx > 0, c_hi - c_lo, c_hi)
    u = u_idx + du
    theta = v * math.pi / h
    phi = (2.0 * u * math.pi) / w
    dirs = compute_direction(theta.flatten(), phi.flatten()).float()
    emission = lookup_envmap(envmap, dirs)
    pdf = (h * w) * (pdf_c * pdf_m) / (2.0 * math.pi * math.pi * torch.sin(theta))
    return dirs, pdf, emission


def sample_env1(wo, normals, mat, use_mesh_normals, device, envmap_t):
    from .mi_plugin import eval_brdf
    sample2 = torch.rand(2, len(normals), device=device)
    wi, pdf = sample_envmap(envmap_t, sample2)
    brdf, pdf_brdf = eval_brdf(wi, wo, normals, mat, use_mesh_normals)
    brdf_weight = torch.nan_to_num(brdf / (pdf + 1e-6), nan=0, posinf=0, neginf=0)
    return wi, pdf, brdf_weight


def sample_brdf1(wo, normals, mat, use_mesh_normals, device):
    from .mi_plugin import sample_brdf
    sample1 = torch.rand(len(normals), device=device)
    sample2 = torch.rand(len(normals), 2, device=device)
    return sample_brdf(sample1, sample2, wo, normals, mat, use_mesh_normals)


# ------------------------------------------------------------------ small CDF helpers of the reference, same semantics
def cdf_search_1d(cdf, x):
    """:131-132"""
    return torch.searchsorted(cdf, x)


def cdf_search_2d(cdf, x, row_offset):
    """:135-136"""
    return torch.searchsorted(cdf[row_offset, :], x)


def get_pdf_from_cdf_1d(cdf, idx):
    """:109-113 — cdf[idx] - cdf[idx-1] (cdf[idx] at idx 0)."""
    mask = (idx > 0).float()
    return mask * (cdf[idx] - cdf[(idx - 1).clamp_min(0)]) + (1 - mask) * cdf[idx]


def get_pdf_from_cdf_2d(cdf, idx, row_offset=0):
    """:115-124 — per-row variant; like the reference, indices equal to the row length are clamped to the last column
    (the reference hard-codes 32 / 31, :120)."""
    mask = (idx > 0).float()
    idx = idx.clamp_max(cdf.shape[1] - 1)
    idx1 = (idx - 1).clamp_min(0)
    return mask * (cdf[row_offset, idx] - cdf[row_offset, idx1]) + (1 - mask) * cdf[row_offset, idx]


def interp_1d(buf, x, index):
    """:92-97 — position of x inside CDF bin `index`, in [0, 1]."""
    mask = (index > 0).float()
    lo = buf[(index - 1).clamp_min(0)]
    interp_if = (x - lo) / (buf[index] - lo)
    interp_else = x / buf[index]
    return torch.where(mask > 0, interp_if, interp_else)


def interp_2d(buf, x, index, row=0, ref_exact_nan=False):
    """:99-107.  The reference evaluates BOTH branches and blends them with a 0/1 mask, so index 0 yields 0 * (x - c) / (c - c)
    = NaN; `ref_exact_nan=True` reproduces that, the default selects the valid branch."""
    mask = (index > 0).float()
    index = index.clamp_max(buf.shape[1] - 1)
    index1 = (index - 1).clamp_min(0)
    interp_if = (x - buf[row, index1]) / (buf[row, index] - buf[row, index1])
    interp_else = x / buf[row, index]
    if ref_exact_nan:
        return mask * interp_if + (1 - mask) * interp_else
    return torch.where(mask > 0, interp_if, interp_else)


def build_envmap_np(envmap):
    """:68-90 — numpy-input variant of build_envmap.  NOTE it differs from build_envmap exactly as in the reference:
    Rec.601 luminance via `luminance`, NO +1e-6 in the normalisation, marginal from the sum of the CUMULATIVE rows."""
    envmap = np.asarray(envmap)
    h, w, _ = envmap.shape
    h01 = (np.arange(h) + 0.5) / h
    lum = 0.299 * envmap[..., 0] + 0.587 * envmap[..., 1] + 0.114 * envmap[..., 2]     # np.apply_along_axis(luminance, 2, envmap)
    lum_sin = lum * np.sin(np.pi * h01).reshape(h, 1)
    c_cdf = np.cumsum(lum_sin, axis=1)
    m_cdf = np.cumsum(np.sum(c_cdf, axis=1))
    c_cdf = c_cdf / c_cdf[:, -1].reshape(h, 1)
    m_cdf = m_cdf / m_cdf[-1]
    if not torch.cuda.is_available():
        raise RuntimeError("build_envmap_np places its tables on the GPU like the reference (.cuda()); no CUDA device")
    return {"envmap": torch.from_numpy(envmap).cuda(), "c_cdf": torch.from_numpy(c_cdf).cuda(), "m_cdf": torch.from_numpy(m_cdf).cuda()}
