"""Drop-in for the reference's myutils/computeSH.py (order-4 real spherical harmonics of lat-long envmaps; the
reference copies lzqsd/SingleImageShapeAndSVBRDF computeSH.py).  numpy in / numpy out like the reference, fp64, the
per-texel Python loops replaced by the CUDA kernels mb200_sh_project / mb200_sh_reconstruct.

  computeK :58-68                 (np.math.factorial fixed: numpy >= 2 has no np.math)
  computeSHFromImage :299-347     per-texel jittered sample, MC integration of the 25 basis functions -> coef (25, 3).
                                  Jitter comes from the global numpy RNG exactly like the reference (seed numpy for
                                  reproducibility); pass `jitter=(h*w, 2)` to supply the (y, x) uniforms explicitly.
  reconstImageFromSH :226-240
  compute_sh_coeff_torch :410-430, reconstruct_envmap_from_sh :480-494, torch_sph_harm :468-478 — broken as shipped
                                  (computeK(int, int) raises, legendre_polynomial takes a product over the whole tensor);
                                  implemented here with the INTENDED maths: Riemann-sum projection 4pi/(W H) sum L Y sin(theta).
  compute_sh_coefficients :395-409, compute_sh_coeff_minh :432-456 — NameError in the reference (sph_harm / lpmv never
                                  imported); they raise the same NameError here.
"""
import math

import numpy as np
import torch

from .. import _abi

LARR = np.array([0, 1, 1, 1, 2, 2, 2, 2, 2, 3, 3, 3, 3, 3, 3, 3, 4, 4, 4, 4, 4, 4, 4, 4, 4], dtype=np.int32)
MARR = np.array([0, -1, 0, 1, -2, -1, 0, 1, 2, -3, -2, -1, 0, 1, 2, 3, -4, -3, -2, -1, 0, 1, 2, 3, 4], dtype=np.int32)


def computeK(l, m):
    l = np.atleast_1d(np.asarray(l)); m = np.absolute(np.atleast_1d(np.asarray(m)))
    l_s_m = np.array([math.factorial(int(x)) for x in (l - m)]).astype(np.float32)
    l_a_m = np.array([math.factorial(int(x)) for x in (l + m)]).astype(np.float32)
    return np.sqrt((2 * l + 1) * l_s_m / l_a_m / 4 / np.pi)


def angleToUV(theta, phi):
    return (phi + np.pi) / 2 / np.pi, 1 - theta / np.pi


def _dev():
    if not torch.cuda.is_available():
        raise RuntimeError("materialist_b200.computeSH needs a CUDA device (no CPU fallback)")
    return torch.device("cuda")


def computeSHFromImage(im, jitter=None):
    im = np.asarray(im, dtype=np.float64)
    h, w = im.shape[0], im.shape[1]
    if jitter is None:                                  # same draw order as the reference loop: y then x, row-major
        jitter = np.random.random(h * w * 2).reshape(h * w, 2)
    r, c = np.divmod(np.arange(h * w), w)
    y = (r + jitter[:, 0]) / float(h); x = (c + jitter[:, 1]) / float(w)
    angles = np.stack([2 * np.arccos(np.sqrt(1 - y)), 2 * np.pi * x - np.pi], 1)
    dev = _dev()
    t_im = torch.from_numpy(np.ascontiguousarray(im[..., :3])).to(dev); t_ang = torch.from_numpy(angles).to(dev)
    coef = torch.empty(25, 3, dtype=torch.float64, device=dev)
    _abi.check(_abi.lib.mb200_sh_project(_abi.ptr(t_im), h, w, _abi.ptr(t_ang), h * w, _abi.ptr(coef), _abi.stream_ptr()), "mb200_sh_project")
    return coef.cpu().numpy()


def reconstImageFromSH(coef, nrows, ncols, K=None, isClip=True):
    dev = _dev()
    t_coef = torch.from_numpy(np.ascontiguousarray(np.asarray(coef, dtype=np.float64))).to(dev)
    img = torch.empty(nrows, ncols, 3, dtype=torch.float64, device=dev)
    _abi.check(_abi.lib.mb200_sh_reconstruct(_abi.ptr(t_coef), nrows, ncols, int(bool(isClip)), _abi.ptr(img), _abi.stream_ptr()), "mb200_sh_reconstruct")
    return img.cpu().numpy()


# ------------------------------------------------------------------ S4: the "AfterRotate" variants (:242-297, :349-391)
def uvToEnvmap(envmap, u, v):
    """:75-85, vectorised over arrays of (u, v): bilinear fetch with c = u (W-1), r = (1-v)(H-1), int() truncation and the
    clamped +1 neighbour."""
    envmap = np.asarray(envmap)
    height, width = envmap.shape[0], envmap.shape[1]
    u = np.asarray(u, dtype=np.float64); v = np.asarray(v, dtype=np.float64)
    c, r = u * (width - 1), (1 - v) * (height - 1)
    cs, rs = c.astype(np.int64), r.astype(np.int64)                 # int(): truncation towards zero
    ce, re = np.minimum(width - 1, cs + 1), np.minimum(height - 1, rs + 1)
    wc, wr = (c - cs)[..., None], (r - rs)[..., None]
    color1 = (1 - wc) * envmap[rs, cs, :] + wc * envmap[rs, ce, :]
    color2 = (1 - wc) * envmap[re, cs, :] + wc * envmap[re, ce, :]
    return (1 - wr) * color1 + wr * color2


def _camera_frame(cameraLoc, cameraUp, isInv):
    cameraLoc = np.asarray(cameraLoc, dtype=np.float32); cameraUp = np.asarray(cameraUp, dtype=np.float32)
    cameraLoc = cameraLoc / np.sqrt(np.sum(cameraLoc * cameraLoc), dtype=np.float32)
    cameraUp = cameraUp / np.sqrt(np.sum(cameraUp * cameraUp), dtype=np.float32)
    rz, ry = cameraLoc, cameraUp
    rx = np.cross(ry, rz); rx = rx / np.sqrt(np.sum(rx * rx))
    ry = np.cross(rz, rx); ry = ry / np.sqrt(np.sum(ry * ry))
    if isInv:
        rot = np.stack([rx, ry, rz], axis=1).transpose([1, 0])
        rx, ry, rz = rot[:, 0], rot[:, 1], rot[:, 2]
    return rx, ry, rz


def _rotate_latlong(envmap, cameraLoc, cameraUp, isInv):
    """The per-texel loop shared by :272-295 and :366-388, vectorised: direction of texel (r, c) in the camera frame ->
    (theta, phi) in the world frame -> bilinear fetch.  float64 like the reference's Python scalars, float32 result."""
    envmap = np.asarray(envmap)
    rx, ry, rz = _camera_frame(cameraLoc, cameraUp, isInv)
    height, width = envmap.shape[0], envmap.shape[1]
    r, c = np.meshgrid(np.arange(height), np.arange(width), indexing="ij")
    theta = r / float(height - 1) * np.pi
    phi = c / float(width) * np.pi * 2 - np.pi
    x, y, z = np.sin(theta) * np.cos(phi), np.sin(theta) * np.sin(phi), np.cos(theta)
    coord = x[..., None] * rx + y[..., None] * ry + z[..., None] * rz
    nx, ny, nz = coord[..., 0], coord[..., 1], coord[..., 2]
    with np.errstate(invalid="ignore"):
        thetaNew = np.arccos(nz)
        den = np.sqrt(1 - nz * nz) + 1e-12
    nx = np.clip(nx / den, -1, 1); ny = np.clip(ny / den, -1, 1)
    phiNew = np.arccos(nx)
    phiNew = np.where(ny < 0, -phiNew, phiNew)
    u, v = angleToUV(thetaNew, phiNew)
    return uvToEnvmap(envmap, u, v).astype(np.float32)


def reconstImageFromSHAfterRotate(coef, cameraLoc, cameraUp, nrows=512, ncols=1024, K=None, isClip=True, isInv=False):
    """:242-297 — reconstruct the lat-long map from the 25 coefficients (CUDA), then resample it into the camera frame."""
    envmap = reconstImageFromSH(coef, nrows, ncols, K, isClip)
    return _rotate_latlong(envmap, cameraLoc, cameraUp, isInv)


def computeSHFromImageAfterRotate(envmap, cameraLoc, cameraUp, isInv=False, jitter=None):
    """:349-391 — resample the map into the camera frame, then project (CUDA; jitter as in computeSHFromImage)."""
    return computeSHFromImage(_rotate_latlong(envmap, cameraLoc, cameraUp, isInv), jitter=jitter)


# ------------------------------------------------------------------ intended behaviour of the broken torch variants
def _assoc_legendre(l, m, x):
    """P_l^m(x) with the Condon-Shortley phase, elementwise (the recurrence legendre_polynomial :458-466 intends)."""
    pmm = torch.ones_like(x)
    if m > 0:
        somx2 = torch.sqrt((1 - x) * (1 + x))
        fact = 1.0
        for _ in range(m):
            pmm = -pmm * fact * somx2; fact += 2.0
    if l == m:
        return pmm
    pmmp1 = x * (2 * m + 1) * pmm
    if l == m + 1:
        return pmmp1
    for ll in range(m + 2, l + 1):
        pll = ((2 * ll - 1) * x * pmmp1 - (ll + m - 1) * pmm) / (ll - m)
        pmm, pmmp1 = pmmp1, pll
    return pmmp1


def torch_sph_harm(m, l, phi, theta):
    P_lm = _assoc_legendre(l, abs(m), torch.cos(theta))
    K = float(computeK(l, m)[0])
    if m > 0:
        return math.sqrt(2) * torch.cos(m * phi) * P_lm * K
    if m < 0:
        return math.sqrt(2) * torch.sin(-m * phi) * P_lm * K
    return P_lm * K


def compute_sh_coeff_torch(hdr_image, l_max=2):
    device = hdr_image.device
    height, width, _ = hdr_image.shape
    coeffs = torch.zeros((l_max + 1, 2 * l_max + 1, 3), dtype=torch.float32, device=device)
    phis = torch.linspace(0, 2 * math.pi, width, device=device); thetas = torch.linspace(0, math.pi, height, device=device)
    phis, thetas = torch.meshgrid(phis, thetas, indexing="xy")
    sin_thetas = torch.sin(thetas)
    for l in range(l_max + 1):
        for m in range(-l, l + 1):
            Ylm_sin = torch_sph_harm(m, l, phis, thetas) * sin_thetas
            coeffs[l, m + l] += (hdr_image * Ylm_sin[..., None]).sum((0, 1))
    return coeffs * (4 * math.pi / (width * height))


def reconstruct_envmap_from_sh(sh_coeffs, width, height, l_max=2):
    device = sh_coeffs.device
    envmap = torch.zeros((height, width, 3), dtype=torch.float32, device=device)
    phis = torch.linspace(0, 2 * math.pi, width, device=device); thetas = torch.linspace(0, math.pi, height, device=device)
    phis, thetas = torch.meshgrid(phis, thetas, indexing="xy")
    for l in range(l_max + 1):
        for m in range(-l, l + 1):
            envmap += sh_coeffs[l, m + l][None, None, :] * torch_sph_harm(m, l, phis, thetas)[..., None]
    return envmap


def compute_sh_coefficients(hdr_map, num_samples):
    raise NameError("name 'sph_harm' is not defined")        # as in the reference (computeSH.py:395-409)


def compute_sh_coeff_minh(hdr_image, l_max=2):
    raise NameError("name 'lpmv' is not defined")            # as in the reference (computeSH.py:432-456)
