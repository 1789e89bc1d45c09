"""CPU ORACLE (numpy) for the helper rows of the hot path — TEST INFRASTRUCTURE ONLY (see oracle/oracle.py).

Restates, with the reference file:line each function follows:
  PosMLP            mymodels/mlps.py:8-54 (embedder), :129-251 (PosMLP), as instantiated at inverse_img_w_mi.py:117-124, :163
  build_envmap      myutils/envmap_utils.py:43-66
  sample_envmap     myutils/envmap_utils.py:172-201
  lookup_envmap     myutils/envmap_utils.py:29-36
  computeK / SH     myutils/computeSH.py:13-68, :87-162 (Integration), :165-224 (projection), :226-240, :299-347
Pinned against tests/golden/{posmlp,envmap_utils,compute_sh}.npz, which were produced by importing the reference
itself (tests/golden/make_golden.py).
"""
import math

import numpy as np


# ====================================================================================== PosMLP
class PosMLPOracle:
    """dims = [in] + [256]*4 + [out]; skip_connection = [1, 3]; SineLayer = sin(Wx + b); lin4 plain."""

    def __init__(self, weights, biases, n_color, n_out, output_type, n_freq=2, hidden=256):
        self.W = [np.asarray(w, np.float64) for w in weights]
        self.b = [np.asarray(b, np.float64) for b in biases]
        self.n_color, self.n_out, self.output_type, self.n_freq, self.hidden = n_color, n_out, output_type, n_freq, hidden

    @staticmethod
    def grid_shape(N):
        """img2points mlps.py:190-198: shape inferred from N."""
        if N > 512:
            h = int(round(N ** 0.5)); return h, h
        h = int(round((N / 2) ** 0.5)); return h, 2 * h

    def points(self, img, H=None, W=None):
        N = img.shape[0]
        if H is None:
            H, W = self.grid_shape(N)
        rows, cols = np.meshgrid(np.arange(H, dtype=np.float64), np.arange(W, dtype=np.float64), indexing="ij")
        p = np.stack([rows.reshape(-1), cols.reshape(-1)], 1)
        feats = [p]
        for k in range(self.n_freq):                    # freq_bands = 2 ** linspace(0, n_freq-1, n_freq)
            f = 2.0 ** k
            feats += [np.sin(p * f), np.cos(p * f)]
        return np.concatenate(feats + [np.asarray(img, np.float64)], 1)

    def forward(self, img, H=None, W=None):
        pts = self.points(img, H, W)
        x = pts
        self.cache = {"pts": pts, "x": [], "z": []}
        for l in range(5):
            if l in (1, 3):
                x = np.concatenate([x, pts], 1)
            self.cache["x"].append(x)
            z = x @ self.W[l].T + self.b[l]
            self.cache["z"].append(z)
            x = np.sin(z) if l < 4 else z
        img = np.asarray(img, np.float64)
        if self.output_type == "envmap":
            y = np.logaddexp(0.0, x)                     # softplus
        else:                                            # 'arm': 1.3*tanh(x) + img, straight-through clamp
            y = 1.3 * np.tanh(x) + img
            self.cache["y_pre"] = y
            y = np.clip(y, 0, 1)
        self.cache["out_pre"] = x
        return y

    def backward(self, gy):
        """Returns (gW list, gb list, g_img). STE clamp: gradient passes straight through (mlps.py:234)."""
        c = self.cache
        gy = np.asarray(gy, np.float64)
        if self.output_type == "envmap":
            g = gy / (1.0 + np.exp(-c["out_pre"]))
            g_img = np.zeros((gy.shape[0], self.n_color))
        else:
            g = gy * 1.3 * (1 - np.tanh(c["out_pre"]) ** 2)
            g_img = gy.copy()
        gW, gb = [None] * 5, [None] * 5
        n_pts = c["pts"].shape[1]
        g_pts = np.zeros_like(c["pts"])
        for l in range(4, -1, -1):
            if l < 4:
                g = g * np.cos(c["z"][l])
            gW[l] = g.T @ c["x"][l]; gb[l] = g.sum(0)
            gx = g @ self.W[l]
            if l in (1, 3):
                g_pts += gx[:, -n_pts:]; gx = gx[:, :-n_pts]
            g = gx
        g_pts += g
        g_img = g_img + g_pts[:, -self.n_color:]
        return gW, gb, g_img


# ====================================================================================== envmap_utils
def build_envmap(env):
    """envmap_utils.py:43-66. float32 in / out; cumsum accumulates in float64 and rounds each prefix to float32
    (torch CPU acc_type<float> = double); the row sum behind the marginal is taken in float32 pairwise order by
    torch and is only reproducible to rounding."""
    env = np.asarray(env, np.float32)
    h, w, _ = env.shape
    h01 = np.array([(v + 0.5) / h for v in range(h)], dtype=np.float32)
    lum = (np.float32(0.299) * env[:, :, 0] + np.float32(0.587) * env[:, :, 1]) + np.float32(0.114) * env[:, :, 2]
    sin_theta = np.sin((np.float32(np.pi) * h01).astype(np.float32)).astype(np.float32).reshape(h, 1)
    lum_sin = (lum * sin_theta).astype(np.float32)
    c_cdf = np.cumsum(lum_sin.astype(np.float64), axis=1).astype(np.float32)
    marg = c_cdf.astype(np.float64).sum(axis=1).astype(np.float32)              # sum of the CUMULATIVE row (:53)
    m_cdf = np.cumsum(marg.astype(np.float64)).astype(np.float32)
    c_cdf = (c_cdf / (c_cdf[:, -1].reshape(h, 1) + np.float32(1e-6))).astype(np.float32)
    m_cdf = (m_cdf / (m_cdf[-1] + np.float32(1e-6))).astype(np.float32)
    return {"envmap": env, "c_cdf": c_cdf, "m_cdf": m_cdf}


def angle2xyz(theta, phi):
    """mi_plugin.py:46-58 (z-up)."""
    st = np.sin(theta)
    v = np.stack([st * np.cos(phi), st * np.sin(phi), np.cos(theta)], -1)
    return v / np.maximum(np.linalg.norm(v, axis=-1, keepdims=True), 1e-12)


def sample_envmap(c_cdf, m_cdf, sample2):
    """envmap_utils.py:172-201. sample2 (2, n) float32 -> dirs (n,3), pdf (n,1), v_idx (n,1), u_idx (n,1) int64."""
    c_cdf = np.asarray(c_cdf, np.float32); m_cdf = np.asarray(m_cdf, np.float32); s = np.asarray(sample2, np.float32)
    h, w = c_cdf.shape
    x0, x1 = s[0], s[1]
    v_idx = np.searchsorted(m_cdf, x0, side="left")
    vi = np.minimum(v_idx, h - 1)                            # torch would raise on v_idx == h; never hit for x0 < m_cdf[-1]
    prev = np.where(v_idx > 0, m_cdf[np.maximum(vi - 1, 0)], np.float32(0))
    dv = np.where(v_idx > 0, (x0 - prev) / (m_cdf[vi] - prev), x0 / m_cdf[vi]).astype(np.float32)
    pdf_m = np.where(v_idx > 0, m_cdf[vi] - prev, m_cdf[vi]).astype(np.float32)
    v = (v_idx.astype(np.float32) + dv).astype(np.float32)
    u_idx = np.array([np.searchsorted(c_cdf[r], x, side="left") for r, x in zip(vi, x1)], dtype=np.int64)
    ui = u_idx.copy()
    ui1 = ui - 1; ui1[ui1 == -1] = 0
    ui[ui == 32] = 31                                        # hard-coded in the reference (:120)
    ui = np.minimum(ui, w - 1)
    pdf_c = np.where(u_idx > 0, c_cdf[vi, ui] - c_cdf[vi, ui1], c_cdf[vi, ui]).astype(np.float32)
    u = u_idx.astype(np.float32)
    theta = (v * np.float32(math.pi) / np.float32(h)).astype(np.float32)
    phi = (np.float32(2.0) * u * np.float32(math.pi) / np.float32(w)).astype(np.float32)
    dirs = angle2xyz(theta.astype(np.float32), phi.astype(np.float32)).astype(np.float32)
    two_pi_pi = np.float32(2.0 * math.pi * math.pi)
    pdf = (np.float32(h * w) * (pdf_c * pdf_m) / (two_pi_pi * np.sin(theta) + np.float32(1e-6))).astype(np.float32)
    return dirs, pdf.reshape(-1, 1), v_idx.reshape(-1, 1).astype(np.int64), u_idx.reshape(-1, 1)


def lookup_envmap(env, w):
    """envmap_utils.py:29-36 (nearest texel, y-up)."""
    env = np.asarray(env, np.float32); w = np.asarray(w, np.float32)
    height, width = env.shape[:2]
    phi = np.arctan2(w[..., 0], -w[..., 2]) / np.float32(2.0 * math.pi)
    u = np.clip(np.mod(phi * width + width, width), 0, width - 1).astype(np.int32)
    theta = np.arccos(w[..., 1]) / np.float32(math.pi)
    v = np.clip(theta * height, 0, height - 1).astype(np.int32)
    return env[v, u]


# ====================================================================================== computeSH
LARR = np.array([0, 1, 1, 1, 2, 2, 2, 2, 2, 3, 3, 3, 3, 3, 3, 3, 4, 4, 4, 4, 4, 4, 4, 4, 4])
MARR = np.array([0, -1, 0, 1, -2, -1, 0, 1, 2, -3, -2, -1, 0, 1, 2, 3, -4, -3, -2, -1, 0, 1, 2, 3, 4])


def computeK(l=LARR, m=MARR):
    """computeSH.py:58-68 — note the float32 round-trip of the factorials."""
    m = np.abs(m)
    a = np.array([math.factorial(int(x)) for x in (l - m)], dtype=np.float32)
    b = np.array([math.factorial(int(x)) for x in (l + m)], dtype=np.float32)
    return np.sqrt((2 * l + 1) * a / b / 4 / np.pi)


def legendre(theta):
    """computeSH.py:13-56, order (l, |m|) as used by Integration/projection. Returns dict[(l,m)]."""
    c, s = np.cos(theta), np.sin(theta)
    return {(0, 0): np.ones_like(theta), (1, 0): c, (1, 1): -s,
            (2, 0): 0.5 * (3 * c ** 2 - 1), (2, 1): -3 * c * s, (2, 2): 3 * s ** 2,
            (3, 0): 0.5 * (5 * c ** 3 - 3 * c), (3, 1): -1.5 * (5 * c ** 2 - 1) * s, (3, 2): 15 * c * s ** 2, (3, 3): -15 * s ** 3,
            (4, 0): 0.125 * (35 * c ** 4 - 30 * c ** 2 + 3), (4, 1): -2.5 * (7 * c ** 3 - 3 * c) * s,
            (4, 2): 7.5 * (7 * c ** 2 - 1) * s ** 2, (4, 3): -105 * c * s ** 3, (4, 4): 105 * s ** 4}


def sh_basis(theta, phi, K=None):
    """The 25 real basis functions in the reference's order (Integration :87-162): (n, 25)."""
    K = computeK() if K is None else K
    P = legendre(theta)
    cols = []
    for i, (l, m) in enumerate(zip(LARR, MARR)):
        if m == 0:
            cols.append(K[i] * P[(l, 0)])
        elif m < 0:
            cols.append(np.sqrt(2) * K[i] * np.sin(-m * phi) * P[(l, -m)])
        else:
            cols.append(np.sqrt(2) * K[i] * np.cos(m * phi) * P[(l, m)])
    return np.stack(cols, 1)


def uv_to_envmap(im, u, v):
    """computeSH.py:75-85 bilinear fetch (vectorised)."""
    h, w = im.shape[:2]
    c, r = u * (w - 1), (1 - v) * (h - 1)
    cs, rs = c.astype(np.int64), r.astype(np.int64)
    ce, re = np.minimum(w - 1, cs + 1), np.minimum(h - 1, rs + 1)
    wc, wr = (c - cs)[:, None], (r - rs)[:, None]
    c1 = (1 - wc) * im[rs, cs] + wc * im[rs, ce]
    c2 = (1 - wc) * im[re, cs] + wc * im[re, ce]
    return (1 - wr) * c1 + wr * c2


def sh_angles(h, w, jitter):
    """Sample directions of computeSHFromImage (:299-311); jitter (h*w, 2) = the (y, x) uniforms it would draw."""
    r, c = np.divmod(np.arange(h * w), w)
    y = (r + jitter[:, 0]) / float(h); x = (c + jitter[:, 1]) / float(w)
    return np.stack([2 * np.arccos(np.sqrt(1 - y)), 2 * np.pi * x - np.pi], 1)        # (theta, phi)


def sh_project(im, angles):
    """computeSHFromImage (:299-347) for given sample angles: coef (25, 3)."""
    im = np.asarray(im, np.float64)
    theta, phi = angles[:, 0], angles[:, 1]
    u = (phi + np.pi) / 2 / np.pi; v = 1 - theta / np.pi
    colors = uv_to_envmap(im, u, v)
    return (4 * np.pi / angles.shape[0]) * sh_basis(theta, phi).T @ colors


def sh_reconstruct(coef, nrows, ncols, clip=True):
    """reconstImageFromSH (:226-240)."""
    x, y = np.meshgrid(np.linspace(-1, 1, ncols + 1), np.linspace(0, 1, nrows + 1))
    phi, theta = (np.pi * x)[:nrows, :ncols].reshape(-1), (np.pi * y)[:nrows, :ncols].reshape(-1)
    img = (sh_basis(theta, phi) @ np.asarray(coef, np.float64)).reshape(nrows, ncols, 3)
    return np.clip(img, 0, 1) if clip else img
