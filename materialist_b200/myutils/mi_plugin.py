"""Drop-in for the hot-path part of the reference's myutils/mi_plugin.py.

* `MatDiffBSDF` (mi_plugin.py:1229-1475): the spatially-varying Disney-diffuse + GGX BSDF whose a/r/m(/n) maps are
  looked up by projecting the hit point to the screen.  In the reference it is a Mitsuba `mi.BSDF` traced by
  Dr.Jit; here it is a plain object with the same methods operating on arrays of lanes ((L,3) CUDA tensors), each
  method one CUDA kernel launch (mb200_bsdf_eval_pdf / mb200_bsdf_sample).  The render operator fuses the same
  device functions into the shading kernels, so a lane-level call is exactly what a path vertex evaluates.
* the microfacet helpers D_GGX / G1_GGX_Schlick / G_Smith / fresnelSchlick (:60-97) and the projection helpers
  (:585-595, :645-671) as generic arithmetic, usable on torch tensors;
* the torch 'scratch' BRDF (:26-58, :136-177, :285-386) used by envmap_utils.sample_env1 / sample_brdf1.  NOTE it
  is a DIFFERENT formula from MatDiffBSDF (Lambert diffuse, a 2.0x factor, +1e-4) — kept as the reference has it.

* `TransBSDF` (mi_plugin.py:1477-1770): the transparency-editing plugin of trans_edit.py — MatDiffBSDF outside the edit
  mask, a diffuse + metal + glass (reflection / transmission with the background looked up at the twice-refracted
  screen position) BSDF inside; lane-level eval_pdf / sample / calculate_refracted_screen_coor, forward only.

Out of scope: MatBSDF, RefractBaseBRDF, MatrefractBSDF, BRDF4scratch (other editing features).
"""
import ctypes as C
import math

import numpy as np
import torch
import torch.nn.functional as NF

from .. import _abi
from ..scene import Camera


# ------------------------------------------------------------------ microfacet helpers (generic arithmetic)
def G1_GGX_Schlick(NoV, eta):
    k = (eta + 1)
    k = k * k / 8
    return 1 / (NoV * (1 - k) + k + 1e-6)


def G_Smith(NoV, NoL, eta):
    return G1_GGX_Schlick(NoL, eta) * G1_GGX_Schlick(NoV, eta)


def fresnelSchlick(VoH, F0):
    x = (1 - VoH) ** 5
    return F0 + (1 - F0) * x


def fresnelSchlick_sep(VoH):
    x = (1 - VoH) ** 5
    return (1 - x), x


def D_GGX(cos_h, eta):
    alpha2 = (eta * eta) ** 2
    denom = (cos_h * cos_h * (alpha2 - 1.0) + 1.0) + 1e-6
    return alpha2 / (math.pi * denom * denom)


def perspective_projection_matrix(fov, aspect, near, far):
    f = 1.0 / torch.tan(torch.as_tensor(fov, dtype=torch.float32) / 2.0)
    return torch.tensor([[f / aspect, 0, 0, 0], [0, f, 0, 0],
                         [0, 0, (far + near) / (near - far), (2 * far * near) / (near - far)], [0, 0, -1, 0]], dtype=torch.float32)


def mi_world_to_screen(world_coords, view_matrix, projection_matrix, screen_width, screen_height):
    """(L,3) -> (L,2) = (x_screen, y_screen); no y flip, no epsilon (mi_plugin.py:645-671)."""
    V = torch.as_tensor(view_matrix, dtype=world_coords.dtype, device=world_coords.device)
    P = torch.as_tensor(projection_matrix, dtype=world_coords.dtype, device=world_coords.device)
    h = torch.cat([world_coords, torch.ones_like(world_coords[..., :1])], -1)
    clip = (h @ V.T) @ P.T
    ndc = clip[..., :3] / clip[..., 3:4]
    return torch.stack([(ndc[..., 0] + 1) * 0.5 * screen_width, (ndc[..., 1] + 1) * 0.5 * screen_height], -1)


# ------------------------------------------------------------------ the torch 'scratch' BRDF (envmap_utils uses it)
def get_normal_space(normal):
    v1 = torch.zeros_like(normal); v1[..., 0] = 1.0
    up = torch.zeros_like(normal); up[..., 1] = 1.0
    use_x = (v1 * normal).sum(-1, keepdim=True).abs() <= 1e-1
    tangent = NF.normalize(torch.where(use_x, torch.cross(v1, normal, dim=-1), torch.cross(up, normal, dim=-1)), dim=-1)
    bitangent = torch.cross(normal, tangent, dim=-1)
    return torch.stack([tangent, bitangent, normal], dim=-1)


def angle2xyz(theta, phi):
    st = torch.sin(theta)
    return NF.normalize(torch.stack([st * torch.cos(phi), st * torch.sin(phi), torch.cos(theta)], dim=-1), dim=-1)


def diffuse_sampler(sample2, normal):
    wi = angle2xyz(torch.asin(sample2[..., 0].sqrt()), math.pi * 2 * sample2[..., 1])
    return (wi[:, None] @ get_normal_space(normal).permute(0, 2, 1)).squeeze(1)


def specular_sampler(sample2, roughness, wo, normal):
    roughness = torch.where(roughness <= 0.0, torch.ones_like(roughness), roughness)
    alpha = (roughness * roughness).squeeze(-1).detach()
    theta = torch.acos(((1 - sample2[..., 0]) / (sample2[..., 0] * (alpha * alpha - 1) + 1)).sqrt())
    wh = angle2xyz(theta, 2 * math.pi * sample2[..., 1])
    wh = (wh[:, None] @ get_normal_space(normal).permute(0, 2, 1)).squeeze(1)
    return NF.normalize(2 * (wo * wh).sum(-1, keepdim=True) * wh - wo, dim=-1)


def eval_brdf(wi, wo, normal_geo, mat, use_mesh_normal, *args, **kwargs):
    albedo, metallic = mat["albedo"].reshape(-1, 3), mat["metallic"].reshape(-1, 1)
    roughness = mat["roughness"].reshape(-1, 1)
    normal = normal_geo if use_mesh_normal else mat["normal"].reshape(-1, 3)
    h = NF.normalize(wi + wo, dim=-1)
    NoL = (wi * normal).sum(-1, keepdim=True).relu(); NoV = (wo * normal).sum(-1, keepdim=True).relu()
    VoH = (wo * h).sum(-1, keepdim=True).relu(); NoH = (normal * h).sum(-1, keepdim=True).relu()
    nn_ = lambda t: torch.nan_to_num(t, nan=0, posinf=0, neginf=0)
    D = nn_(D_GGX(NoH, roughness))
    pdf = 0.5 * (D.detach() / (4 * VoH.clamp_min(1e-6)) * NoH) + 0.5 * (NoL / math.pi)
    kd = albedo * (1 - metallic); ks = 0.04 * (1 - metallic) + albedo * metallic
    G = nn_(G_Smith(NoV, NoL, roughness)); F = nn_(fresnelSchlick(VoH, ks))
    brdf = 2.0 * (kd / math.pi + D * G * F / 4.0 * NoH) * NoL
    return nn_(brdf), nn_(pdf)


def sample_brdf(sample1, sample2, wo, normal_geo, mat, use_mesh_normal, *args, **kwargs):
    B, device = sample2.shape[0], sample2.device
    roughness = mat["roughness"].reshape(-1, 1)
    normal = normal_geo if use_mesh_normal else mat["normal"].reshape(-1, 3)
    if sample1 is None:
        sample1 = torch.rand(B, device=device)
    mask = sample1 > 0.5
    wi = torch.zeros(B, 3, device=device)
    wi[mask] = diffuse_sampler(sample2[mask], normal[mask])
    wi[~mask] = specular_sampler(sample2[~mask], roughness[~mask], wo[~mask], normal[~mask])
    brdf, pdf = eval_brdf(wi, wo, normal, mat, use_mesh_normal)
    w = torch.where(pdf > 0, brdf / (pdf + 1e-4), torch.zeros_like(brdf))
    return wi, pdf, torch.nan_to_num(w, nan=0, posinf=0, neginf=0)


# ------------------------------------------------------------------ MatDiffBSDF
class Frame3f:
    """mi.Frame3f(n): Duff et al. orthonormal basis; to_world / to_local on (L,3) tensors."""

    def __init__(self, n):
        sign = torch.copysign(torch.ones_like(n[..., 2]), n[..., 2])
        a = -1.0 / (sign + n[..., 2]); b = n[..., 0] * n[..., 1] * a
        self.s = torch.stack([sign * n[..., 0] * n[..., 0] * a + 1, sign * b, -sign * n[..., 0]], -1)
        self.t = torch.stack([b, n[..., 1] * n[..., 1] * a + sign, -n[..., 1]], -1)
        self.n = n

    def to_world(self, v):
        return self.s * v[..., 0:1] + self.t * v[..., 1:2] + self.n * v[..., 2:3]

    def to_local(self, v):
        return torch.stack([(v * self.s).sum(-1), (v * self.t).sum(-1), (v * self.n).sum(-1)], -1)


class SurfaceInteraction:
    """The fields of mi.SurfaceInteraction3f the plugin reads: p, n (geometric normal = shading frame), wi (local)."""

    def __init__(self, p, n, wi_local):
        self.p, self.n, self.wi = p.contiguous().float(), n.contiguous().float(), wi_local.contiguous().float()
        self.sh_frame = Frame3f(self.n)

    def to_world(self, v):
        return self.sh_frame.to_world(v)

    def to_local(self, v):
        return self.sh_frame.to_local(v)


class BSDFSample3f:
    def __init__(self, wo, pdf, flags):
        self.wo, self.pdf, self.eta = wo, pdf, 1.0
        self.sampled_component = 0
        self.sampled_type = flags


class BSDFFlags:
    DiffuseReflection, SpatiallyVarying, FrontSide = 0x2, 0x1000, 0x10000


class MatDiffBSDF:
    """props: dict with 'cam_meta' (path to the camera json) and 'use_mesh_normal' (mi_plugin.py:1230-1275)."""

    def __init__(self, props=None):
        props = props or {}
        self.m_flags = BSDFFlags.SpatiallyVarying | BSDFFlags.DiffuseReflection | BSDFFlags.FrontSide
        self.m_components = [self.m_flags]
        self.use_mesh_normal = bool(props.get("use_mesh_normal", True))
        self.camera = Camera.from_json(props.get("cam_meta"))
        self.width, self.height = self.camera.width, self.camera.height
        dev = props.get("device", "cuda")
        H, W = self.height, self.width
        self.a = torch.full((H, W, 3), 0.5, device=dev); self.r = torch.full((H, W, 1), 0.5, device=dev)
        self.m = torch.full((H, W, 1), 0.5, device=dev); self.n = torch.full((H, W, 3), 0.5, device=dev)
        self.view_matrix = self.camera.view_matrix
        self.persp_proj_matx = self.camera.proj_matrix
        self.flags = _abi.FLAG_WO_WORLD_QUIRK | _abi.FLAG_ENV_HALF_TEXEL | (_abi.FLAG_ROW_STRIDE_H if H == W else 0)

    # -- plumbing
    def _cfg(self):
        c = _abi.Cfg()
        c.H, c.W, c.spp, c.max_depth = self.height, self.width, 1, 4
        c.flags, c.use_mesh_normal, c.row0, c.rows = self.flags, int(self.use_mesh_normal), 0, self.height
        c.view[:] = self.view_matrix.reshape(-1).tolist(); c.proj[:] = self.persp_proj_matx.reshape(-1).tolist()
        c.cam_to_world[:] = self.camera.to_world.astype(np.float32).reshape(-1).tolist()
        c.tan_half_fov_x = self.camera.tan_half_fov_x
        return c

    def _maps(self):
        return [t.detach().contiguous().float() for t in (self.a, self.r, self.m)] + \
               [None if self.use_mesh_normal else self.n.detach().contiguous().float()]

    # -- the mi.BSDF protocol on lanes
    def eval_pdf(self, ctx, si, wo, active=True):
        """wo: local direction (light). Returns (brdf*cos (L,3), pdf (L)) — mi_plugin.py:1449-1460."""
        wo_w = si.to_world(wo).contiguous(); wi_w = si.to_world(si.wi).contiguous()
        L = wo_w.shape[0]
        f = torch.empty(L, 3, device=wo_w.device); pdf = torch.empty(L, device=wo_w.device)
        a, r, m, n = self._maps()
        cfg = self._cfg()
        _abi.check(_abi.lib.mb200_bsdf_eval_pdf(C.byref(cfg), L, _abi.ptr(si.p), _abi.ptr(si.n), _abi.ptr(wi_w), _abi.ptr(wo_w),
                                                _abi.ptr(a), _abi.ptr(r), _abi.ptr(m), _abi.ptr(n), _abi.ptr(f), _abi.ptr(pdf),
                                                _abi.stream_ptr()), "mb200_bsdf_eval_pdf")
        return f, pdf

    def eval_pdf_backward(self, ctx, si, wo, grad_f):
        """What `dr.backward` through `eval_pdf` delivers to the traversed parameters for a cotangent `grad_f` (L,3) on the rgb value
        (the pdf is detached by the path integrator): per lane (g_a (L,3), g_r (L), g_m (L), g_n (L,3)) at the lane's texel."""
        wo_w = si.to_world(wo).contiguous(); wi_w = si.to_world(si.wi).contiguous()
        L, dev = wo_w.shape[0], wo_w.device
        ga = torch.empty(L, 3, device=dev); gr = torch.empty(L, device=dev); gm = torch.empty(L, device=dev); gn = torch.empty(L, 3, device=dev)
        a, r, m, n = self._maps()
        cfg = self._cfg()
        _abi.check(_abi.lib.mb200_bsdf_eval_grad(C.byref(cfg), L, _abi.ptr(si.p), _abi.ptr(si.n), _abi.ptr(wi_w), _abi.ptr(wo_w),
                                                 _abi.ptr(a), _abi.ptr(r), _abi.ptr(m), _abi.ptr(n), _abi.ptr(grad_f.contiguous().float()),
                                                 _abi.ptr(ga), _abi.ptr(gr), _abi.ptr(gm), _abi.ptr(gn), _abi.stream_ptr()), "mb200_bsdf_eval_grad")
        return ga, gr, gm, gn

    def sample(self, ctx, si, sample1, sample2, active=True):
        """Returns (BSDFSample3f, weight (L,3)); bs.wo is in WORLD space like the reference (mi_plugin.py:1444)."""
        wi_w = si.to_world(si.wi).contiguous()
        L = wi_w.shape[0]
        wo = torch.empty(L, 3, device=wi_w.device); pdf = torch.empty(L, device=wi_w.device); w = torch.empty(L, 3, device=wi_w.device)
        a, r, m, n = self._maps()
        cfg = self._cfg()
        _abi.check(_abi.lib.mb200_bsdf_sample(C.byref(cfg), L, _abi.ptr(si.p), _abi.ptr(si.n), _abi.ptr(wi_w),
                                              _abi.ptr(sample1.contiguous().float()), _abi.ptr(sample2.contiguous().float()),
                                              _abi.ptr(a), _abi.ptr(r), _abi.ptr(m), _abi.ptr(n), _abi.ptr(wo), _abi.ptr(pdf), _abi.ptr(w),
                                              _abi.stream_ptr()), "mb200_bsdf_sample")
        return BSDFSample3f(wo, pdf, self.m_flags), w

    def eval(self, ctx, si, wo, active=True):
        """mi_plugin.py:1349-1357 — note the reference passes (wi, wo) in the OPPOSITE order from eval_pdf here
        (the cosine lands on the view direction); reproduced. Not called by the path integrator."""
        swapped = SurfaceInteraction(si.p, si.n, wo)
        return self.eval_pdf(ctx, swapped, si.wi)[0]

    def pdf(self, ctx, si, wo, active=True):
        swapped = SurfaceInteraction(si.p, si.n, wo)
        return self.eval_pdf(ctx, swapped, si.wi)[1]

    def traverse(self, callback):
        for k in ("a", "r", "m", "n"):
            callback.put_parameter(k, getattr(self, k), "Differentiable")
        callback.put_parameter("use_mesh_normal", self.use_mesh_normal, "NonDifferentiable")

    def parameters_changed(self, keys=None):
        pass

    def to_string(self):
        return "MatDiffBSDF"


class TransBSDF(MatDiffBSDF):
    """props: MatDiffBSDF's plus 'ior' (default 1.3) and 'keep_albedo_color' (its PRESENCE sets refract_distance = 100,
    mi_plugin.py:1484-1489).  Attributes specTrans / bg / mask / ior are what TransBSDF.traverse exposes (:1763-1770)."""

    def __init__(self, props=None):
        super().__init__(props)
        props = props or {}
        self.ior = float(props.get("ior", 1.3))
        if "keep_albedo_color" in props:
            self.keep_albedo_color = bool(props["keep_albedo_color"]); self.refract_distance = 1.0 * 100
        else:
            self.keep_albedo_color = False; self.refract_distance = 1.0
        self.specTrans = 0.8
        dev = self.a.device
        self.bg = torch.full((self.height, self.width, 3), 0.5, device=dev)
        self.mask = torch.zeros(self.height, self.width, dtype=torch.bool, device=dev)

    def _trans(self):
        self._keep = (self.bg.detach().contiguous().float(), (self.mask != 0).to(torch.uint8).contiguous())
        return _abi.Trans(float(self.ior), float(self.specTrans), float(self.refract_distance), 0, self._keep[0].data_ptr(), self._keep[1].data_ptr())

    def calculate_refracted_screen_coor(self, wi, normal, ior_ratio, position, screen_coor=None):
        """(L,2) screen position of the background texel seen through the edited surface (:1503-1519); `ior_ratio` is accepted
        for signature compatibility — the reference always passes 1/self.ior (:1528, :1757) and so does the kernel."""
        L = wi.shape[0]
        sc = torch.empty(L, 2, device=wi.device); flat = torch.empty(L, dtype=torch.int64, device=wi.device)
        cfg = self._cfg(); t = self._trans()
        _abi.check(_abi.lib.mb200_trans_refracted_texel(C.byref(cfg), C.byref(t), L, _abi.ptr(position.contiguous()), _abi.ptr(normal.contiguous()),
                                                        _abi.ptr(wi.contiguous()), _abi.ptr(sc), _abi.ptr(flat), _abi.stream_ptr()),
                   "mb200_trans_refracted_texel")
        return sc

    def eval_pdf(self, ctx, si, wo, active=True):
        wo_w = si.to_world(wo).contiguous(); wi_w = si.to_world(si.wi).contiguous()
        L = wo_w.shape[0]
        f = torch.empty(L, 3, device=wo_w.device); pdf = torch.empty(L, device=wo_w.device)
        a, r, m, n = self._maps()
        cfg = self._cfg(); t = self._trans()
        _abi.check(_abi.lib.mb200_trans_eval_pdf(C.byref(cfg), C.byref(t), L, _abi.ptr(si.p), _abi.ptr(si.n), _abi.ptr(wi_w), _abi.ptr(wo_w),
                                                 _abi.ptr(a), _abi.ptr(r), _abi.ptr(m), _abi.ptr(n), _abi.ptr(f), _abi.ptr(pdf),
                                                 _abi.stream_ptr()), "mb200_trans_eval_pdf")
        return f, pdf

    def sample(self, ctx, si, sample1, sample2, active=True):
        wi_w = si.to_world(si.wi).contiguous()
        L = wi_w.shape[0]
        wo = torch.empty(L, 3, device=wi_w.device); pdf = torch.empty(L, device=wi_w.device); w = torch.empty(L, 3, device=wi_w.device)
        a, r, m, n = self._maps()
        cfg = self._cfg(); t = self._trans()
        _abi.check(_abi.lib.mb200_trans_sample(C.byref(cfg), C.byref(t), L, _abi.ptr(si.p), _abi.ptr(si.n), _abi.ptr(wi_w),
                                               _abi.ptr(sample1.contiguous().float()), _abi.ptr(sample2.contiguous().float()),
                                               _abi.ptr(a), _abi.ptr(r), _abi.ptr(m), _abi.ptr(n), _abi.ptr(wo), _abi.ptr(pdf), _abi.ptr(w),
                                               _abi.stream_ptr()), "mb200_trans_sample")
        bs = BSDFSample3f(wo, pdf, self.m_flags)
        bs.eta = self.ior                                   # :1542
        return bs, w

    def traverse(self, callback):
        for k in ("a", "r", "m", "bg", "mask", "specTrans", "ior"):
            callback.put_parameter(k, getattr(self, k), "Differentiable")

    def to_string(self):
        return "TransBSDF"


_REGISTRY = {}


def register_bsdf(name, factory):
    """mi.register_bsdf stand-in (inverse_img_w_mi.py:6)."""
    _REGISTRY[name] = factory
