"""Drop-in for the reference's mymodels/mlps.py: `PosMLP` (mlps.py:129-251) with the same constructor, parameter
names (`lin{l}.linear.weight|bias`, `lin4.weight|bias` — reference state_dicts load as they are) and forward
semantics, executed by the fused sm_100a kernels (mb200_posmlp_fwd / _bwd) instead of five cuBLAS SGEMMs.

Supported instantiations = the ones the reference makes (inverse_img_w_mi.py:117-124, :163):
dims=[256]*4, skip_connection=[1,3], multires_view=2, weight_norm=False, output_type in {'envmap', 'arm'}.
Anything else raises NotImplementedError (no silent fallback).  The reference infers the pixel grid from N
(img2points, mlps.py:190-198: sqrt(N) x sqrt(N), or h x 2h when N <= 512); pass `hw=(H, W)` for other shapes.
The per-layer `torch.isnan(...).any()` host syncs of the reference (mlps.py:218-229) are not reproduced.
"""
import ctypes as C
import math

import numpy as np
import torch
import torch.nn as nn

from .. import _abi


def get_embedder(multires, input_dims):
    """mlps.py:42-54 — kept for API compatibility (the kernels embed in-register)."""
    freqs = 2.0 ** torch.linspace(0.0, multires - 1, multires)

    def embed(x):
        out = [x]
        for f in freqs:
            out += [torch.sin(x * f), torch.cos(x * f)]
        return torch.cat(out, -1)
    return embed, input_dims * (1 + 2 * multires)


class SineLayer(nn.Module):
    """mlps.py:69-103 (sin(Wx+b); omega_0 stored but unused; default nn.Linear init)."""

    def __init__(self, in_features, out_features, bias=True, is_first=False, omega_0=30, weight_norm=False):
        super().__init__()
        if weight_norm:
            raise NotImplementedError("weight_norm=True is not used by the reference scripts and is not implemented")
        self.omega_0, self.is_first, self.in_features = omega_0, is_first, in_features
        self.linear = nn.Linear(in_features, out_features, bias=bias)

    def forward(self, input):
        return torch.sin(self.linear(input))


class _PosMLPFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, desc, img, flat):
        N = img.shape[0]
        out = torch.empty(N, desc.n_out, device=img.device)
        need_grad = flat.requires_grad or img.requires_grad
        cache = None
        if need_grad:
            nbytes = _abi.lib.mb200_posmlp_cache_bytes(C.byref(desc), N)
            cache = torch.empty((nbytes + 3) // 4, dtype=torch.float32, device=img.device)
        wbytes = _abi.lib.mb200_posmlp_workspace_bytes(C.byref(desc))
        work = torch.empty((wbytes + 3) // 4, dtype=torch.float32, device=img.device) if wbytes else None
        _abi.check(_abi.lib.mb200_posmlp_fwd(C.byref(desc), _abi.fptr(flat), _abi.fptr(img), N, _abi.fptr(out), _abi.ptr(cache),
                                             _abi.ptr(work), _abi.stream_ptr()), "mb200_posmlp_fwd")
        ctx.desc, ctx.cache, ctx.N, ctx.work = desc, cache, N, work
        ctx.save_for_backward(img, flat)
        return out

    @staticmethod
    def backward(ctx, g_out):
        img, flat = ctx.saved_tensors
        g_flat = torch.zeros_like(flat)
        g_img = torch.empty_like(img) if ctx.needs_input_grad[1] else None
        _abi.check(_abi.lib.mb200_posmlp_bwd(C.byref(ctx.desc), _abi.fptr(flat), _abi.fptr(img), ctx.N, _abi.ptr(ctx.cache),
                                             _abi.ptr(g_out.contiguous().float()), _abi.ptr(g_flat), _abi.ptr(g_img), _abi.ptr(ctx.work),
                                             _abi.stream_ptr()),
                   "mb200_posmlp_bwd")
        return None, g_img, g_flat


class PosMLP(nn.Module):
    def __init__(self, in_dims, out_dims, dims, skip_connection=(), weight_norm=True, multires_view=0, output_type="envmap",
                 color_ch=5):
        super().__init__()
        if weight_norm or list(dims) != [256] * 4 or list(skip_connection) != [1, 3] or multires_view != 2 or \
                output_type not in ("envmap", "arm"):
            raise NotImplementedError(
                "materialist_b200.PosMLP implements the reference's instantiations only: dims=[256]*4, skip_connection=[1,3], "
                "multires_view=2, weight_norm=False, output_type in ('envmap','arm')")
        self.init_range = np.sqrt(3 / dims[0])
        self.output_type, self.color_ch, self.out_dims = output_type, color_ch, out_dims
        if output_type == "arm" and out_dims != color_ch:
            raise ValueError("'arm' adds the input image to the output: out_dims must equal color_ch")
        d0 = 10 + color_ch                          # dims[0] += (input_ch - in_dims) + color_ch with input_ch = 10 (mlps.py:147-152)
        if in_dims + (10 - in_dims) + color_ch != d0:
            raise ValueError("inconsistent in_dims")
        sizes = [d0, 256, 256, 256, 256, out_dims]
        self.num_layers, self.skip_connection = len(sizes), list(skip_connection)
        for l in range(5):
            out_dim = sizes[l + 1] - sizes[0] if (l + 1) in self.skip_connection else sizes[l + 1]
            if l < 4:
                lin = SineLayer(sizes[l], out_dim, True, False, 1, False)
            else:
                lin = nn.Linear(sizes[l], out_dim)
                nn.init.zeros_(lin.weight); nn.init.zeros_(lin.bias)          # mlps.py:174-176
            setattr(self, "lin" + str(l), lin)
        self.last_active_fun = nn.Softplus()
        self.impl = _abi.POSMLP_TCGEN05            # _abi.POSMLP_FFMA selects the FP32-FFMA kernels (A/B measurement only)

    # ------------------------------------------------------------------
    def _linears(self):
        return [getattr(self, f"lin{l}").linear if l < 4 else self.lin4 for l in range(5)]

    def flat_params(self):
        """[W0 | b0 | W1 | b1 | ... | W4 | b4] — the packing mb200_posmlp_* expects; differentiable w.r.t. the parameters."""
        parts = []
        for lin in self._linears():
            parts += [lin.weight.reshape(-1), lin.bias.reshape(-1)]
        return torch.cat(parts)

    def _desc(self, N, hw, row0=0):
        if hw is None and row0:
            raise ValueError("row0 needs an explicit hw=(H, W)")
        if hw is None:                              # img2points, mlps.py:190-198
            if N > 512:
                h = int(round(N ** 0.5))
                if h * h != N:
                    raise ValueError("N is not a square number: pass hw=(H, W)")
                hw = (h, h)
            else:
                h = (N / 2) ** 0.5
                if not float(h).is_integer():
                    raise ValueError("width should be double of height")
                hw = (int(h), 2 * int(h))
        if N % hw[1] != 0 or row0 < 0 or row0 + N // hw[1] > hw[0]:
            raise ValueError("img must hold whole rows [row0, row0 + N / W) of the (H, W) image")
        d = _abi.PosMLPDesc()
        d.n_color, d.n_out, d.hidden, d.n_freq = self.color_ch, self.out_dims, 256, 2
        d.output_type = 0 if self.output_type == "envmap" else 1
        d.H, d.W = hw
        d.impl = self.impl
        d.row0 = int(row0)
        return d

    def forward(self, img, hw=None, row0=0):
        """img: (N, color_ch) = the pixels of rows [row0, row0 + N / W) of an (H, W) image, row-major (`hw`, `row0` default to the
        whole image as the reference's img2points lays it out).  Row shards give bitwise the rows of the full evaluation."""
        if img.ndim != 2 or img.shape[1] != self.color_ch:
            raise ValueError(f"img must be (N, {self.color_ch})")
        if not img.is_cuda:
            raise ValueError("PosMLP runs on CUDA tensors only (no CPU fallback)")
        img = img.contiguous().float()
        desc = self._desc(img.shape[0], hw, row0)
        flat = self.flat_params()
        if flat.numel() != _abi.lib.mb200_posmlp_param_count(C.byref(desc)):
            raise RuntimeError("parameter packing mismatch")
        y = _PosMLPFn.apply(desc, img, flat)
        if self.output_type == "arm":
            # x.clamp(0,1).detach() + x - x.detach(): the kernel returns the clamped value and a straight-through gradient
            pass
        return y
