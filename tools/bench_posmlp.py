#!/usr/bin/env python
"""PosMLP micro-benchmark: fused sm_100a kernels vs the same network in stock PyTorch (cuBLAS FP32, TF32 off),
forward + backward, N = H*W pixels.  Prints one JSON line."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from materialist_b200.mymodels.mlps import PosMLP  # noqa: E402


def torch_forward(net, img, H, W):
    """mymodels/mlps.py:211-234 in plain torch (what the reference executes)."""
    r, c = torch.meshgrid(torch.arange(H, device=img.device), torch.arange(W, device=img.device), indexing="ij")
    p = torch.stack([r.flatten(), c.flatten()], 1).float()
    pts = torch.cat([p, torch.sin(p), torch.cos(p), torch.sin(p * 2), torch.cos(p * 2), img], 1)
    x = pts
    for l in range(5):
        lin = getattr(net, f"lin{l}")
        if l in (1, 3):
            x = torch.cat([x, pts], -1)
        x = lin(x)
    x = 1.3 * torch.tanh(x) + img
    return x.clamp(0, 1).detach() + x - x.detach()


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    H = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    W = int(sys.argv[2]) if len(sys.argv) > 2 else H
    torch.backends.cuda.matmul.allow_tf32 = False
    net = PosMLP(in_dims=7, out_dims=5, dims=[256] * 4, skip_connection=[1, 3], weight_norm=False, multires_view=2,
                 output_type="arm", color_ch=5).cuda()
    with torch.no_grad():
        net.lin4.weight.normal_(0, 0.05); net.lin4.bias.normal_(0, 0.05)
    img = torch.rand(H * W, 5, device="cuda")
    gy = torch.randn(H * W, 5, device="cuda")

    def ours():
        net.zero_grad(set_to_none=True)
        net(img, hw=(H, W)).backward(gy)

    def ours_fwd():
        with torch.no_grad():
            net(img, hw=(H, W))

    def ref():
        net.zero_grad(set_to_none=True)
        torch_forward(net, img, H, W).backward(gy)

    y1 = net(img, hw=(H, W)); y2 = torch_forward(net, img, H, W)
    err = float(((y1 - y2).norm() / y2.norm()).item())
    flops = H * W * 2 * (15 * 241 + 256 * 256 + 256 * 241 + 256 * 256 + 256 * 5)
    t_f, t_fb, t_ref = timeit(ours_fwd), timeit(ours), timeit(ref)
    print(json.dumps({"N": H * W, "fused_fwd_ms": t_f, "fused_fwd_bwd_ms": t_fb, "torch_fwd_bwd_ms": t_ref, "rel_l2_vs_torch": err,
                      "fwd_tflops": flops / t_f / 1e9, "fwd_bwd_tflops": 3 * flops / t_fb / 1e9}))


if __name__ == "__main__":
    main()
