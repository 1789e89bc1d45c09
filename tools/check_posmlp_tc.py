#!/usr/bin/env python
"""PosMLP tcgen05 forward vs the FFMA kernels vs plain torch (fp32, TF32 off / float64): errors and timing."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from materialist_b200 import _abi  # noqa: E402
from materialist_b200.mymodels.mlps import PosMLP  # noqa: E402
from bench_posmlp import torch_forward, timeit  # noqa: E402


def main():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    for (H, W) in ((8, 16), (37, 53), (512, 512)):
        net = PosMLP(in_dims=7, out_dims=5, dims=[256] * 4, skip_connection=[1, 3], weight_norm=False, multires_view=2,
                     output_type="arm", color_ch=5).cuda()
        with torch.no_grad():
            net.lin4.weight.normal_(0, 0.05); net.lin4.bias.normal_(0, 0.05)
        img = torch.rand(H * W, 5, device="cuda")
        res = {"H": H, "W": W}
        with torch.no_grad():
            y_ref64 = torch_forward(net.double(), img.double(), H, W).float(); net.float()
            y_t = torch_forward(net, img, H, W)
            for name, impl in (("tc", _abi.POSMLP_TCGEN05), ("ffma", _abi.POSMLP_FFMA)):
                net.impl = impl
                y = net(img, hw=(H, W))
                torch.cuda.synchronize()
                res[name + "_err_vs_f64"] = float(((y - y_ref64).norm() / y_ref64.norm()).item())
                res[name + "_maxabs"] = float((y - y_ref64).abs().max().item())
                res[name + "_ms"] = timeit(lambda: net(img, hw=(H, W)), n=5, warm=2)
            res["torch32_err_vs_f64"] = float(((y_t - y_ref64).norm() / y_ref64.norm()).item())
        # gradients: tcgen05 backward and FFMA backward against float64 autograd of the plain-torch network
        gy = torch.randn(H * W, 5, device="cuda") * 1e-4
        net.double(); net.zero_grad(set_to_none=True)
        torch_forward(net, img.double(), H, W).backward(gy.double())
        ref = [p.grad.float().clone() for p in net.parameters()]
        net.float()
        names = [n for n, _ in net.named_parameters()]
        for name, impl in (("tc", _abi.POSMLP_TCGEN05), ("ffma", _abi.POSMLP_FFMA)):
            net.impl = impl
            net.zero_grad(set_to_none=True)
            net(img, hw=(H, W)).backward(gy)
            torch.cuda.synchronize()
            errs = {n: float(((p.grad - r).norm() / r.norm().clamp_min(1e-30)).item()) for n, p, r in zip(names, net.parameters(), ref)}
            res[name + "_grad_err_max"] = max(errs.values())
            res[name + "_grad_err_worst"] = max(errs, key=errs.get)
            if impl == _abi.POSMLP_TCGEN05:
                res["tc_grad_errs"] = {k: round(v, 9) for k, v in errs.items()}

            def fb():
                net.zero_grad(set_to_none=True)
                net(img, hw=(H, W)).backward(gy)
            res[name + "_fwd_bwd_ms"] = timeit(fb, n=5, warm=2)
        flops = H * W * 2 * (15 * 241 + 256 * 256 + 256 * 241 + 256 * 256 + 256 * 5)
        res["tc_fwd_tflops"] = flops / res["tc_ms"] / 1e9
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
