#!/usr/bin/env python
"""Envmap-phase iteration (inverse_img_w_mi.py:237-256: render_envmap forward + adjoint with gradients to the envmap
texels) timing at the reference's 16x32 learned envmap and at 256x128; prints one JSON line per case."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import materialist_b200 as mb  # noqa: E402
from materialist_b200 import renderop as mbr, synthetic  # noqa: E402
from materialist_b200.inverse import EnvmapOptimizer  # noqa: E402
from materialist_b200.scene import Camera  # noqa: E402


def main():
    H = W = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    spp = 64
    cam = Camera(width=W, height=H)
    pos, nrm, valid = synthetic.gbuffer(H, W, cam)
    a, r, m = (t.cuda() for t in synthetic.materials(H, W, seed_base=1))
    for He, We, sun in ((16, 32, 0.0), (16, 32, 2000.0), (128, 256, 2000.0)):
        env = synthetic.envmap(He, We, seed=4, sun=sun)
        scene = mb.Scene(pos, nrm, valid, camera=cam, envmap=env, device="cuda")
        scene.a, scene.r, scene.m = a, r, m
        gt = mb.render(scene, spp=spp, seed=999)
        opt = EnvmapOptimizer(scene, torch.log(torch.expm1(env.cuda().clamp_min(1e-3))), gt, spp=spp)
        for i in range(3):
            opt.step(i)
        torch.cuda.synchronize()
        mbr.KERNEL_EVENTS = []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 10
        e0.record()
        for i in range(n):
            opt.step(100 + i)
        e1.record(); torch.cuda.synchronize()
        kt = {}
        for name, x, y in mbr.KERNEL_EVENTS:
            kt.setdefault(name, []).append(x.elapsed_time(y))
        mbr.KERNEL_EVENTS = None
        print(json.dumps({"image": [H, W], "spp": spp, "envmap": [He, We], "sun": sun, "ms_per_step": e0.elapsed_time(e1) / n,
                          "kernel_ms": {k: sum(v) / len(v) for k, v in kt.items()}}), flush=True)


if __name__ == "__main__":
    main()
