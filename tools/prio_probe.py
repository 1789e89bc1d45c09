"""Probe: does a high-priority main stream let the film-weights kernel (side stream, default priority) fill only the idle slots of
the iteration?  C2, fused optimiser; {default, high-priority} main stream x film weights {after, before} the forward render."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
import bench
import materialist_b200 as mb
from materialist_b200.inverse import FusedBRDFOptimizer
from materialist_b200.parallel import ShardContext

dev = torch.device("cuda:0"); torch.cuda.set_device(dev)
wl = bench.WORKLOADS["c2"]
case = bench.build_case(wl, 1)
H, W, spp = case["H"], case["W"], wl["spp"]
scene = mb.Scene(case["pos"], case["nrm"], case["valid"], camera=case["cam"], envmap=case["env"], device=dev)
to = lambda t: t.to(dev)
gt = mb.render(scene, spp=64, seed=999, albedo=to(case["a2"]), roughness=to(case["r2"]), metallic=to(case["m2"]))
for prio in (0, -1, -5):
    for early in ("0", "1", "2"):
        os.environ["MB200_FILM_WEIGHTS_EARLY"] = early
        mat = {"albedo": to(case["a"]), "roughness": to(case["r"]), "metallic": to(case["m"])}
        st = torch.cuda.Stream(dev, priority=prio) if prio != 0 else torch.cuda.current_stream(dev)
        with torch.cuda.stream(st):
            opt = FusedBRDFOptimizer(scene, mat, gt, "arm", spp=spp, shard=ShardContext(H, W, 0, 1))
            for i in range(5):
                opt.step(i)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(40):
                opt.step(1000 + i)
            e1.record(); torch.cuda.synchronize()
        print(f"main-stream priority {prio:2d}  film weights { {'0': 'after the forward render', '1': 'launched before the forward render', '2': 'launched behind the forward render, no dependency on it'}[early]}: {e0.elapsed_time(e1) / 40:.4f} ms / iteration", flush=True)
