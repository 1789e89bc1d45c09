"""Per-iteration times of the fused C2 iteration right after a short warm-up (is the timed region of a short bench run steady?)."""
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
mat = {"albedo": to(case["a"]), "roughness": to(case["r"]), "metallic": to(case["m"])}
opt = FusedBRDFOptimizer(scene, mat, gt, "arm", spp=spp, shard=ShardContext(H, W, 0, 1))
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 5):
    opt.step(i)
torch.cuda.synchronize()
n = 40
ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
ev[0].record()
for i in range(n):
    opt.step(1000 + i)
    ev[i + 1].record()
torch.cuda.synchronize()
ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(n)]
print("per-step ms:", " ".join(f"{t:.3f}" for t in ts))
print(f"first 20: {sum(ts[:20]) / 20:.4f}  last 20: {sum(ts[20:]) / 20:.4f}")
