#!/usr/bin/env python
"""torchrun --nproc-per-node 2 tools/check_mesh_shard_equivalence.py
Mesh mode under row sharding: K iterations of FusedBRDFOptimizer on 2 ranks must reproduce the 1-GPU run (paths cross shard
borders at their secondary vertices: the map gradients are all-reduced and every rank steps the whole image)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import materialist_b200 as mb  # noqa: E402
from materialist_b200 import synthetic  # noqa: E402
from materialist_b200.inverse import FusedBRDFOptimizer  # noqa: E402
from materialist_b200.parallel import ShardContext  # noqa: E402


def run(dev, shard, K=3, H=64, W=64, spp=32):
    cam = mb.Camera(width=W, height=H)
    verts, tris = synthetic.grid_mesh(synthetic.bumpy_positions(H, W, cam))
    env = synthetic.envmap(16, 32, seed=4)
    scene = mb.Scene.from_mesh(verts, tris, cam, device=dev, envmap=env)
    a, r, m = (t.to(dev) for t in synthetic.materials(H, W, seed_base=1))
    a2, r2, m2 = (t.to(dev) for t in synthetic.materials(H, W, seed_base=5))
    scene.set_shard(0, H)
    gt = mb.render(scene, spp=spp, seed=999, albedo=a2, roughness=r2, metallic=m2)
    opt = FusedBRDFOptimizer(scene, {"albedo": a, "roughness": r, "metallic": m}, gt, "arm", spp=spp, lr=0.01, shard=shard)
    for k in range(K):
        opt.step(100 + k)
    return {k: v.clone() for k, v in opt.mat.items()}, shard


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    sharded, sh = run(dev, ShardContext(64, 64, rank, world))
    single, _ = run(dev, ShardContext(64, 64, 0, 1))
    ok = True
    for k in sharded:
        e = float((sharded[k] - single[k]).abs().max()); moved = float((single[k] - single[k].mean()).abs().max())
        d = float((sharded[k] - single[k]).norm() / single[k].norm())
        print(f"rank {rank} {k}: max |sharded - single| = {e:.3e}, rel-L2 {d:.3e}", flush=True)
        ok &= d < 1e-5
    dist.barrier()
    dist.destroy_process_group()
    if not ok:
        raise SystemExit("sharded mesh-mode optimisation differs from the single-GPU run")
    if rank == 0:
        print("OK: 2-rank mesh-mode optimisation == 1-GPU run")


if __name__ == "__main__":
    main()
