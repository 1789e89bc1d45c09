#!/usr/bin/env python
"""torchrun --nproc-per-node 2 tools/check_shard_equivalence.py
Row sharding must reproduce the 1-GPU run:
  * mesh mode, FusedBRDFOptimizer: paths cross shard borders at their secondary vertices -> the map gradients are all-reduced
    and every rank steps the whole image;
  * G-buffer mode, FusedBRDFOptimizer / DirectBRDFOptimizer: a rank steps its own rows; its forward reads the neighbours' maps in the
    2-row film halo, so the stepped boundary rows are exchanged after every iteration (ShardContext.map_halo_exchange);
  * G-buffer mode, PosMLPBRDFOptimizer (model_name=pos_mlp): every rank evaluates brdf_net on its own rows + film halo only
    (PosMLP row0), weight gradients summed over the ranks."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import materialist_b200 as mb  # noqa: E402
from materialist_b200 import synthetic  # noqa: E402
from materialist_b200.inverse import DirectBRDFOptimizer, FusedBRDFOptimizer, PosMLPBRDFOptimizer  # noqa: E402
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


def run_gbuffer(dev, shard, cls, K=5, H=64, W=64, spp=32):
    """G-buffer mode with the maps as direct parameters: returns this rank's OWN rows of the optimised maps."""
    cam = mb.Camera(width=W, height=H)
    pos, nrm, valid = synthetic.gbuffer(H, W, cam)
    scene = mb.Scene(pos, nrm, valid, camera=cam, envmap=synthetic.envmap(16, 32, seed=4), device=dev)
    a, r, m = (t.to(dev) for t in synthetic.materials(H, W, seed_base=1))
    a2, r2, m2 = (t.to(dev) for t in synthetic.materials(H, W, seed_base=5))
    gt = mb.render(scene, spp=spp, seed=999, albedo=a2, roughness=r2, metallic=m2)
    opt = cls(scene, {"albedo": a, "roughness": r, "metallic": m}, gt, "arm", spp=spp, lr=0.02, shard=shard)
    for k in range(K):
        opt.step(100 + k)
    assert (scene.row0, scene.rows) == (0, H)                 # the optimiser leaves the scene unsharded
    maps = opt.mat if cls is FusedBRDFOptimizer else {k: v.detach() for k, v in opt.params.items()}
    out = {k: v.clone() for k, v in maps.items()}
    if getattr(opt, "peer", None) is not None:
        print(f"rank {shard.rank}: {cls.__name__} exchanged over peer memory (no NCCL in the iteration)", flush=True)
    if hasattr(opt, "close"):
        opt.close()
    return out


def run_posmlp(dev, shard, K=3, H=64, W=64, spp=32):
    cam = mb.Camera(width=W, height=H)
    pos, nrm, valid = synthetic.gbuffer(H, W, cam)
    scene = mb.Scene(pos, nrm, valid, camera=cam, envmap=synthetic.envmap(16, 32, seed=4), device=dev)
    a, r, m = (t.to(dev) for t in synthetic.materials(H, W, seed_base=1))
    a2, r2, m2 = (t.to(dev) for t in synthetic.materials(H, W, seed_base=5))
    scene.set_shard(0, H)
    gt = mb.render(scene, spp=spp, seed=999, albedo=a2, roughness=r2, metallic=m2)
    torch.manual_seed(0)
    opt = PosMLPBRDFOptimizer(scene, {"albedo": a, "roughness": r, "metallic": m}, gt, "arm", spp=spp, lr=1e-3, shard=shard)
    with torch.no_grad():
        opt.net.lin4.weight.normal_(0, 0.02, generator=torch.Generator(device=dev).manual_seed(1))
    losses = [float(opt.step(100 + k)) for k in range(K)]
    return {"params": opt.net.flat_params().detach().clone(), "losses": torch.tensor(losses)}, shard


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
    # G-buffer mode: an even split (64 rows) and an UNEVEN one (70 rows x 68 columns: the ranks' halo buffers differ in size, so the
    # peer arenas have different layouts and the neighbours' offset tables are what addresses them)
    for (Hg, Wg) in ((64, 64), (70, 68)):
        for cls in (FusedBRDFOptimizer, DirectBRDFOptimizer):
            shc = ShardContext(Hg, Wg, rank, world)
            sharded = run_gbuffer(dev, shc, cls, H=Hg, W=Wg)
            single = run_gbuffer(dev, ShardContext(Hg, Wg, 0, 1), cls, H=Hg, W=Wg)
            rows = slice(shc.row0, shc.row0 + shc.rows)
            for k in sharded:
                d = float((sharded[k][rows] - single[k][rows]).norm() / single[k][rows].norm())
                lim = 1e-5
                print(f"rank {rank} G-buffer {Hg}x{Wg} {cls.__name__} {k}: own rows rel-L2 sharded vs single {d:.3e}", flush=True)
                ok &= d < lim
    sharded, _ = run_posmlp(dev, ShardContext(64, 64, rank, world))
    single, _ = run_posmlp(dev, ShardContext(64, 64, 0, 1))
    d = float((sharded["params"] - single["params"]).norm() / single["params"].norm())
    moved = float((single["params"]).norm())
    print(f"rank {rank} pos_mlp: params rel-L2 sharded vs single {d:.3e}; losses {sharded['losses'].tolist()} vs {single['losses'].tolist()}", flush=True)
    ok &= d < 1e-4 and bool(torch.allclose(sharded["losses"], single["losses"], rtol=1e-4))
    dist.barrier()
    dist.destroy_process_group()
    if not ok:
        raise SystemExit("a sharded optimisation differs from the single-GPU run")
    if rank == 0:
        print(f"OK: {world}-rank mesh-mode, G-buffer (fused / direct) and pos_mlp optimisations == 1-GPU runs")


if __name__ == "__main__":
    main()
