"""CPU, world_size 2 over gloo: the host-side sharding logic of an iteration (row shards, scalar all-reduce,
2-row halo exchange of the image gradient, gradient all-reduce) — SURVEY §8e."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from materialist_b200.parallel import ShardContext, shard_rows


def test_shard_rows_cover_image():
    for H, ws in ((2160, 8), (512, 3), (17, 4), (10, 1)):
        rows = [shard_rows(H, ws, r) for r in range(ws)]
        assert rows[0][0] == 0 and sum(n for _, n in rows) == H
        for (a, n), (b, _) in zip(rows, rows[1:]):
            assert a + n == b
        assert max(n for _, n in rows) - min(n for _, n in rows) <= 1


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, H, W, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sh = ShardContext(H, W, rank, world)
        full = torch.arange(H * W * 3, dtype=torch.float32).reshape(H, W, 3)
        mine = full[sh.row0:sh.row0 + sh.rows].clone()
        # 1. global scalar from per-shard sums
        s = sh.all_reduce_sum(mine.sum().reshape(1).double())
        # 2. halo exchange reproduces the neighbouring rows of the full image
        halo = sh.halo_exchange(mine)
        lo, hi = max(0, sh.row0 - 2), min(H, sh.row0 + sh.rows + 2)
        ok_halo = torch.equal(halo, full[lo:hi])
        buf, view = sh.halo_buffer(3, "cpu")
        view.copy_(mine)
        ok_halo = ok_halo and torch.equal(sh.halo_exchange_inplace(buf), full[lo:hi])
        # 3. gradient all-reduce
        g = sh.all_reduce_sum(torch.full((4,), float(rank + 1)))
        # 4. shared envmap, per-rank image rows: the envmap gradient every rank sees is the sum over ranks
        from materialist_b200.inverse import _SumGradOverRanks
        env = torch.ones(3, requires_grad=True)
        (_SumGradOverRanks.apply(env, sh) * float(rank + 1)).sum().backward()
        # 5. after the optimiser step: every rank has changed ITS rows of two full-image maps; the exchange refreshes the 2-row film
        #    halo with the owners' values and nothing else
        owner = torch.zeros(H, dtype=torch.long)
        for q in range(world):
            q0, qn = shard_rows(H, world, q); owner[q0:q0 + qn] = q
        truth = [(owner.float() + 1.0)[:, None, None] * torch.ones(H, W, c) * (10.0 ** i) for i, c in enumerate((3, 1))]
        maps = [torch.zeros(H, W, 3), torch.zeros(H, W, 1)]
        for mp_, t in zip(maps, truth):
            mp_[sh.row0:sh.row0 + sh.rows] = t[sh.row0:sh.row0 + sh.rows]
        sh.map_halo_exchange(maps)
        ok_maps = all(torch.equal(mp_[lo:hi], t[lo:hi]) for mp_, t in zip(maps, truth))
        untouched = all(float(mp_[:lo].abs().sum() + mp_[hi:].abs().sum()) == 0.0 for mp_ in maps)
        out[rank] = (float(s.item()), bool(ok_halo), g.tolist(), env.grad.tolist(), bool(ok_maps and untouched))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
@pytest.mark.parametrize("world", [2, 3])
def test_collectives_gloo(world):
    H, W = 11, 5
    mgr = mp.Manager(); out = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(world, port, H, W, out), nprocs=world, join=True)
    total = float(np.arange(H * W * 3, dtype=np.float64).sum())
    tri = float(world * (world + 1) // 2)
    for r in range(world):
        s, ok_halo, g, ge, ok_maps = out[r]
        assert s == total and ok_halo and g == [tri] * 4 and ge == [tri] * 3 and ok_maps
