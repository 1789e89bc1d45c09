"""Pixel-sharded data parallelism over the GPUs of one box: contiguous row blocks per rank, one
process per GPU, torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU tests) for the three exchange
steps of an iteration:

  1. scalar all-reduce of the image sums behind `ratio`, `loss_mse`, `loss_l1`  (inverse_img_w_mi.py:388-395)
  2. 2-row halo exchange of d(loss)/d(image): the gaussian film couples a pixel's samples to rows of the
     neighbouring shards, so the adjoint render needs their image gradients
  3. sum all-reduce of the envmap (and MLP) gradients — material-map gradients stay shard-local.

RNG streams are indexed by the GLOBAL lane id, so any sharding reproduces the single-GPU samples exactly.
"""
import torch
import torch.distributed as dist

FILM_HALO = 2


def shard_rows(H, world_size, rank):
    """Contiguous row block [row0, row0+rows) of rank `rank`; remainders go to the first ranks."""
    base, rem = divmod(H, world_size)
    rows = base + (1 if rank < rem else 0)
    row0 = rank * base + min(rank, rem)
    return row0, rows


class ShardContext:
    """Row shard of one rank + the collectives of an iteration. world_size == 1 degenerates to no-ops."""

    def __init__(self, H, W, rank=0, world_size=1, group=None, halo=FILM_HALO):
        self.H, self.W, self.rank, self.world_size, self.group = H, W, rank, world_size, group
        self.row0, self.rows = shard_rows(H, world_size, rank)
        self.halo = halo
        if world_size > 1 and min(shard_rows(H, world_size, r)[1] for r in range(world_size)) < halo:
            raise ValueError("every shard needs at least `halo` rows")

    # -- 1 / 3
    def all_reduce_sum(self, t):
        if self.world_size > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    # -- 2
    def halo_exchange(self, grad_rows):
        """(rows, W, C) -> (rows + halo above + halo below, W, C), halos filled with the neighbours' boundary rows
        (absent at the image border)."""
        if self.world_size == 1 or self.halo == 0:
            return grad_rows
        h, up, down = self.halo, self.rank - 1, self.rank + 1
        grad_rows = grad_rows.contiguous()
        ops, recv_up, recv_down = [], None, None
        if up >= 0:
            recv_up = torch.empty_like(grad_rows[:h])
            ops += [dist.P2POp(dist.isend, grad_rows[:h].contiguous(), up, self.group),
                    dist.P2POp(dist.irecv, recv_up, up, self.group)]
        if down < self.world_size:
            recv_down = torch.empty_like(grad_rows[:h])
            ops += [dist.P2POp(dist.isend, grad_rows[-h:].contiguous(), down, self.group),
                    dist.P2POp(dist.irecv, recv_down, down, self.group)]
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        parts = ([recv_up] if recv_up is not None else []) + [grad_rows] + ([recv_down] if recv_down is not None else [])
        return torch.cat(parts, 0)
