"""Pixel-sharded data parallelism over the GPUs of one box: contiguous row blocks per rank, one
process per GPU, torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU tests) for the three exchange
steps of an iteration:

  1. scalar all-reduce of the image sums behind `ratio`, `loss_mse`, `loss_l1`  (inverse_img_w_mi.py:388-395)
  2. 2-row halo exchange of d(loss)/d(image): the gaussian film couples a pixel's samples to rows of the
     neighbouring shards, so the adjoint render needs their image gradients
  3. sum all-reduce of the envmap (and MLP) gradients — material-map gradients stay shard-local.
  4. after the optimiser step: the updated material maps of the boundary rows go to the neighbours (`map_halo_exchange`): a rank's
     forward render shades its own rows plus the 2-row film halo, i.e. it READS the neighbours' a / r / m there, and the neighbours
     have just stepped them.

RNG streams are indexed by the GLOBAL lane id, so any sharding reproduces the single-GPU samples exactly.

On GPUs the fused BRDF iteration does steps 1, 2 and 4 WITHOUT a collective library in the loop (`PeerArena`): every rank maps
the other ranks' peer-visible memory (CUDA IPC over NVLink / NVSwitch), the loss kernels exchange their scalar sums through
mailboxes from inside the kernels, and the halo rows are stored straight into the neighbours' buffers (include/materialist_b200.h,
"multi-GPU: exchange steps over peer memory").  NCCL remains for set-up, for the envmap / MLP gradient sums and as the fallback.
"""
import ctypes as C
import os

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

    # -- 4
    def map_halo_exchange(self, maps, halo=None):
        """In place: rows [row0 - halo, row0) and [row0 + rows, row0 + rows + halo) of every full-image contiguous (H, W, C) tensor in
        `maps` are overwritten with the owners' current values (the neighbours' boundary rows); this rank's own boundary rows go the
        other way.  Row blocks of a contiguous image are contiguous, so every transfer goes straight from / into the maps: ONE
        grouped NCCL launch, no packing kernels."""
        if self.world_size == 1:
            return
        h = self.halo if halo is None else halo
        if h == 0:
            return
        r0, r1, up, down = self.row0, self.row0 + self.rows, self.rank - 1, self.rank + 1
        ops = []
        for m in maps:
            if not m.is_contiguous():
                raise ValueError("map_halo_exchange needs contiguous (H, W, C) tensors")
            if up >= 0:
                ops += [dist.P2POp(dist.isend, m[r0:r0 + h], up, self.group), dist.P2POp(dist.irecv, m[r0 - h:r0], up, self.group)]
            if down < self.world_size:
                ops += [dist.P2POp(dist.isend, m[r1 - h:r1], down, self.group), dist.P2POp(dist.irecv, m[r1:r1 + h], down, self.group)]
        for req in dist.batch_isend_irecv(ops):
            req.wait()                       # NCCL: orders the current stream after the transfer, the host does not block

    # -- 2, without copies
    def halo_buffer(self, channels, device):
        """(rows + halos present, W, C) buffer + the view of this rank's own rows inside it: a kernel writes d(loss)/d(image) into the
        view, `halo_exchange_inplace` then fills the halo rows straight from the neighbours."""
        top = self.halo if (self.world_size > 1 and self.rank > 0) else 0
        bot = self.halo if (self.world_size > 1 and self.rank < self.world_size - 1) else 0
        full = torch.zeros(top + self.rows + bot, self.W, channels, device=device)
        return full, full[top:top + self.rows]

    def halo_exchange_inplace(self, full):
        if self.world_size == 1 or self.halo == 0:
            return full
        h, up, down = self.halo, self.rank - 1, self.rank + 1
        top = h if up >= 0 else 0
        ops = []
        if up >= 0:
            ops += [dist.P2POp(dist.isend, full[top:top + h], up, self.group), dist.P2POp(dist.irecv, full[:h], up, self.group)]
        if down < self.world_size:
            n = full.shape[0]
            ops += [dist.P2POp(dist.isend, full[n - 2 * h:n - h], down, self.group), dist.P2POp(dist.irecv, full[n - h:], down, self.group)]
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        return full


class _DevMem:
    """A raw device allocation as a CUDA-array-interface object (torch.as_tensor wraps it without copying)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


class PeerArena:
    """Peer-visible device memory of this rank (mailbox + named buffers) and the mappings of every other rank's arena.

    Collective: all ranks of the shard's group construct it together, `alloc` the same names, then `finalize()` (which exchanges the
    IPC handles and the offset tables through torch.distributed — set-up only).  `tensor(name)` is this rank's buffer;
    `remote_ptr(rank, name, byte_offset)` is where that buffer of another rank is mapped in this process."""

    def __init__(self, shard, device):
        from . import _abi
        if shard.world_size < 2 or shard.world_size > _abi.MAX_PEERS:
            raise ValueError("PeerArena needs 2..%d ranks" % _abi.MAX_PEERS)
        self._abi, self.shard, self.device = _abi, shard, torch.device(device)
        self.box_bytes = int(_abi.lib.mb200_peer_box_bytes())
        self._plan, self._cursor = {}, self.box_bytes
        self.base, self._own, self._mem, self._tensors = None, None, None, {}

    def alloc(self, name, shape, dtype=torch.float32):
        n = 1
        for d in shape:
            n *= int(d)
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        self._plan[name] = (self._cursor, tuple(int(d) for d in shape), dtype, nbytes)
        self._cursor += (nbytes + 255) & ~255

    def finalize(self):
        _abi, sh = self._abi, self.shard
        with torch.cuda.device(self.device):
            ptr = C.c_void_p(); handle = (C.c_ubyte * 64)()
            _abi.check(_abi.lib.mb200_peer_alloc(self._cursor, C.byref(ptr), handle), "mb200_peer_alloc")
            self._own = ptr.value
            mine = (bytes(handle), {k: v[0] for k, v in self._plan.items()})
            everyone = [None] * sh.world_size
            dist.all_gather_object(everyone, mine, group=sh.group)
            self.base, self.offsets = [None] * sh.world_size, [e[1] for e in everyone]
            for r, (h, _) in enumerate(everyone):
                if r == sh.rank:
                    self.base[r] = self._own
                else:
                    q = C.c_void_p(); hb = (C.c_ubyte * 64).from_buffer_copy(h)
                    _abi.check(_abi.lib.mb200_peer_open(hb, C.byref(q)), "mb200_peer_open")
                    self.base[r] = q.value
            self._mem = torch.as_tensor(_DevMem(self._own, self._cursor), device=self.device)
            for name, (off, shape, dtype, nbytes) in self._plan.items():
                self._tensors[name] = self._mem[off:off + nbytes].view(dtype).view(shape)
            dist.barrier(group=sh.group)          # every mailbox is zeroed and mapped before anyone's first kernel writes into it
        return self

    def tensor(self, name):
        return self._tensors[name]

    def remote_ptr(self, rank, name, byte_offset=0):
        return self.base[rank] + self.offsets[rank][name] + int(byte_offset)

    def peer(self, seq):
        p = self._abi.Peer()
        p.rank, p.world, p.seq = self.shard.rank, self.shard.world_size, int(seq)
        for r in range(self.shard.world_size):
            p.box[r] = self.base[r]
        return p

    def close(self):
        if self.base is None:
            return
        lib = self._abi.lib
        try:
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self.shard.group)   # nobody unmaps / frees while a peer's kernel may still write here
        except Exception:
            pass
        for r, b in enumerate(self.base):
            if r != self.shard.rank and b:
                lib.mb200_peer_close(C.c_void_p(b))
        self._tensors, self._mem = {}, None
        lib.mb200_peer_free(C.c_void_p(self._own))
        self.base = None


def peer_exchange_available(shard, device):
    """The peer-memory exchange needs one GPU per rank on one node with peer access (NVLink / NVSwitch) and the NCCL-era set-up
    channel; MB200_PEER=0 forces the NCCL path."""
    if os.environ.get("MB200_PEER", "1") == "0" or shard.world_size < 2 or shard.world_size > 16:
        return False
    if not (dist.is_available() and dist.is_initialized()) or dist.get_backend(shard.group) != "nccl":
        return False
    dev = torch.device(device)
    if dev.type != "cuda" or torch.cuda.device_count() < shard.world_size:
        return False
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    return all(p == idx or torch.cuda.can_device_access_peer(idx, p) for p in range(shard.world_size))
