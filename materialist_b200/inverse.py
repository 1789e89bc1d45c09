"""Host-side optimisation iteration around the render operator — the body of the BRDF phase of
`optimize_envmap_ARMN` with `model_name == 'none'` (inverse_img_w_mi.py:343-446) and of the envmap phase
(:237-256), minus file I/O, tqdm and the per-iteration `.item()` host syncs.  Everything stays on the device;
with a ShardContext the three scalar sums and the image-gradient halo are exchanged between ranks.
"""
import ctypes as C
import os

import torch
import torch.nn.functional as NF

from . import _abi
from . import renderop as _rop
from .parallel import PeerArena, ShardContext, peer_exchange_available, shard_rows
from .renderop import render


def linear_to_srgb(image):
    """myutils/misc.py:167-170"""
    return image ** (1.0 / 2.2)


def brdf_phase_lr(k, lr0=3e-4, step_size=100, gamma=0.8, floor=1.5e-4):
    """Learning rate of iteration k (0-based) of the BRDF phase: StepLR(step_size=100, gamma=0.8) that the reference
    only advances while `current_lr > 1.5e-4` (inverse_img_w_mi.py:359, :424-432) — a pure function of k, so the
    host never reads the device."""
    lr, epoch = lr0, 0
    for _ in range(k):
        if lr > floor:
            epoch += 1
            if epoch % step_size == 0:
                lr *= gamma
    return lr


class _ShardedStep:
    """`step(seed)` renders this rank's rows and leaves the Scene's shard as it found it (a Scene is shared between phases and with
    plain `render(scene)` calls, which must keep seeing the whole image)."""

    def step(self, seed):
        with self.scene.shard(self.shard.row0, self.shard.rows):
            return self._step(seed)


class DirectBRDFOptimizer(_ShardedStep):
    """`Directly optimizing {a,r,m} without neural network` (inverse_img_w_mi.py:346-446).

    mat: dict of full-image CUDA tensors albedo (H,W,3), roughness (H,W,1), metallic (H,W,1);
    gt_image: (H,W,3) linear radiance.  Each rank renders and optimises its own row shard.
    """

    def __init__(self, scene, mat, gt_image, optimize_part="arm", spp=64, lr=3e-4, scale_delta=0.1, shard=None):
        self.scene, self.spp, self.scale_delta, self.part = scene, spp, scale_delta, optimize_part
        self.shard = shard or ShardContext(scene.H, scene.W)
        self.mat = {k: v.detach().clone() for k, v in mat.items()}
        self.ori = {k: v.detach().clone() for k, v in mat.items()}
        self.params = {}
        if "a" in optimize_part: self.params["albedo"] = torch.nn.Parameter(mat["albedo"].clone())
        if "r" in optimize_part: self.params["roughness"] = torch.nn.Parameter(mat["roughness"].clone())
        if "m" in optimize_part: self.params["metallic"] = torch.nn.Parameter(mat["metallic"].clone())
        self.normal_ori = None
        if not scene.use_mesh_normal:                          # :356-357: the normal map is a parameter only with 'n' in the part
            if "normal" not in mat:
                raise ValueError("use_mesh_normal=False needs mat['normal'] (H, W, 3)")
            self.normal_ori = NF.normalize(mat["normal"].detach(), p=2, dim=-1)            # :196
            if "n" in optimize_part:
                if scene.mesh is not None and self.shard.world_size > 1:
                    raise NotImplementedError("'n' with a traced scene under sharding is not supported")
                self.params["normal"] = torch.nn.Parameter(mat["normal"].clone())
        self.opt = torch.optim.Adam(self.params.values(), lr=lr)
        self.sched = torch.optim.lr_scheduler.StepLR(self.opt, step_size=100, gamma=0.8)
        r0, r1 = self.shard.row0, self.shard.row0 + self.shard.rows
        self.rows = slice(r0, r1)
        self.gt = gt_image[self.rows].contiguous()
        self.gt_srgb = linear_to_srgb(self.gt)
        self.n_img = float(scene.H * scene.W * 3)
        self.gt_sum = self.shard.all_reduce_sum(self.gt.sum().reshape(1).clone())
        self.last = {}

    def _step(self, seed):
        p, sh, rows = self.params, self.shard, self.rows
        mat = dict(self.mat)
        if "albedo" in p: mat["albedo"] = p["albedo"].clamp(0, 1)
        if "roughness" in p: mat["roughness"] = p["roughness"].clamp(0.07, 1)
        if "metallic" in p: mat["metallic"] = p["metallic"].clamp(0, 1)
        normal = None
        if not self.scene.use_mesh_normal:                     # :375-376, :384
            normal = NF.normalize(p["normal"], p=2, dim=-1) if "normal" in p else mat["normal"]
        pred = render(self.scene, spp=self.spp, seed=seed, albedo=mat["albedo"], roughness=mat["roughness"],
                      metallic=mat["metallic"], normal=normal, halo_exchange=sh.halo_exchange if sh.world_size > 1 else None)
        # ratio = gt.mean() / pred.detach().mean()  — a GLOBAL scalar over all pixels (:388-389)
        pred_sum = sh.all_reduce_sum(pred.detach().sum().reshape(1))
        pred = pred * (self.gt_sum / pred_sum)
        pred_srgb = linear_to_srgb(pred)
        diff = pred_srgb - self.gt_srgb
        sums = sh.all_reduce_sum(torch.stack([(diff * diff).sum().detach(), diff.abs().sum().detach()]))
        loss_mse_l = (diff * diff).sum() / self.n_img          # this rank's share of the global means
        loss_l1_l = diff.abs().sum() / self.n_img
        npx = float(self.scene.H * self.scene.W)
        aux = 0.0
        if "albedo" in p: aux = aux + (mat["albedo"][rows] - self.ori["albedo"][rows]).abs().sum() / (npx * 3)
        if "roughness" in p: aux = aux + (mat["roughness"][rows] - self.ori["roughness"][rows]).abs().sum() / npx
        if "metallic" in p: aux = aux + (mat["metallic"][rows] - self.ori["metallic"][rows]).abs().sum() / npx
        if "normal" in p: aux = aux + (normal[rows] - self.normal_ori[rows]).abs().sum() / (npx * 3)      # :407-408
        scale_ratio = sums[1] / sums[0]                        # loss_l1.detach() / loss_mse.detach()
        loss = 3 * scale_ratio * loss_mse_l + loss_l1_l + aux * self.scale_delta
        loss.backward()
        if sh.world_size > 1 and self.scene.mesh is not None:
            # traced scene: a path that starts in this rank's rows reads the maps — and scatters their gradients — at whatever texels
            # its secondary vertices hit, in any shard: sum the map gradients over the ranks, every rank steps the whole image
            for q in p.values():
                sh.all_reduce_sum(q.grad)
        self.opt.step()
        self.opt.zero_grad(set_to_none=True)
        if sh.world_size > 1 and self.scene.mesh is None:
            # G-buffer mode: only this rank's rows were stepped; its forward also READS the 2-row film halo of the neighbours'
            # maps, which they have just stepped -> fetch them (parameters, so that the clamps above apply next iteration)
            with torch.no_grad():
                sh.map_halo_exchange([q.data for q in p.values()])
        if self.opt.param_groups[0]["lr"] > 1.5e-4:            # :431-432 (a host-side float, no device read)
            self.sched.step()
        self.last = {"loss_mse": sums[0] / self.n_img, "loss_l1": sums[1] / self.n_img, "pred": pred_srgb}
        return loss.detach()


class FusedBRDFOptimizer(_ShardedStep):
    """Same iteration as DirectBRDFOptimizer (inverse_img_w_mi.py:346-446, `model_name == 'none'`), with everything
    between the two render kernels done by the fused loss / Adam kernels of csrc/mb200_optim.cu instead of ~100
    elementwise torch launches + autograd: 9 kernel launches per iteration, no host sync, scalars stay on the device.

    Per iteration: shade_fwd -> film_develop -> image_sum [all-reduce] -> loss_srgb_sums [all-reduce] ->
    loss_srgb_grad [halo exchange of d loss / d image] -> film_adjoint -> shade_bwd -> adam_clamped [halo exchange of the stepped
    maps]; film_weights (a function of seed_grad alone) runs on a side stream from the end of shade_fwd, under the loss kernels
    and the collectives.
    """

    _RANGE = {"albedo": (0.0, 1.0), "roughness": (0.07, 1.0), "metallic": (0.0, 1.0)}

    def __init__(self, scene, mat, gt_image, optimize_part="arm", spp=64, lr=3e-4, scale_delta=0.1, shard=None,
                 betas=(0.9, 0.999), eps=1e-8):
        self.scene, self.spp, self.scale_delta, self.part = scene, spp, scale_delta, optimize_part
        self.lr0, self.betas, self.eps = lr, betas, eps
        self.shard = sh = shard or ShardContext(scene.H, scene.W)
        dev, H, W = scene.device, scene.H, scene.W
        self.names = [n for n, k in (("albedo", "a"), ("roughness", "r"), ("metallic", "m")) if k in optimize_part]
        # Multi-GPU, G-buffer mode: the exchange steps of the iteration go over peer memory from inside / between the kernels
        # (parallel.PeerArena) instead of through NCCL; the buffers the neighbours write into — the rendered-with maps and the
        # d(loss)/d(image) buffer with its film halo — then live in this rank's peer-visible arena.
        ch = {"albedo": 3, "roughness": 1, "metallic": 1}
        self.peer = None
        if (sh.world_size > 1 and scene.mesh is None and W % 4 == 0 and scene.filter == _abi.FILTER_GAUSSIAN
                and scene.use_mesh_normal and peer_exchange_available(sh, dev)):
            top = sh.halo if sh.rank > 0 else 0
            bot = sh.halo if sh.rank < sh.world_size - 1 else 0
            ar = PeerArena(sh, dev)
            for k in ("albedo", "roughness", "metallic"):
                ar.alloc(k, (H, W, ch[k]))
            ar.alloc("grad_full", (top + sh.rows + bot, W, 3))
            ar.alloc("ticket", (64,), torch.int32)
            self.peer = ar.finalize()
        # the maps the kernels render with (full image; this rank only ever touches its own rows)
        if self.peer is not None:
            self.mat = {k: self.peer.tensor(k) for k in ("albedo", "roughness", "metallic")}
            for k in self.mat:
                self.mat[k].copy_(mat[k].detach().float())
        else:
            self.mat = {k: mat[k].detach().float().clone().contiguous() for k in ("albedo", "roughness", "metallic")}
        for k in self.names:
            self.mat[k].clamp_(*self._RANGE[k])
        self.ori = {k: mat[k].detach().float().clone().contiguous() for k in self.names}
        self.params = {k: mat[k].detach().float().clone().contiguous() for k in self.names}
        # one flat gradient buffer (one memset per iteration) and one flat Adam state
        sizes = [H * W * ch[k] for k in ("albedo", "roughness", "metallic")]
        self.gflat = torch.zeros(sum(sizes), device=dev)
        ga, gr, gm = torch.split(self.gflat, sizes)
        self.grads = {"albedo": ga.view(H, W, 3), "roughness": gr.view(H, W, 1), "metallic": gm.view(H, W, 1)}
        self.exp_avg = {k: torch.zeros_like(self.params[k]) for k in self.names}
        self.exp_avg_sq = {k: torch.zeros_like(self.params[k]) for k in self.names}
        self.rows = slice(sh.row0, sh.row0 + sh.rows)
        self.gt = gt_image[self.rows].float().contiguous()
        self.gt_srgb = linear_to_srgb(self.gt)
        self.n_total = H * W * 3
        # device scalars: scal = (Σ gt, Σ pred), sums2 = (Σ diff², Σ |diff|)
        self.scal = torch.zeros(2, device=dev)
        self.scal[0:1] = sh.all_reduce_sum(self.gt.sum().reshape(1).clone())
        self.sums2 = torch.zeros(2, device=dev)
        self.scratch = torch.zeros(_abi.lib.mb200_reduce_scratch_bytes() // 4 + 1, dtype=torch.int32, device=dev)
        if self.peer is not None:                                      # d loss / d image: own rows (a view) inside the rows + film-halo buffer
            self.grad_full = self.peer.tensor("grad_full")
            self.grad_img = self.grad_full[top:top + sh.rows]
            self._peer_plan(top)
        else:
            self.grad_full, self.grad_img = sh.halo_buffer(3, dev)
        self.pred_srgb = torch.empty(sh.rows, W, 3, device=dev)
        self.k, self._lr, self._epoch = 0, lr, 0
        self.last = {}
        self._side, self._wpart = None, None                  # side stream + film-weight buffer of the adjoint render (see _step)
        # Normal map (use_mesh_normal False, inverse_img_w_mi.py:356-357, :375-376, :384, :407-410): rendered with mat['normal']; with
        # 'n' in optimize_part the PARAMETER is the un-normalised map, the kernels see normalize(p), and the aux term is
        # l1(normalize(p), normal_ori).  The a / r / m part stays in the fused kernels; the normal's chain rule through the
        # normalisation and its Adam step are a handful of elementwise torch ops on one (H, W, 3) tensor.
        self.normal = self.normal_ori = self.n_param = self.n_opt = None
        if not scene.use_mesh_normal:
            if "normal" not in mat:
                raise ValueError("use_mesh_normal=False needs mat['normal'] (H, W, 3)")
            self.normal = mat["normal"].detach().float().clone().contiguous()
            if "n" in optimize_part:
                if scene.mesh is not None and sh.world_size > 1:
                    raise NotImplementedError("'n' with a traced scene under sharding: use DirectBRDFOptimizer")
                self.normal_ori = NF.normalize(self.normal, p=2, dim=-1)
                self.n_param = torch.nn.Parameter(self.normal.clone())
                self.n_opt = torch.optim.Adam([self.n_param], lr=lr, betas=betas, eps=eps)
                self.normal = NF.normalize(self.n_param.detach(), p=2, dim=-1).contiguous()
        # Adam segments over this rank's rows (contiguous in the row-major maps).  Mesh mode under sharding: a path that starts in
        # this rank's rows scatters material gradients to whatever texels its secondary vertices hit, and reads the maps there —
        # so the map gradients are summed over the ranks and every rank steps the WHOLE image (replicated state, 5 floats / pixel).
        self.replicated = scene.mesh is not None and sh.world_size > 1
        segs = (_abi.AdamSeg * len(self.names))()
        npx = float(H * W)
        for i, k in enumerate(self.names):
            c = ch[k]
            off = 0 if self.replicated else sh.row0 * W * c * 4
            n = (H if self.replicated else sh.rows) * W * c
            segs[i].p = self.params[k].data_ptr() + off; segs[i].mat = self.mat[k].data_ptr() + off
            segs[i].g = self.grads[k].data_ptr() + off; segs[i].ori = self.ori[k].data_ptr() + off
            segs[i].m = self.exp_avg[k].data_ptr() + off; segs[i].v = self.exp_avg_sq[k].data_ptr() + off
            segs[i].n = n; segs[i].lo, segs[i].hi = self._RANGE[k]
            segs[i].aux_coeff = scale_delta / (npx * c)
        self.segs = segs

    def _peer_plan(self, top):
        """Push descriptors of the two halo exchanges (constant over the optimisation): where this rank's boundary rows land in the
        neighbours' arenas."""
        sh, ar, W, h = self.shard, self.peer, self.scene.W, self.shard.halo
        up = sh.rank - 1 if sh.rank > 0 else -1
        down = sh.rank + 1 if sh.rank < sh.world_size - 1 else -1
        self._nb = (up, down)
        row_b = W * 3 * 4
        segs = []
        if up >= 0:       # my first h rows -> the upper neighbour's bottom halo
            up_top = h if up > 0 else 0
            up_rows = shard_rows(sh.H, sh.world_size, up)[1]
            segs.append((self.grad_full[top:top + h].data_ptr(), ar.remote_ptr(up, "grad_full", (up_top + up_rows) * row_b), h * row_b // 16))
        if down >= 0:     # my last h rows -> the lower neighbour's top halo
            segs.append((self.grad_full[top + sh.rows - h:top + sh.rows].data_ptr(), ar.remote_ptr(down, "grad_full", 0), h * row_b // 16))
        self._push_halo = (_abi.PushSeg * max(1, len(segs)))(*[_abi.PushSeg(a, b, n) for a, b, n in segs]); self._n_push_halo = len(segs)
        segs = []
        chn = {"albedo": 3, "roughness": 1, "metallic": 1}
        r0, r1 = sh.row0, sh.row0 + sh.rows
        for k in self.names:
            rb = W * chn[k] * 4
            if up >= 0:
                segs.append((self.mat[k][r0:r0 + h].data_ptr(), ar.remote_ptr(up, k, r0 * rb), h * rb // 16))
            if down >= 0:
                segs.append((self.mat[k][r1 - h:r1].data_ptr(), ar.remote_ptr(down, k, (r1 - h) * rb), h * rb // 16))
        self._push_map = (_abi.PushSeg * max(1, len(segs)))(*[_abi.PushSeg(a, b, n) for a, b, n in segs]); self._n_push_map = len(segs)
        self._ticket = self.peer.tensor("ticket")

    def _launch_film_weights(self, main, seed_grad, env_pack):
        """The film weights of the adjoint render depend on nothing but seed_grad: they run on a side stream.  MB200_FILM_WEIGHTS_EARLY
        selects where: "2" (default) launches them BEHIND the forward render with no dependency on it — the forward's CTAs are
        dispatched first and leave no room for the weights' CTAs, which then fill the forward's tail and the idle slots under the
        latency-bound loss kernels; "1" launches them before the forward render (co-resident from its start); "0" after it (a
        dependency on the forward render: its per-kernel time in the bench stays clean).  Measured at C2 (profiles/r6h_prio_probe.log):
        2.500 / 2.468 / 2.464 ms per iteration for "0" / "1" / "2"; a high-priority stream for everything else: within noise of that."""
        sc = self.scene
        if sc.filter != _abi.FILTER_GAUSSIAN:
            return
        if self._side is None:
            self._side = torch.cuda.Stream(sc.device)
            self._ev_fwd, self._ev_w = torch.cuda.Event(), torch.cuda.Event()
            self._fw_mode = os.environ.get("MB200_FILM_WEIGHTS_EARLY", "2")
            self._fw_late = self._fw_mode != "1"
        if self._fw_late:
            self._pending_fw = (seed_grad, env_pack)
            if self._fw_mode == "2":
                self._ev_fwd.record(main)         # dependency: the previous iteration; the launch itself follows the forward render's
            return
        self._ev_fwd.record(main)                 # after the previous iteration's adjoint render (which read the weight buffer)
        with torch.cuda.stream(self._side):
            self._side.wait_event(self._ev_fwd)
            self._wpart = _rop._film_weights(sc, self.spp, seed_grad, env_pack[2].res_x, out=self._wpart)
            self._ev_w.record(self._side)

    def _late_film_weights(self, main):
        if self.scene.filter == _abi.FILTER_GAUSSIAN and self._fw_late:
            seed_grad, env_pack = self._pending_fw
            if self._fw_mode != "2":
                self._ev_fwd.record(main)
            with torch.cuda.stream(self._side):
                self._side.wait_event(self._ev_fwd)
                self._wpart = _rop._film_weights(self.scene, self.spp, seed_grad, env_pack[2].res_x, out=self._wpart)
                self._ev_w.record(self._side)

    def close(self):
        """Unmaps / frees the peer arena (collective; call on every rank when the optimisation is over)."""
        if self.peer is not None:
            keep = {k: v.clone() for k, v in self.mat.items()}
            self.peer.close(); self.peer = None
            self.mat = keep
            # everything that lived in the arena moves to ordinary device memory, so that the optimiser stays usable (NCCL path)
            self.grad_full, self.grad_img = self.shard.halo_buffer(3, self.scene.device)
            for i, k in enumerate(self.names):
                c = {"albedo": 3, "roughness": 1, "metallic": 1}[k]
                off = self.shard.row0 * self.scene.W * c * 4
                self.segs[i].mat = self.mat[k].data_ptr() + off

    def _step(self, seed):
        if self.peer is not None:
            return self._step_peer(seed)
        sc, sh, lib, st = self.scene, self.shard, _abi.lib, _abi.stream_ptr()
        a, r, m = self.mat["albedo"], self.mat["roughness"], self.mat["metallic"]
        env_pack = sc.prepared_env()
        seed_grad = _rop.default_seed_grad(int(seed))
        nmap = self.normal
        main = torch.cuda.current_stream(sc.device)
        self._launch_film_weights(main, seed_grad, env_pack)
        img = _rop._forward(sc, self.spp, int(seed), a, r, m, nmap, env_pack)
        self._late_film_weights(main)
        n = img.numel()
        _abi.check(lib.mb200_image_sum(_abi.ptr(img), n, C.c_void_p(self.scal.data_ptr() + 4), _abi.ptr(self.scratch), st), "mb200_image_sum")
        if sh.world_size > 1:
            sh.all_reduce_sum(self.scal[1:2])
        _abi.check(lib.mb200_loss_srgb_sums(_abi.ptr(img), _abi.ptr(self.gt_srgb), n, _abi.ptr(self.scal), _abi.ptr(self.sums2),
                                            _abi.ptr(self.pred_srgb), _abi.ptr(self.scratch), st), "mb200_loss_srgb_sums")
        if sh.world_size > 1:
            sh.all_reduce_sum(self.sums2)
        _abi.check(lib.mb200_loss_srgb_grad(_abi.ptr(img), _abi.ptr(self.gt_srgb), n, _abi.ptr(self.scal), _abi.ptr(self.sums2),
                                            self.n_total, _abi.ptr(self.grad_img), st), "mb200_loss_srgb_grad")
        grad = sh.halo_exchange_inplace(self.grad_full) if sh.world_size > 1 else self.grad_img
        self.gflat.zero_()
        if sc.filter == _abi.FILTER_GAUSSIAN:
            main.wait_event(self._ev_w)
        g = _rop._backward(sc, self.spp, seed_grad, a, r, m, nmap, env_pack, grad,
                           "albedo" in self.names, "roughness" in self.names, "metallic" in self.names, self.n_param is not None, False,
                           out=(self.grads["albedo"], self.grads["roughness"], self.grads["metallic"]), wpart=self._wpart)
        if self.replicated:
            sh.all_reduce_sum(self.gflat)
        if self.n_param is not None:
            # d loss / d p through n = p / |p|, with the aux term scale_delta * l1(n, n_ori) (mean over H*W*3) added on n first
            rows = self.rows
            gn = g[3][rows] + (self.scale_delta / float(sc.H * sc.W * 3)) * torch.sign(self.normal[rows] - self.normal_ori[rows])
            p = self.n_param.data[rows]
            inv = 1.0 / p.norm(dim=-1, keepdim=True).clamp_min(1e-12)
            nn_ = p * inv
            gp = torch.zeros_like(self.n_param.data)
            gp[rows] = (gn - nn_ * (nn_ * gn).sum(-1, keepdim=True)) * inv
            self.n_param.grad = gp
        self.k += 1
        lr = self._lr
        if self._lr > 1.5e-4:                                  # StepLR(100, 0.8), advanced only above the floor (:431-432)
            self._epoch += 1
            if self._epoch % 100 == 0:
                self._lr *= 0.8
        _abi.check(lib.mb200_adam_clamped(self.segs, len(self.names), lr, self.betas[0], self.betas[1], self.eps, self.k, st),
                   "mb200_adam_clamped")
        if self.n_param is not None:
            for gr_ in self.n_opt.param_groups:
                gr_["lr"] = lr
            self.n_opt.step()
            self.normal = NF.normalize(self.n_param.detach(), p=2, dim=-1).contiguous()
        if sh.world_size > 1 and not self.replicated:
            # only this rank's rows were stepped; its next forward READS the neighbours' maps in the 2-row film halo, and they have
            # just stepped those rows: fetch them (without this the sharded run drifts from the single-GPU run after iteration 1)
            sh.map_halo_exchange([self.mat[k] for k in self.names] + ([self.normal] if self.n_param is not None else []))
        self.last = {"loss_mse": self.sums2[0] / self.n_total, "loss_l1": self.sums2[1] / self.n_total, "pred": self.pred_srgb}
        return self.last["loss_mse"]


    def _step_peer(self, seed):
        """The iteration with its exchange steps over peer memory (see __init__): shade_fwd -> film_develop -> image_sum [publishes
        Σ pred] -> loss_srgb_sums [collects Σ pred; publishes Σ diff², Σ |diff|] -> loss_srgb_grad [collects] -> push of the film halo of
        d loss / d image into the neighbours -> wait for theirs -> film_adjoint -> shade_bwd -> adam_clamped -> push of the stepped
        boundary rows of the maps into the neighbours (awaited at the start of the next iteration).  No NCCL call."""
        sc, sh, lib, st, ar = self.scene, self.shard, _abi.lib, _abi.stream_ptr(), self.peer
        a, r, m = self.mat["albedo"], self.mat["roughness"], self.mat["metallic"]
        env_pack = sc.prepared_env()
        seed_grad = _rop.default_seed_grad(int(seed))
        seq = self.k + 1
        up, down = self._nb
        pr = ar.peer(seq)
        if seq > 1:       # the neighbours' stepped boundary rows of the previous iteration have landed in my maps
            prev = ar.peer(seq - 1)
            _abi.check(lib.mb200_peer_wait(C.byref(prev), _abi.PEER_MAP, up, down, st), "mb200_peer_wait")
        main = torch.cuda.current_stream(sc.device)
        self._launch_film_weights(main, seed_grad, env_pack)
        img = _rop._forward(sc, self.spp, int(seed), a, r, m, None, env_pack)
        self._late_film_weights(main)
        n = img.numel()
        _abi.check(lib.mb200_image_sum_peer(_abi.ptr(img), n, C.c_void_p(self.scal.data_ptr() + 4), _abi.ptr(self.scratch), C.byref(pr), st),
                   "mb200_image_sum_peer")
        _abi.check(lib.mb200_loss_srgb_sums_peer(_abi.ptr(img), _abi.ptr(self.gt_srgb), n, _abi.ptr(self.scal), _abi.ptr(self.sums2),
                                                 _abi.ptr(self.pred_srgb), _abi.ptr(self.scratch), C.byref(pr), st), "mb200_loss_srgb_sums_peer")
        _abi.check(lib.mb200_loss_srgb_grad_peer(_abi.ptr(img), _abi.ptr(self.gt_srgb), n, _abi.ptr(self.scal), _abi.ptr(self.sums2),
                                                 self.n_total, _abi.ptr(self.grad_img), C.byref(pr), st), "mb200_loss_srgb_grad_peer")
        _abi.check(lib.mb200_peer_push(C.byref(pr), _abi.PEER_HALO, self._push_halo, self._n_push_halo, up, down, _abi.ptr(self._ticket), st),
                   "mb200_peer_push")
        self.gflat.zero_()
        main.wait_event(self._ev_w)
        _abi.check(lib.mb200_peer_wait(C.byref(pr), _abi.PEER_HALO, up, down, st), "mb200_peer_wait")
        _rop._backward(sc, self.spp, seed_grad, a, r, m, None, env_pack, self.grad_full,
                       "albedo" in self.names, "roughness" in self.names, "metallic" in self.names, False, False,
                       out=(self.grads["albedo"], self.grads["roughness"], self.grads["metallic"]), wpart=self._wpart)
        self.k += 1
        lr = self._lr
        if self._lr > 1.5e-4:
            self._epoch += 1
            if self._epoch % 100 == 0:
                self._lr *= 0.8
        _abi.check(lib.mb200_adam_clamped(self.segs, len(self.names), lr, self.betas[0], self.betas[1], self.eps, self.k, st), "mb200_adam_clamped")
        _abi.check(lib.mb200_peer_push(C.byref(pr), _abi.PEER_MAP, self._push_map, self._n_push_map, up, down,
                                       C.c_void_p(self._ticket.data_ptr() + 64), st), "mb200_peer_push")
        self.last = {"loss_mse": self.sums2[0] / self.n_total, "loss_l1": self.sums2[1] / self.n_total, "pred": self.pred_srgb}
        return self.last["loss_mse"]


class PosMLPBRDFOptimizer(_ShardedStep):
    """BRDF phase with `model_name == 'pos_mlp'` (inverse_img_w_mi.py:471-552): the maps are the output of `brdf_net`
    (PosMLP on the tensor cores, mymodels/mlps.py) applied to the initial estimate `start_arm`; the network weights are
    optimised with AdamW(lr=3e-4) + the guarded StepLR(100, 0.8).

    Row sharding (`shard`): each rank evaluates the network only on the rows it renders — its own rows plus the 2-row film halo
    (the whole image in mesh mode, where paths read the maps anywhere) — through PosMLP's `row0`; the three image sums of the
    loss are all-reduced as in FusedBRDFOptimizer, and the 198 662 weight gradients are summed over the ranks before the
    (replicated) AdamW step."""

    def __init__(self, scene, mat, gt_image, optimize_part="arm", spp=64, lr=3e-4, scale_delta=0.1, net=None, shard=None):
        from .mymodels.mlps import PosMLP
        self.scene, self.spp, self.scale_delta, self.part = scene, spp, scale_delta, optimize_part
        self.shard = sh = shard or ShardContext(scene.H, scene.W)
        dev, H, W = scene.device, scene.H, scene.W
        self.net = net if net is not None else PosMLP(in_dims=7, out_dims=5, dims=[256] * 4, skip_connection=[1, 3], weight_norm=False,
                                                      multires_view=2, output_type="arm", color_ch=5).to(dev)   # :163
        self.mat = {k: v.detach().clone() for k, v in mat.items()}
        self.ori = {k: v.detach().clone() for k, v in mat.items()}
        self.start_arm = torch.cat([mat["albedo"].reshape(-1, 3), mat["roughness"].reshape(-1, 1), mat["metallic"].reshape(-1, 1)],
                                   dim=-1).clamp(0, 1).contiguous()                                               # :205
        if scene.mesh is not None or sh.world_size == 1:
            self.er0, self.er1 = 0, H
        else:
            self.er0, self.er1 = max(0, sh.row0 - sh.halo), min(H, sh.row0 + sh.rows + sh.halo)
        self.opt = torch.optim.AdamW(self.net.parameters(), lr=lr, fused=True)     # one launch for the 10 parameter tensors (the for-each path: 16 launches, 0.36 ms at C3)
        self.sched = torch.optim.lr_scheduler.StepLR(self.opt, step_size=100, gamma=0.8)
        self.rows = slice(sh.row0, sh.row0 + sh.rows)
        self.gt_srgb = linear_to_srgb(gt_image[self.rows])
        self.gt_mean = gt_image.mean()                       # gt is replicated: the global mean needs no exchange
        self.n_total = H * W * 3
        self.last = {}

    def _maps(self):
        H, W = self.scene.H, self.scene.W
        e0, e1 = self.er0, self.er1
        arm = self.net(self.start_arm[e0 * W:e1 * W], hw=(H, W), row0=e0)                                        # :493
        albedo, roughness, metallic = arm[..., 0:3].clamp(0, 1), (arm[..., 3:4] * 0.93 + 0.07).clamp(0, 1), arm[..., 4:5].clamp(0, 1)
        mat = dict(self.mat)
        for key, k, val, c in (("albedo", "a", albedo, 3), ("roughness", "r", roughness, 1), ("metallic", "m", metallic, 1)):
            if k in self.part:
                if (e0, e1) == (0, H):
                    mat[key] = val.reshape(H, W, c)
                else:                                         # rows outside [e0, e1) are never read by this rank's kernels
                    mat[key] = torch.cat([self.mat[key][:e0], val.reshape(e1 - e0, W, c), self.mat[key][e1:]], 0)
        return mat

    def _step(self, seed):
        if os.environ.get("MB200_POSMLP_AUTOGRAD", "0") == "1":
            return self._step_autograd(seed)
        return self._step_fused(seed)

    def _fused_setup(self):
        sc, sh = self.scene, self.shard
        dev, H, W = sc.device, sc.H, sc.W
        self._maps_buf = {k: self.mat[k].detach().float().clone().contiguous() for k in ("albedo", "roughness", "metallic")}
        sizes = [H * W * 3, H * W, H * W]
        self._gflat = torch.zeros(sum(sizes), device=dev)
        ga, gr, gm = torch.split(self._gflat, sizes)
        self._g = {"albedo": ga.view(H, W, 3), "roughness": gr.view(H, W, 1), "metallic": gm.view(H, W, 1)}
        self._scal = torch.zeros(2, device=dev)
        self._scal[0:1] = (self.gt_mean * self.n_total).reshape(1)
        self._sums2 = torch.zeros(2, device=dev)
        self._scratch = torch.zeros(_abi.lib.mb200_reduce_scratch_bytes() // 4 + 1, dtype=torch.int32, device=dev)
        self._grad_full, self._grad_img = sh.halo_buffer(3, dev)
        self._pred_srgb = torch.empty(sh.rows, W, 3, device=dev)
        self._gt_srgb = self.gt_srgb.float().contiguous()
        self._side = torch.cuda.Stream(dev)
        self._ev_fwd, self._ev_w, self._wpart = torch.cuda.Event(), torch.cuda.Event(), None

    def _step_fused(self, seed):
        """The iteration of :471-552 with everything between the network and the two render kernels done by the fused loss kernels
        of csrc/mb200_optim.cu (as in FusedBRDFOptimizer) instead of ~60 elementwise autograd launches: the head post-processing
        (:494-506: clamp, roughness * 0.93 + 0.07) and its chain rule, the aux l1 terms and the loss gradient are a handful of
        no-grad tensor ops on the (n, 5) network output; autograd only carries arm -> network weights."""
        if not hasattr(self, "_gflat"):
            self._fused_setup()
        sc, sh, lib, st = self.scene, self.shard, _abi.lib, _abi.stream_ptr()
        H, W, e0, e1 = sc.H, sc.W, self.er0, self.er1
        arm = self.net(self.start_arm[e0 * W:e1 * W], hw=(H, W), row0=e0)                                        # :493
        with torch.no_grad():
            x = arm.detach()
            rgh_pre = x[:, 3:4] * 0.93 + 0.07
            head = {"albedo": x[:, 0:3].clamp(0, 1), "roughness": rgh_pre.clamp(0, 1), "metallic": x[:, 4:5].clamp(0, 1)}
            for key, k, c in (("albedo", "a", 3), ("roughness", "r", 1), ("metallic", "m", 1)):
                if k in self.part:
                    self._maps_buf[key][e0:e1].copy_(head[key].view(e1 - e0, W, c))
            a, r, m = self._maps_buf["albedo"], self._maps_buf["roughness"], self._maps_buf["metallic"]
            env_pack = sc.prepared_env()
            seed_grad = _rop.default_seed_grad(int(seed))
            img = _rop._forward(sc, self.spp, int(seed), a, r, m, None, env_pack)
            main = torch.cuda.current_stream(sc.device)
            if sc.filter == _abi.FILTER_GAUSSIAN:
                self._ev_fwd.record(main)
                with torch.cuda.stream(self._side):
                    self._side.wait_event(self._ev_fwd)
                    self._wpart = _rop._film_weights(sc, self.spp, seed_grad, env_pack[2].res_x, out=self._wpart)
                    self._ev_w.record(self._side)
            n = img.numel()
            _abi.check(lib.mb200_image_sum(_abi.ptr(img), n, C.c_void_p(self._scal.data_ptr() + 4), _abi.ptr(self._scratch), st), "mb200_image_sum")
            if sh.world_size > 1:
                sh.all_reduce_sum(self._scal[1:2])
            _abi.check(lib.mb200_loss_srgb_sums(_abi.ptr(img), _abi.ptr(self._gt_srgb), n, _abi.ptr(self._scal), _abi.ptr(self._sums2),
                                                _abi.ptr(self._pred_srgb), _abi.ptr(self._scratch), st), "mb200_loss_srgb_sums")
            if sh.world_size > 1:
                sh.all_reduce_sum(self._sums2)
            _abi.check(lib.mb200_loss_srgb_grad(_abi.ptr(img), _abi.ptr(self._gt_srgb), n, _abi.ptr(self._scal), _abi.ptr(self._sums2),
                                                self.n_total, _abi.ptr(self._grad_img), st), "mb200_loss_srgb_grad")
            grad = sh.halo_exchange_inplace(self._grad_full) if sh.world_size > 1 else self._grad_img
            self._gflat.zero_()
            if sc.filter == _abi.FILTER_GAUSSIAN:
                main.wait_event(self._ev_w)
            _rop._backward(sc, self.spp, seed_grad, a, r, m, None, env_pack, grad, "a" in self.part, "r" in self.part, "m" in self.part, False, False,
                           out=(self._g["albedo"], self._g["roughness"], self._g["metallic"]), wpart=self._wpart)
            # chain rule of the head post-processing + the aux l1 terms (own rows only), on the (n, 5) network output
            g_arm = torch.zeros_like(x)
            npx = float(H * W)
            own = torch.zeros(e1 - e0, 1, 1, device=x.device)
            own[sh.row0 - e0:sh.row0 - e0 + sh.rows] = 1.0
            for key, k, c, sl, pre, scale in (("albedo", "a", 3, slice(0, 3), x[:, 0:3], 1.0), ("roughness", "r", 1, slice(3, 4), rgh_pre, 0.93),
                                              ("metallic", "m", 1, slice(4, 5), x[:, 4:5], 1.0)):
                if k not in self.part:
                    continue
                gk = self._g[key][e0:e1] + (self.scale_delta / (npx * c)) * own * torch.sign(self._maps_buf[key][e0:e1] - self.ori[key][e0:e1])
                inside = ((pre >= 0) & (pre <= 1)).view(-1, c)
                g_arm[:, sl] = gk.reshape(-1, c) * inside * scale
        arm.backward(g_arm)
        if sh.world_size > 1:
            params = [p for p in self.net.parameters() if p.grad is not None]
            flat = torch.cat([p.grad.reshape(-1) for p in params])
            sh.all_reduce_sum(flat)
            off = 0
            for p in params:
                p.grad.copy_(flat[off:off + p.numel()].view_as(p)); off += p.numel()
        self.opt.step()
        self.opt.zero_grad(set_to_none=True)
        if self.opt.param_groups[0]["lr"] > 1.5e-4:
            self.sched.step()
        self.last = {"loss_mse": self._sums2[0] / self.n_total, "loss_l1": self._sums2[1] / self.n_total, "pred": self._pred_srgb}
        return self.last["loss_mse"]

    def _step_autograd(self, seed):
        """Reference formulation through torch autograd (kept for A/B and as the checker of _step_fused)."""
        sh = self.shard
        mat = self._maps()
        pred = render(self.scene, spp=self.spp, seed=seed, albedo=mat["albedo"], roughness=mat["roughness"], metallic=mat["metallic"],
                      halo_exchange=sh.halo_exchange if sh.world_size > 1 else None)
        s_pred = sh.all_reduce_sum(pred.detach().sum().reshape(1).clone())
        pred = pred * (self.gt_mean / (s_pred / self.n_total))
        pred_srgb = linear_to_srgb(pred)
        diff = pred_srgb - self.gt_srgb
        mse_local, l1_local = (diff * diff).sum() / self.n_total, diff.abs().sum() / self.n_total
        g = sh.all_reduce_sum(torch.stack([mse_local.detach(), l1_local.detach()]))
        loss_mse, loss_l1 = g[0], g[1]
        aux = 0.0
        npx = self.scene.H * self.scene.W
        for key, k, c in (("albedo", "a", 3), ("roughness", "r", 1), ("metallic", "m", 1)):
            if k in self.part:
                aux = aux + (mat[key][self.rows] - self.ori[key][self.rows]).abs().sum() / (npx * c)
        loss = 3 * (loss_l1 / loss_mse) * mse_local + l1_local + aux * self.scale_delta      # summed over the ranks = the reference's loss
        loss.backward()
        if sh.world_size > 1:
            params = [p for p in self.net.parameters() if p.grad is not None]
            flat = torch.cat([p.grad.reshape(-1) for p in params])
            sh.all_reduce_sum(flat)
            off = 0
            for p in params:
                p.grad.copy_(flat[off:off + p.numel()].view_as(p)); off += p.numel()
        self.opt.step()
        self.opt.zero_grad(set_to_none=True)
        if self.opt.param_groups[0]["lr"] > 1.5e-4:
            self.sched.step()
        self.last = {"loss_mse": loss_mse.detach(), "loss_l1": loss_l1.detach(), "pred": pred_srgb.detach()}
        return loss_mse.detach()


class _SumGradOverRanks(torch.autograd.Function):
    """Identity whose backward all-reduces (sums) the gradient over the ranks of a ShardContext: every rank renders its own
    rows, the envmap they all share receives the sum of their texel gradients (exchange step 3)."""

    @staticmethod
    def forward(ctx, x, shard):
        ctx.shard = shard
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        ctx.shard.all_reduce_sum(g)
        return g, None


class EnvmapNetOptimizer(_ShardedStep):
    """Envmap phase exactly as the reference runs it (inverse_img_w_mi.py:117-124, :222-256): `envmap_net` = PosMLP(in_dims=5,
    out_dims=3, 'envmap') applied to a constant all-ones (env_h*env_w, 3) input gives the (16, 32, 3) envmap that
    `render_envmap` shades with; loss = mse + l1 in sRGB; Adam(lr=1e-3) + StepLR(100, 0.8) in the first outer loop.
    With the image sharded, every rank holds a replica of the network; the texel gradients are summed over ranks before they
    enter the (replicated, therefore identical) network backward."""

    def __init__(self, scene, gt_image, env_h=16, env_w=32, spp=64, lr=1e-3, shard=None, net=None):
        from .mymodels.mlps import PosMLP
        self.scene, self.spp, self.env_h, self.env_w = scene, spp, env_h, env_w
        self.shard = shard or ShardContext(scene.H, scene.W)
        dev = scene.device
        self.net = net if net is not None else PosMLP(in_dims=5, out_dims=3, dims=[256] * 4, skip_connection=[1, 3], weight_norm=False,
                                                      multires_view=2, output_type="envmap", color_ch=3).to(dev)      # :117-124
        self.start_envmap = torch.ones(env_h * env_w, 3, device=dev)
        self.opt = torch.optim.Adam(self.net.parameters(), lr=lr, fused=True)
        self.sched = torch.optim.lr_scheduler.StepLR(self.opt, step_size=100, gamma=0.8)
        rows = slice(self.shard.row0, self.shard.row0 + self.shard.rows)
        self.gt_srgb = linear_to_srgb(gt_image[rows].contiguous())
        self.n_img = float(scene.H * scene.W * 3)
        self.last = {}

    def _step(self, seed):
        sh = self.shard
        envmap_pred = self.net(self.start_envmap, hw=(self.env_h, self.env_w)).reshape(self.env_h, self.env_w, 3)           # :238-239
        env = _SumGradOverRanks.apply(envmap_pred, sh) if sh.world_size > 1 else envmap_pred
        pred = render(self.scene, spp=self.spp, seed=seed, envmap=env, halo_exchange=sh.halo_exchange if sh.world_size > 1 else None)
        diff = linear_to_srgb(pred) - self.gt_srgb
        loss = (diff * diff).sum() / self.n_img + diff.abs().sum() / self.n_img          # this rank's share of mse + l1 (:243-245)
        loss.backward()
        self.opt.step()
        self.opt.zero_grad(set_to_none=True)
        self.sched.step()
        self.last = {"loss": loss.detach(), "envmap": envmap_pred.detach()}
        return loss.detach()


class EnvmapOptimizer(_ShardedStep):
    """Envmap phase (inverse_img_w_mi.py:237-256) with the envmap texels as direct parameters (the reference
    drives them through `envmap_net`; see mymodels/mlps.py for that module).  Gradients of the envmap are summed
    over ranks once per iteration."""

    def __init__(self, scene, env_init, gt_image, spp=64, lr=1e-3, shard=None):
        self.scene, self.spp = scene, spp
        self.shard = shard or ShardContext(scene.H, scene.W)
        self.env = torch.nn.Parameter(env_init.detach().clone())
        self.opt = torch.optim.Adam([self.env], lr=lr)
        self.rows = slice(self.shard.row0, self.shard.row0 + self.shard.rows)
        self.gt_srgb = linear_to_srgb(gt_image[self.rows].contiguous())
        self.n_img = float(scene.H * scene.W * 3)

    def _step(self, seed):
        sh = self.shard
        env = NF.softplus(self.env)
        pred = render(self.scene, spp=self.spp, seed=seed, envmap=env,
                      halo_exchange=sh.halo_exchange if sh.world_size > 1 else None)
        diff = linear_to_srgb(pred) - self.gt_srgb
        loss = (diff * diff).sum() / self.n_img + diff.abs().sum() / self.n_img
        loss.backward()
        sh.all_reduce_sum(self.env.grad)
        self.opt.step()
        self.opt.zero_grad(set_to_none=True)
        return loss.detach()
