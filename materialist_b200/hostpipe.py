"""Host-buffer front end of `render_w_brdf` (inverse_img_w_mi.py:69-80) for callers whose maps and gradients live in HOST
memory (a CPU-side optimiser, a serving process): every step uploads albedo / roughness / metallic and d(loss)/d(image) from
pinned buffers and downloads the image and the material gradients — with the copies on their own CUDA streams so that they
overlap the two render kernels instead of serialising with them:

    step i:   [H2D maps i+1, grad i+1]            (copy-in stream, double buffered, during step i's kernels)
              fwd ──► [D2H image i] ──────────┐   (copy-out stream, overlaps the adjoint render)
              adjoint ──► [D2H gradients i]   ┘   (copy-out stream, overlaps step i+1's forward)

Nothing is skipped: each step's inputs cross PCIe in that step, each step's outputs are read back; `synchronize()` (or the
next step's reuse of a buffer) orders the host's view.  PyTorch is plumbing here (streams, events, pinned memory).

Row sharding: a rank moves only what its kernels touch — in G-buffer mode the maps of its own rows plus the 2-row film halo
(`map_rows`) up, and the image / gradients of its own rows (`out_rows`) down; a traced scene reads and scatters anywhere, so there
the whole maps travel.  (Uploading whole-image maps on every rank made the end-to-end rate stop scaling at 8 GPUs.)"""
import torch

from .renderop import render


class HostPipelinedRenderWBRDF:
    def __init__(self, scene, spp, halo_exchange=None):
        self.scene, self.spp, self.halo_exchange = scene, int(spp), halo_exchange
        dev = scene.device
        self.main = torch.cuda.current_stream(dev)
        self.s_in, self.s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        H, W = scene.H, scene.W
        r0, r1 = scene.row0, scene.row0 + scene.rows
        whole = scene.mesh is not None
        self.map_rows = (0, H) if whole else (max(0, r0 - 2), min(H, r1 + 2))       # rows of a / r / m the kernels read
        self.out_rows = (0, H) if whole else (r0, r1)                               # rows that receive material gradients
        mk = lambda *shape: [torch.zeros(*shape, device=dev) for _ in range(2)]
        self.a, self.r, self.m, self.g = mk(H, W, 3), mk(H, W, 1), mk(H, W, 1), mk(scene.rows, W, 3)
        # device-side staging of the outputs, owned by the pipe: the copy-out stream never touches a tensor the caching allocator
        # may recycle (no record_stream, no deferred frees)
        no = self.out_rows[1] - self.out_rows[0]
        self.o_img, self.o_ga, self.o_gr, self.o_gm = mk(scene.rows, W, 3), mk(no, W, 3), mk(no, W, 1), mk(no, W, 1)
        ev = lambda: [torch.cuda.Event(), torch.cuda.Event()]
        self.ev_in, self.ev_free, self.ev_img, self.ev_bwd, self.ev_out = ev(), ev(), ev(), ev(), ev()

    def stage(self, slot, ha, hr, hm, hgrad):
        """Upload one step's inputs (pinned host tensors: rows `map_rows` of the maps, the shard rows of d(loss)/d(image)) into
        buffer set `slot` on the copy-in stream."""
        m0, m1 = self.map_rows
        with torch.cuda.stream(self.s_in):
            self.s_in.wait_event(self.ev_free[slot])            # (a never-recorded event is a no-op)
            self.a[slot][m0:m1].copy_(ha, non_blocking=True); self.r[slot][m0:m1].copy_(hr, non_blocking=True)
            self.m[slot][m0:m1].copy_(hm, non_blocking=True); self.g[slot].copy_(hgrad, non_blocking=True)
            self.ev_in[slot].record(self.s_in)

    def step(self, seed, slot, himg, hga, hgr, hgm, next_inputs=None):
        """Render forward + adjoint with buffer set `slot` (already staged); downloads go to the pinned tensors himg / hga /
        hgr / hgm.  `next_inputs` = (ha, hr, hm, hgrad) of the following step is uploaded into the other set meanwhile."""
        main, s_out = self.main, self.s_out
        main.wait_event(self.ev_in[slot])
        main.wait_event(self.ev_out[slot])                      # the downloads of two steps ago have left this slot's staging
        if next_inputs is not None:
            self.stage(1 - slot, *next_inputs)
        a = self.a[slot].detach().requires_grad_(True); r = self.r[slot].detach().requires_grad_(True); m = self.m[slot].detach().requires_grad_(True)
        img = render(self.scene, spp=self.spp, seed=seed, albedo=a, roughness=r, metallic=m, halo_exchange=self.halo_exchange)
        self.o_img[slot].copy_(img.detach())
        self.ev_img[slot].record(main)
        with torch.cuda.stream(s_out):
            s_out.wait_event(self.ev_img[slot])
            himg.copy_(self.o_img[slot], non_blocking=True)
        img.backward(self.g[slot])
        o0, o1 = self.out_rows
        self.o_ga[slot].copy_(a.grad[o0:o1]); self.o_gr[slot].copy_(r.grad[o0:o1]); self.o_gm[slot].copy_(m.grad[o0:o1])
        self.ev_bwd[slot].record(main)
        self.ev_free[slot].record(main)
        with torch.cuda.stream(s_out):
            s_out.wait_event(self.ev_bwd[slot])
            hga.copy_(self.o_ga[slot], non_blocking=True); hgr.copy_(self.o_gr[slot], non_blocking=True); hgm.copy_(self.o_gm[slot], non_blocking=True)
            self.ev_out[slot].record(s_out)
        self.last_out = self.ev_out[slot]

    def synchronize(self):
        self.s_in.synchronize(); self.s_out.synchronize(); self.main.synchronize()
