"""Host-buffer front end of `render_w_brdf` (inverse_img_w_mi.py:69-80) for callers whose maps and gradients live in HOST
memory (a CPU-side optimiser, a serving process): every step uploads albedo / roughness / metallic and d(loss)/d(image) from
pinned buffers and downloads the image and the material gradients — with the copies on their own CUDA streams so that they
overlap the two render kernels instead of serialising with them:

    step i:   [H2D maps i+1, grad i+1]            (copy-in stream, double buffered, during step i's kernels)
              fwd ──► [D2H image i] ──────────┐   (copy-out stream, overlaps the adjoint render)
              adjoint ──► [D2H gradients i]   ┘   (copy-out stream, overlaps step i+1's forward)

Nothing is skipped: each step's inputs cross PCIe in that step, each step's outputs are read back; `synchronize()` (or the
next step's reuse of a buffer) orders the host's view.  PyTorch is plumbing here (streams, events, pinned memory)."""
import torch

from .renderop import render


class HostPipelinedRenderWBRDF:
    def __init__(self, scene, spp, halo_exchange=None):
        self.scene, self.spp, self.halo_exchange = scene, int(spp), halo_exchange
        dev = scene.device
        self.main = torch.cuda.current_stream(dev)
        self.s_in, self.s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        H, W = scene.H, scene.W
        mk = lambda *shape: [torch.empty(*shape, device=dev) for _ in range(2)]
        self.a, self.r, self.m, self.g = mk(H, W, 3), mk(H, W, 1), mk(H, W, 1), mk(scene.rows, W, 3)
        self.ev_in = [torch.cuda.Event(), torch.cuda.Event()]
        self.ev_free = [torch.cuda.Event(), torch.cuda.Event()]       # buffer set no longer read by the main stream
        self.ev_out = torch.cuda.Event()                              # previous step's downloads have left their source tensors
        self._staged = None
        self._keep = []

    def stage(self, slot, ha, hr, hm, hgrad):
        """Upload one step's inputs (pinned host tensors) into buffer set `slot` on the copy-in stream."""
        with torch.cuda.stream(self.s_in):
            self.s_in.wait_event(self.ev_free[slot])            # (a never-recorded event is a no-op)
            self.a[slot].copy_(ha, non_blocking=True); self.r[slot].copy_(hr, non_blocking=True)
            self.m[slot].copy_(hm, non_blocking=True); self.g[slot].copy_(hgrad, non_blocking=True)
            self.ev_in[slot].record(self.s_in)
        self._staged = slot

    def step(self, seed, slot, himg, hga, hgr, hgm, next_inputs=None):
        """Render forward + adjoint with buffer set `slot` (already staged); downloads go to the pinned tensors himg / hga /
        hgr / hgm.  `next_inputs` = (ha, hr, hm, hgrad) of the following step is uploaded into the other set meanwhile."""
        main = self.main
        main.wait_event(self.ev_in[slot])
        if next_inputs is not None:
            self.stage(1 - slot, *next_inputs)
        a = self.a[slot].detach().requires_grad_(True); r = self.r[slot].detach().requires_grad_(True); m = self.m[slot].detach().requires_grad_(True)
        img = render(self.scene, spp=self.spp, seed=seed, albedo=a, roughness=r, metallic=m, halo_exchange=self.halo_exchange)
        ev_img = torch.cuda.Event(); ev_img.record(main)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(ev_img)
            himg.copy_(img.detach(), non_blocking=True)
        img.detach().record_stream(self.s_out)
        img.backward(self.g[slot])
        ev_bwd = torch.cuda.Event(); ev_bwd.record(main)
        self.ev_free[slot].record(main)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(ev_bwd)
            hga.copy_(a.grad, non_blocking=True); hgr.copy_(r.grad, non_blocking=True); hgm.copy_(m.grad, non_blocking=True)
            self.ev_out.record(self.s_out)
        for t in (a.grad, r.grad, m.grad):
            t.record_stream(self.s_out)

    def synchronize(self):
        self.s_in.synchronize(); self.s_out.synchronize(); self.main.synchronize()
