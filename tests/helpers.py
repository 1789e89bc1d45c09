"""Shared test plumbing: build the same synthetic case for the CPU oracle (numpy) and the CUDA path (torch)."""
import numpy as np
import torch

from oracle import oracle as orc
from materialist_b200 import synthetic
from materialist_b200.camera import Camera

REF_FLAGS = orc.FLAG_WO_WORLD_QUIRK | orc.FLAG_ROW_STRIDE_H | orc.FLAG_ENV_HALF_TEXEL


def rel_l2(x, y):
    x = np.asarray(x, np.float64); y = np.asarray(y, np.float64)
    return float(np.linalg.norm(x - y) / max(np.linalg.norm(y), 1e-30))


class Case:
    """A synthetic G-buffer case (SURVEY §8d generators)."""

    def __init__(self, H, W, spp, He, We, *, sun=2000.0, invalid_border=0, env_mode=orc.ENV_ASSIGNED, gaussian=True,
                 use_mesh_normal=True, flags=None, max_depth=4, mat_seed=1):
        self.H, self.W, self.spp, self.gaussian, self.use_mesh_normal = H, W, spp, gaussian, use_mesh_normal
        self.max_depth, self.env_mode = max_depth, env_mode
        self.cam = Camera(width=W, height=H)
        self.pos, self.nrm, self.valid = synthetic.gbuffer(H, W, self.cam, invalid_border=invalid_border)
        a, r, m = synthetic.materials(H, W, seed_base=mat_seed)
        self.a, self.r, self.m = a.numpy(), r.numpy(), m.numpy()
        self.n = synthetic.normal_map(self.nrm).numpy()
        self.env = synthetic.envmap(He, We, sun=sun).numpy()
        self.flags = (REF_FLAGS if H == W else REF_FLAGS & ~orc.FLAG_ROW_STRIDE_H) if flags is None else flags
        self.gpos = np.ascontiguousarray(np.concatenate([self.pos, self.valid[..., None].astype(np.float32)], -1))
        self.gnrm = np.ascontiguousarray(np.concatenate([self.nrm, np.zeros((H, W, 1), np.float32)], -1))

    # ---------------------------------------------------------------- oracle side
    def cfg(self, d, seed, row0=0, rows=None, extra_flags=0):
        cam = self.cam
        c = orc.Cfg()
        c.H, c.W, c.spp, c.max_depth = self.H, self.W, self.spp, self.max_depth
        c.seed = seed & 0xFFFFFFFF
        c.filter = orc.FILTER_GAUSSIAN if self.gaussian else orc.FILTER_BOX
        c.flags = self.flags | extra_flags
        c.use_mesh_normal = int(self.use_mesh_normal)
        c.row0, c.rows = row0, (self.H if rows is None else rows)
        c.view[:] = cam.view_matrix.reshape(-1).tolist()
        c.proj[:] = cam.proj_matrix.reshape(-1).tolist()
        c.cam_to_world[:] = cam.to_world.astype(np.float32).reshape(-1).tolist()
        c.tan_half_fov_x = cam.tan_half_fov_x
        c.env_u_shift = float(np.float32(0.5) / np.float32(d.res_x - 1)) if c.flags & orc.FLAG_ENV_HALF_TEXEL else 0.0
        return c

    def oracle_fwd(self, O, seed, want_indices=False, row0=0, rows=None, extra_flags=0):
        env_int, hier, d = O.env_prepare(self.env, self.env_mode)
        return O.render_fwd(self.cfg(d, seed, row0, rows, extra_flags), self.gpos, self.gnrm, self.a, self.r, self.m,
                            self.n, env_int, hier, d, want_indices=want_indices)

    def oracle_bwd(self, O, seed_grad, grad_img, want=("a", "r", "m", "env"), row0=0, rows=None):
        env_int, hier, d = O.env_prepare(self.env, self.env_mode)
        g = O.render_bwd(self.cfg(d, seed_grad, row0, rows), self.gpos, self.gnrm, self.a, self.r, self.m, self.n,
                         env_int, hier, d, grad_img, want=want)
        if "env_int" in g:
            g["env"] = O.env_grad_finish(g.pop("env_int"), self.env.shape[1], self.env_mode)
        return g

    # ---------------------------------------------------------------- CUDA side
    def scene(self, device="cuda"):
        import materialist_b200 as mb
        s = mb.Scene(self.pos, self.nrm, self.valid, camera=self.cam, envmap=torch.from_numpy(self.env),
                     use_mesh_normal=self.use_mesh_normal, max_depth=self.max_depth,
                     rfilter="gaussian" if self.gaussian else "box", device=device, flags=self.flags)
        s.set_envmap(torch.from_numpy(self.env), self.env_mode)
        return s

    def torch_maps(self, device="cuda", requires_grad=False):
        out = [torch.from_numpy(x).to(device).requires_grad_(requires_grad) for x in (self.a, self.r, self.m, self.n)]
        return out
