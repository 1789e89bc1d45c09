"""Host-side optimisation iteration around the render operator — the body of the BRDF phase of
`optimize_envmap_ARMN` with `model_name == 'none'` (inverse_img_w_mi.py:343-446) and of the envmap phase
(:237-256), minus file I/O, tqdm and the per-iteration `.item()` host syncs.  Everything stays on the device;
with a ShardContext the three scalar sums and the image-gradient halo are exchanged between ranks.
"""
import torch
import torch.nn.functional as NF

from .parallel import ShardContext
from .renderop import render


def linear_to_srgb(image):
    """myutils/misc.py:167-170"""
    return image ** (1.0 / 2.2)


class DirectBRDFOptimizer:
    """`Directly optimizing {a,r,m} without neural network` (inverse_img_w_mi.py:346-446).

    mat: dict of full-image CUDA tensors albedo (H,W,3), roughness (H,W,1), metallic (H,W,1);
    gt_image: (H,W,3) linear radiance.  Each rank renders and optimises its own row shard.
    """

    def __init__(self, scene, mat, gt_image, optimize_part="arm", spp=64, lr=3e-4, scale_delta=0.1, shard=None):
        self.scene, self.spp, self.scale_delta, self.part = scene, spp, scale_delta, optimize_part
        self.shard = shard or ShardContext(scene.H, scene.W)
        scene.set_shard(self.shard.row0, self.shard.rows)
        self.mat = {k: v.detach().clone() for k, v in mat.items()}
        self.ori = {k: v.detach().clone() for k, v in mat.items()}
        self.params = {}
        if "a" in optimize_part: self.params["albedo"] = torch.nn.Parameter(mat["albedo"].clone())
        if "r" in optimize_part: self.params["roughness"] = torch.nn.Parameter(mat["roughness"].clone())
        if "m" in optimize_part: self.params["metallic"] = torch.nn.Parameter(mat["metallic"].clone())
        self.opt = torch.optim.Adam(self.params.values(), lr=lr)
        self.sched = torch.optim.lr_scheduler.StepLR(self.opt, step_size=100, gamma=0.8)
        r0, r1 = self.shard.row0, self.shard.row0 + self.shard.rows
        self.rows = slice(r0, r1)
        self.gt = gt_image[self.rows].contiguous()
        self.gt_srgb = linear_to_srgb(self.gt)
        self.n_img = float(scene.H * scene.W * 3)
        self.gt_sum = self.shard.all_reduce_sum(self.gt.sum().reshape(1).clone())
        self.last = {}

    def step(self, seed):
        p, sh, rows = self.params, self.shard, self.rows
        mat = dict(self.mat)
        if "albedo" in p: mat["albedo"] = p["albedo"].clamp(0, 1)
        if "roughness" in p: mat["roughness"] = p["roughness"].clamp(0.07, 1)
        if "metallic" in p: mat["metallic"] = p["metallic"].clamp(0, 1)
        pred = render(self.scene, spp=self.spp, seed=seed, albedo=mat["albedo"], roughness=mat["roughness"],
                      metallic=mat["metallic"], halo_exchange=sh.halo_exchange if sh.world_size > 1 else None)
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
        scale_ratio = sums[1] / sums[0]                        # loss_l1.detach() / loss_mse.detach()
        loss = 3 * scale_ratio * loss_mse_l + loss_l1_l + aux * self.scale_delta
        loss.backward()
        self.opt.step()
        self.opt.zero_grad(set_to_none=True)
        self.sched.step()          # the reference stops decaying below lr 1.5e-4 (:432) through a host-side read
        self.last = {"loss_mse": sums[0] / self.n_img, "loss_l1": sums[1] / self.n_img, "pred": pred_srgb}
        return loss.detach()


class EnvmapOptimizer:
    """Envmap phase (inverse_img_w_mi.py:237-256) with the envmap texels as direct parameters (the reference
    drives them through `envmap_net`; see mymodels/mlps.py for that module).  Gradients of the envmap are summed
    over ranks once per iteration."""

    def __init__(self, scene, env_init, gt_image, spp=64, lr=1e-3, shard=None):
        self.scene, self.spp = scene, spp
        self.shard = shard or ShardContext(scene.H, scene.W)
        scene.set_shard(self.shard.row0, self.shard.rows)
        self.env = torch.nn.Parameter(env_init.detach().clone())
        self.opt = torch.optim.Adam([self.env], lr=lr)
        self.rows = slice(self.shard.row0, self.shard.row0 + self.shard.rows)
        self.gt_srgb = linear_to_srgb(gt_image[self.rows].contiguous())
        self.n_img = float(scene.H * scene.W * 3)

    def step(self, seed):
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
