"""The operator boundary of the reference's inverse_img_w_mi.py / render_final.py under their own names, so that
the optimisation loop (`optimize_envmap_ARMN`, kept as host code in PyTorch) and the relighting script run on the
B200 operator with `import materialist_b200.inverse_img_w_mi as ...` in place of Mitsuba:

    scene  = load_estimated_mesh(mesh_path, use_mesh_normal, max_path=4)          # inverse_img_w_mi.py:30-56
    params = traverse(scene); params['shape.bsdf.a'] = T; ...; params.update()     # :216-220, :334-342
    img    = render_w_brdf(scene, albedo, roughness, metallic, normal=None, spp=64)   # :69-80   grads -> a, r, m[, n]
    img    = render_envmap(scene, envmap, spp=64)                                     # :59-67   grads -> envmap
    img    = render(scene, spp=64, seed=i)                                            # render_final.py:194 (mi.render)
    frames = render_rolling_envmap(scene, envmap, frames=36, rotation_step=10)        # render_final.py:300-418 (intended behaviour)
"""

import numpy as np
import torch

from . import _abi
from .gbuffer import gbuffer_from_ply, load_estimated_brdf, read_image
from .renderop import render, render_envmap, render_w_brdf  # noqa: F401  (re-exported under the reference's names)
from .scene import Camera, Scene, traverse  # noqa: F401



def load_estimated_mesh(mesh_path, use_mesh_normal, max_path=4, envmap=None, width=512, height=512, device="cuda", mode="mesh"):
    """Scene handle for the depth-derived PLY.  `envmap`: (He,We,3) array/tensor or a path to a .hdr/.exr; the reference
    hard-wires 'envmaps/0.hdr' (inverse_img_w_mi.py:54) — pass it explicitly here.
    mode='mesh' (default): the triangle mesh is traced like the reference does (per-sample hits, shadow rays, max_path-1
    bounces; csrc/mb200_mesh.cu).  mode='gbuffer': the fast approximation — per-pixel G-buffer (vertex k <-> pixel k),
    no occlusion / interreflection."""
    cam = Camera.from_json(width=width, height=height)
    if isinstance(envmap, str):
        envmap = read_image(envmap)[..., :3]
    if envmap is None:
        envmap = np.ones((16, 32, 3), np.float32)
    envmap = torch.as_tensor(np.ascontiguousarray(envmap))
    if mode == "mesh":
        from .mesh import read_ply_mesh
        verts, tris = read_ply_mesh(mesh_path)
        return Scene.from_mesh(verts, tris, cam, device=device, envmap=envmap, use_mesh_normal=use_mesh_normal, max_depth=max_path)
    if mode != "gbuffer":
        raise ValueError("mode must be 'mesh' or 'gbuffer'")
    pos, nrm, valid = gbuffer_from_ply(mesh_path, height, width, cam)
    return Scene(pos, nrm, valid, camera=cam, envmap=envmap, use_mesh_normal=use_mesh_normal, max_depth=max_path, device=device)


def load_estimated_mesh_w_env(mesh_path, envmap_path, mat_dir=None, bsdf="matDiffBSDF", max_depth=4, **kw):
    """render_final.py:19-97: bsdf = 'matDiffBSDF' / {'name': 'matDiffBSDF'}, or {'name': 'TransBSDF', 'ior': ..,
    'keep_albedo_color': ..} as trans_edit.py:18 passes it (`mat_dir` is accepted for signature compatibility: maps are
    assigned through traverse(), as both callers do)."""
    if isinstance(bsdf, str):
        bsdf = {"name": bsdf}
    if bsdf.get("name") in ("MatDiffBSDF",):
        bsdf = dict(bsdf, name="matDiffBSDF")
    if bsdf.get("name") not in ("matDiffBSDF", "TransBSDF"):
        raise ValueError("Invalid bsdf type")
    scene = load_estimated_mesh(mesh_path, True, max_depth, envmap=envmap_path, **kw)
    return scene.set_bsdf(bsdf)


def render_w_mi(scene, mat_dir, n_iter=10, spp=64):
    """render_final.py:148-203 minus the OptiX AI denoiser and file output: average of `n_iter` renders, seeds 0..n-1."""
    mat = load_estimated_brdf(mat_dir)
    params = traverse(scene)
    dev = scene.device
    params["shape.bsdf.a"] = torch.from_numpy(mat["albedo"]).to(dev)
    params["shape.bsdf.r"] = torch.from_numpy(mat["roughness"]).to(dev)
    params["shape.bsdf.m"] = torch.from_numpy(mat["metallic"]).to(dev)
    params.update()
    acc = None
    for i in range(n_iter):
        img = render(scene, spp=spp, seed=i)
        acc = img if acc is None else acc + img
    return acc / n_iter


def rotate_envmap(envmap, angle_deg):
    """render_final.py:290-298: np.roll by int(angle/360 * W) columns."""
    W = envmap.shape[1]
    return torch.roll(envmap, shifts=int(angle_deg / 360.0 * W), dims=1)


def render_rolling_envmap(scene, envmap, frames=36, rotation_step=10, spp=32, n_iter=1, frame_ids=None):
    """render_final.py:300-418 as INTENDED (the CLI path is unreachable as shipped): one relight per rolled
    envmap, frames independent — shard `frame_ids` over GPUs (replicas only, no collective)."""
    out = {}
    for k in (range(frames) if frame_ids is None else frame_ids):
        scene.set_envmap(rotate_envmap(envmap, k * rotation_step), _abi.ENV_FILE)
        acc = None
        for i in range(n_iter):
            img = render(scene, spp=spp, seed=i)
            acc = img if acc is None else acc + img
        out[k] = acc / n_iter
    return out
