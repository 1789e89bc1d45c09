"""Drop-in for the reference's trans_edit.py (transparency editing): the masked region of an optimised material set is
re-rendered as glass with the `TransBSDF` plugin (myutils/mi_plugin.py:1477-1770) on the B200 operator
(mb200_trans_shade_fwd / mb200_trans_mesh_shade_fwd).  Forward only, as in the reference.

    scene = load_estimated_mesh_w_env(mesh_path, env_path, mat_dir, bsdf={'name': 'TransBSDF', 'ior': ior,
                                                                         'keep_albedo_color': keep})   # trans_edit.py:18
    img = transprancy_edit(scene, mat, ior, keep_albedo_color, specTrans, n_iter=10)                     # trans_edit.py:16-49
"""
import torch

from .renderop import render
from .scene import traverse


def edit_materials(mat, keep_albedo_color):
    """trans_edit.py:21-29: inside the mask albedo := 0.7 (unless kept), roughness := 0.3, metallic := 0.  `mat` = dict of CUDA
    tensors albedo (H,W,3), roughness (H,W,1), metallic (H,W,1), mask (H,W) bool, bg (H,W,3), envmap (He,We,3)."""
    mask = mat["mask"].bool()
    albedo, roughness, metallic = mat["albedo"].clone(), mat["roughness"].clone(), mat["metallic"].clone()
    if not keep_albedo_color:
        albedo[mask] = 0.7
    roughness[mask] = roughness[mask] * 0 + 0.3
    metallic[mask] = metallic[mask] * 0.0
    return albedo, roughness, metallic


def transprancy_edit(scene, mat, ior, keep_albedo_color, specTrans, n_iter=10, spp=64):
    """Average of `n_iter` renders with seeds 0..n_iter-1 (trans_edit.py:40-43) of the scene whose shape carries the TransBSDF
    (scene.set_bsdf({'name': 'TransBSDF', ...})).  Returns the (H,W,3) linear image; file output is the caller's business."""
    if scene.trans is None:
        raise ValueError("scene does not carry a TransBSDF: build it with bsdf={'name': 'TransBSDF', 'ior': ..., 'keep_albedo_color': ...}")
    albedo, roughness, metallic = edit_materials(mat, keep_albedo_color)
    p = traverse(scene)
    p["shape.bsdf.a"] = albedo
    p["shape.bsdf.r"] = roughness
    p["shape.bsdf.m"] = metallic
    p["emitter.data"] = mat["envmap"]
    p["shape.bsdf.bg"] = mat["bg"]
    p["shape.bsdf.mask"] = mat["mask"].float() >= 1
    p["shape.bsdf.specTrans"] = specTrans
    p["shape.bsdf.ior"] = ior
    p.update()
    acc = None
    with torch.no_grad():
        for i in range(n_iter):
            img = render(scene, spp=spp, seed=i)
            acc = img if acc is None else acc + img
    return acc / n_iter
