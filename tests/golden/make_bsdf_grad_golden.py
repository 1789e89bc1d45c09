#!/usr/bin/env python
"""Adjoint pin from the REFERENCE'S OWN SOURCE: d f / d(albedo, roughness, metallic, normal) of MatDiffBSDF.eval_pdf
(myutils/mi_plugin.py:1372-1427, :1449-1460), per lane.

mitsuba / drjit are not installable here, so Mitsuba's reverse-mode AD of the plugin cannot be run.  Instead the plugin's
Dr.Jit-typed source is executed from /root/reference on the numpy stand-ins of drjit_np_shim.py switched to FLOAT64
(MB_SHIM_F64=1), and the Jacobian of its rgb value with respect to the material it gathers is taken by central finite
differences (step 1e-6 in double: truncation + rounding ~1e-9 relative).  What Mitsuba's AD would return for a cotangent w
is J^T w — the quantity oracle/mb_oracle.c::eval_brdf_grad and mb200_device.cuh::eval_brdf_grad compute by hand; the pdf is
detached in the path integrator, so only f is differentiated.

Run:  python tests/golden/make_bsdf_grad_golden.py   -> matdiff_bsdf_grad.npz next to this file.
"""
import os
import sys

os.environ["MB_SHIM_F64"] = "1"
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import drjit_np_shim as shim  # noqa: E402
from make_bsdf_golden import REF, lanes, load_reference_plugin  # noqa: E402


def main():
    mp, dr, mi = load_reference_plugin()
    assert shim.F32 is np.float64
    cam_json = os.path.join(REF, "myutils", "default_cam.json")
    H = W = 512
    L = 2048
    ln, maps = lanes(L, H, W, 4321, cam_json)
    # the normal map the plugin gathers when use_mesh_normal is False: a perturbed, re-normalised copy of the lane normals
    # written at each lane's own texel would collide between lanes, so lanes use a constant-per-lane normal through a map that
    # is rebuilt per evaluation: normal map = image whose texel (x, y) of lane i holds n_i (lanes sit on distinct pixels below)
    si = shim.FakeSI(ln["p"], ln["n"], ln["wi"])
    wo_local = si.to_local(mi.Vector3f(ln["wo"]))
    out = {}
    for tag, use_mesh_normal in (("mesh", True), ("nmap", False)):
        b = mp.MatDiffBSDF(mi.Properties(cam_meta=cam_json, use_mesh_normal=use_mesh_normal))
        sc = mp.mi_world_to_screen(si.p, b.view_matrix, b.persp_proj_matx, b.width, b.height).numpy()
        flat = (np.floor(sc[:, 0]).astype(np.int64) + np.floor(sc[:, 1]).astype(np.int64) * H)
        # one lane per texel: drop lanes that share a texel so that a per-lane perturbation of the maps is well defined
        _, first = np.unique(flat, return_index=True)
        keep = np.zeros(L, bool); keep[first] = True
        a0 = maps["a"].astype(np.float64); r0 = maps["r"].astype(np.float64); m0 = maps["m"].astype(np.float64)
        n0 = np.zeros((H, W, 3)); n0.reshape(-1, 3)[:] = (0.0, 0.0, 1.0)
        nl = ln["n"].astype(np.float64) + 0.15 * np.random.RandomState(7).randn(L, 3)
        nl /= np.linalg.norm(nl, axis=-1, keepdims=True)
        n0.reshape(-1, 3)[flat[keep]] = nl[keep]

        def f_of(a, r, m, n):
            b.a = mi.TensorXf(a); b.r = mi.TensorXf(r); b.m = mi.TensorXf(m); b.n = mi.TensorXf(n)
            f, pdf = b.eval_pdf(None, si, wo_local)
            return f.numpy().astype(np.float64)

        h = 1e-6
        f0 = f_of(a0, r0, m0, n0)
        names = ["a0", "a1", "a2", "r", "m"] + ([] if use_mesh_normal else ["n0", "n1", "n2"])
        J = np.zeros((L, len(names), 3))
        for k, nm in enumerate(names):
            def pert(sign):
                a, r, m, n = a0.copy(), r0.copy(), m0.copy(), n0.copy()
                if nm[0] == "a": a[..., int(nm[1])] += sign * h
                elif nm == "r": r += sign * h
                elif nm == "m": m += sign * h
                else: n[..., int(nm[1])] += sign * h
                return f_of(a, r, m, n)
            J[:, k] = (pert(+1) - pert(-1)) / (2 * h)
        out[f"{tag}_J"] = J[keep].astype(np.float32)
        out[f"{tag}_f"] = f0[keep].astype(np.float32)
        out[f"{tag}_keep"] = np.flatnonzero(keep).astype(np.int32)
        if not use_mesh_normal:
            out["nmap_normals"] = nl[keep].astype(np.float32)
            out["nmap_flat"] = flat[keep].astype(np.int64)
    blk = maps["blocks"]
    out["wo_world_used"] = si.to_world(wo_local).numpy().astype(np.float32)
    out["wi_world_used"] = si.to_world(si.wi).numpy().astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "matdiff_bsdf_grad.npz"), H=H, W=W, **ln, a_block=blk["a"], r_block=blk["r"], m_block=blk["m"], **out)
    print("wrote matdiff_bsdf_grad.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
