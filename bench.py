#!/usr/bin/env python
"""bench.py — headline benchmark of the differentiable envmap-shading path (BASELINE.json).

A "step" is ONE inverse-optimisation iteration of the BRDF phase (`--model_name=none --opt_src=arm
--opt_order=arm`, inverse_img_w_mi.py:368-446): forward render (seed) -> ratio / sRGB / MSE+L1+aux loss ->
adjoint render (seed_grad) -> Adam step.  One *sample* = one (pixel, spp-index) path shaded forward AND in the
adjoint render, so  Gsamples/s = H*W*spp / t_step / 1e9  and  iters/s = 1 / t_step.

  python bench.py --gpus N --steps K --warmup W [--workload c2|c5] [--impl reference]

N>1 is launched by torchrun (one rank per GPU, NCCL): rows are sharded, per-GPU work is fixed ("weak": each rank
owns a C2-sized 512-row slab of a (512*N) x 512 image); `--workload c5` is the fixed 4K image (strong scaling).
`--impl reference` times the reference's CPU path — restated by oracle/ because mitsuba==3.5.2 is not installable
here — on the box's host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOADS = {
    # name: (H per rank or total, W, spp, He, We, scaling)
    "c2": dict(H=512, W=512, spp=64, He=128, We=256, scaling="weak",
               desc="inverse_img_w_mi.py --model_name=none --opt_src=arm --opt_order=arm, synthetic 512x512 G-buffer, 64 spp, 256x128 envmap"),
    "c5": dict(H=2160, W=3840, spp=256, He=1024, We=2048, scaling="strong",
               desc="synthetic 4K (3840x2160) G-buffer inverse optimisation, 2048x1024 envmap, 256 spp, rows sharded"),
    "c3": dict(H=768, W=1024, spp=64, He=16, We=32, scaling="weak", pos_mlp=True,
               desc="inverse_img_w_mi.py --model_name=pos_mlp --opt_src=a --opt_order='rm a', synthetic 1024x768 G-buffer, 64 spp, 16x32 envmap, brdf_net = PosMLP"),
    "c4": dict(H=1080, W=1920, spp=32, He=512, We=1024, scaling="weak", rolling=True,
               desc="render_final.py --mode=rolling: rotated-envmap relights of a synthetic 1080p material set, 1024x512 envmap, 32 spp, forward only, frames sharded over the GPUs (replicas, no collective)"),
    "c2m": dict(H=512, W=512, spp=64, He=128, We=256, scaling="weak", mesh=True,
                desc="inverse_img_w_mi.py --model_name=none --opt_src=arm --opt_order=arm with the scene TRACED as the reference does (521 k-triangle "
                     "synthetic height-field mesh, per-sample hits, shadow rays, max_depth 4 bounces), 512x512, 64 spp, 256x128 envmap"),
    "c1": dict(H=512, W=512, spp=64, He=16, We=32, scaling="weak", real=True,
               desc="render_final.py --save_name=indoor --mode=real: the shipped output_imgs/indoor scene (522 220-face PLY, optimised maps, "
                    "16x32 envmap; tests/golden/indoor_pin.npz) TRACED as the reference does (path max_depth 4), 64 spp per mi.render call, forward only; "
                    "seeds sharded over the GPUs (replicas, no collective)"),
    "c1t": dict(H=512, W=512, spp=64, He=16, We=32, scaling="weak", real=True, trans=True,
                desc="trans_edit.py --save_name=indoor --ior 1.2 --specTrans 0.4: the shipped indoor scene TRACED with the TransBSDF editing plugin "
                     "(the scene's shipped edit mask and background image, trans_edit.py's material overrides), 64 spp per mi.render call, forward only"),
    "tinym": dict(H=64, W=64, spp=32, He=16, We=32, scaling="weak", mesh=True, desc="tiny self-test workload, mesh mode"),
    "tiny": dict(H=64, W=64, spp=32, He=16, We=32, scaling="weak", desc="tiny self-test workload"),
}
METRIC = "fwd+adjoint shaded samples/s (inverse-optimisation iteration)"
UNIT = "Gsamples/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-c5", action="store_true", help="skip the C5 strong-scaling rider of the default (c2) run")
    ap.add_argument("--optimizer", default="fused", choices=["fused", "autograd"],
                    help="fused: csrc/mb200_optim.cu loss + Adam kernels (9 launches/step); autograd: torch ops + torch.optim.Adam")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True); self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ workload
def build_case(wl, world):
    """Numpy/torch CPU inputs of the workload (identical bits for the CPU oracle and the GPU path)."""
    import numpy as np
    import torch
    from materialist_b200 import synthetic
    from materialist_b200.camera import Camera
    H = wl["H"] * (world if (wl["scaling"] == "weak" and not wl.get("rolling")) else 1)
    W = wl["W"]
    cam = Camera(width=W, height=H)
    mesh = None
    if wl.get("mesh"):
        mesh = synthetic.grid_mesh(synthetic.bumpy_positions(H, W, cam))
    pos, nrm, valid = synthetic.gbuffer(H, W, cam)
    a, r, m = synthetic.materials(H, W, seed_base=1)
    a2, r2, m2 = synthetic.materials(H, W, seed_base=5)       # the material set behind gt_image
    env = synthetic.envmap(wl["He"], wl["We"], seed=4)
    return dict(mesh=mesh, H=H, W=W, cam=cam, pos=pos, nrm=nrm, valid=valid, a=a, r=r, m=m, a2=a2, r2=r2, m2=m2, env=env)


def alg_bytes_per_pixel():
    # fwd 96 B + bwd 88 B per pixel (G-buffer, a/r/m, RGBW, develop, grad_out, W, material grads)
    return 184.0


# ------------------------------------------------------------------------------------------------ reference arm
def _host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def _stub_product_package():
    """The reference arm must not load the product: `materialist_b200/__init__` dlopens libmaterialist_b200.so.  Register an empty
    package object whose __path__ points at the directory, so that the two lib-free host modules the workload generators live in
    (materialist_b200.camera, materialist_b200.synthetic) import without running __init__."""
    import types
    if "materialist_b200" not in sys.modules:
        pkg = types.ModuleType("materialist_b200"); pkg.__path__ = [os.path.join(ROOT, "materialist_b200")]
        sys.modules["materialist_b200"] = pkg


REF_FLAGS = 1 | 2 | 4          # oracle.FLAG_WO_WORLD_QUIRK | FLAG_ROW_STRIDE_H | FLAG_ENV_HALF_TEXEL


def try_mitsuba_reference(args, wl):
    """when `import mitsuba` works AND the reference tree is reachable (MATERIALIST_REF), time the reference's
    own path — Mitsuba llvm_ad_rgb with its MatDiffBSDF plugin — instead of the restatement.  Neither exists in the build image or on
    the bench boxes of this pool (no wheel, no network, /root/reference is not shipped), so this returns None there."""
    ref = os.environ.get("MATERIALIST_REF", "/root/reference")
    try:
        import mitsuba as mi
        import drjit as dr
    except Exception:
        return None
    if not os.path.isdir(ref) or wl.get("real") or wl.get("mesh") or wl.get("rolling") or wl.get("pos_mlp"):
        return None
    try:
        import numpy as np
        mi.set_variant("llvm_ad_rgb")
        sys.path.insert(0, ref)
        import myutils.mi_plugin as mp
        mi.register_bsdf("MatDiffBSDF", lambda props: mp.MatDiffBSDF(props))
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import upstream_check as uc                                   # scene construction shared with the guarded parity check
        scene, params, leaves = uc.build_reference_scene(mi, dr, mp, ref, wl["He"], wl["We"])
        spp = wl["spp"]

        def step(i):
            img = mi.render(scene, params, spp=spp, seed=i)
            dr.backward(dr.sum(img))
            for t in leaves:
                dr.grad(t)
        for i in range(args.warmup):
            step(i)
        t0 = time.perf_counter()
        for i in range(args.steps):
            step(100 + i)
        dt = (time.perf_counter() - t0) / args.steps
        return {"dt": dt, "samples": 512 * 512 * spp, "kind": "mitsuba", "sample": f"whole 512x512 image, {spp} spp, mi.render + dr.backward (llvm_ad_rgb)"}
    except Exception as e:                                            # API drift: fall back to the restatement, say why
        sys.stderr.write(f"[bench] mitsuba reference arm failed ({e!r}); using the oracle\n")
        return None


def run_reference(args, wl):
    """The reference's CPU path (restated: oracle/) on the host cores, on a bounded row-sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = _host_threads()
    os.environ["OMP_NUM_THREADS"] = str(cores)        # explicit: torchrun exports OMP_NUM_THREADS=1 to its workers (must precede dlopen)
    os.environ.pop("OMP_PROC_BIND", None)
    _stub_product_package()
    import numpy as np
    from oracle import oracle as orc
    O = orc.Oracle()
    cores = int(O.lib.mbo_omp_max_threads())          # what the OpenMP runtime will actually use
    mts = try_mitsuba_reference(args, wl)
    if mts is not None:
        val = mts["samples"] / mts["dt"] / 1e9
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": mts["dt"] * 1e3, "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None,
                          "dtype": "f32", "data": "synthetic", "config": {"workload": wl["desc"], "sample": mts["sample"]},
                          "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "mitsuba", "sample": mts["sample"]},
                          "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)
        return
    if wl.get("real"):                                             # C1: forward render of the shipped scene, traced (as run_real)
        from test_reference_render_pin import pin_cfg             # (tests/: imports oracle + the lib-free camera module only)
        g = np.load(os.path.join(ROOT, "tests", "golden", "indoor_pin.npz"))
        om = O.mesh_create(g["verts"], g["tris"])
        env_int, hier, d = O.env_prepare(g["env"], orc.ENV_ASSIGNED)
        rows, row0, W, spp = 16, 248, 512, wl["spp"]
        for i in range(args.warmup):
            O.mesh_render_fwd(pin_cfg(d, i, row0, rows), om, g["a"], g["r"], g["m"], None, env_int, hier, d)
        t0 = time.perf_counter()
        for i in range(args.steps):
            O.mesh_render_fwd(pin_cfg(d, 100 + i, row0, rows), om, g["a"], g["r"], g["m"], None, env_int, hier, d)
        dt = (time.perf_counter() - t0) / args.steps
        val = rows * W * spp / dt / 1e9
        sample = f"rows [{row0},{row0 + rows}) of the 512x512 image, {spp} spp, forward render"
        print(json.dumps({"impl": "reference", "metric": "forward shaded samples/s (relight of the shipped scene, traced)", "value": val, "unit": UNIT,
                          "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
                          "scaling": wl["scaling"], "vs_baseline": None, "dtype": "f32", "data": "shipped scene fixture (tests/golden/indoor_pin.npz)",
                          "config": {"workload": wl["desc"], "sample": sample},
                          "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                                           "note": "restated Mitsuba path (oracle/, C + OpenMP), not Mitsuba itself: mitsuba==3.5.2 is not installable here"},
                          "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)
        return
    case = build_case(wl, 1)
    H, W, spp = case["H"], case["W"], wl["spp"]
    rows = max(4, min(H, int(2.0e6 // (W * spp)) or 1))           # ~2 M samples per step
    if case["mesh"] is not None:
        rows = max(4, rows // 8)                                   # traced paths are ~10x the work per sample
    row0 = (H - rows) // 2
    env_int, hier, d = O.env_prepare(case["env"].numpy(), orc.ENV_FILE)
    gpos = np.ascontiguousarray(np.concatenate([case["pos"], case["valid"][..., None].astype(np.float32)], -1))
    gnrm = np.ascontiguousarray(np.concatenate([case["nrm"], np.zeros((H, W, 1), np.float32)], -1))
    a, r, m = (case[k].numpy() for k in ("a", "r", "m"))
    cam = case["cam"]

    def cfg(seed):
        c = orc.Cfg()
        c.H, c.W, c.spp, c.max_depth, c.seed = H, W, spp, 4, seed
        c.filter, c.flags, c.use_mesh_normal = orc.FILTER_GAUSSIAN, (REF_FLAGS if H == W else REF_FLAGS & ~orc.FLAG_ROW_STRIDE_H), 1
        c.row0, c.rows = row0, rows
        c.view[:] = cam.view_matrix.reshape(-1).tolist(); c.proj[:] = cam.proj_matrix.reshape(-1).tolist()
        c.cam_to_world[:] = cam.to_world.astype(np.float32).reshape(-1).tolist(); c.tan_half_fov_x = cam.tan_half_fov_x
        c.env_u_shift = float(np.float32(0.5) / np.float32(d.res_x - 1))
        return c

    G = np.ones((H, W, 3), np.float32)
    om = O.mesh_create(*case["mesh"]) if case["mesh"] is not None else None

    def step(i):
        if om is not None:
            O.mesh_render_fwd(cfg(i), om, a, r, m, None, env_int, hier, d)
            O.mesh_render_bwd(cfg(O.seed_grad(i)), om, a, r, m, None, env_int, hier, d, G, want=("a", "r", "m"))
            return
        O.render_fwd(cfg(i), gpos, gnrm, a, r, m, None, env_int, hier, d)
        O.render_bwd(cfg(O.seed_grad(i)), gpos, gnrm, a, r, m, None, env_int, hier, d, G, want=("a", "r", "m"))

    for i in range(args.warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(100 + i)
    dt = (time.perf_counter() - t0) / args.steps
    val = rows * W * spp / dt / 1e9
    sample = f"rows [{row0},{row0 + rows}) of the {H}x{W} image, {spp} spp, fwd + adjoint render per step (no loss/optimiser)"
    try:
        lib_loaded = "libmaterialist_b200" in open("/proc/self/maps").read()
    except Exception:
        lib_loaded = None
    out = {"impl": "reference", "product_lib_loaded": lib_loaded, "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None,
           "dtype": "f32", "data": "synthetic", "config": {"workload": wl["desc"], "sample": sample},
           "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                            "note": "restated Mitsuba-llvm path (oracle/, C + OpenMP), not Mitsuba itself: mitsuba==3.5.2 is not installable here"},
           "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "iters_per_s_full_image": val * 1e9 / (H * W * spp)}
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------ rolling relight (C4)
def run_rolling(args, wl, case, scene, dev, world, rank, local):
    """One step = one frame per rank: roll the envmap by 1 degree (render_final.py:290-298), rebuild the sampling hierarchy ON
    THE GPU (the reference migrates to the host for that), forward render + film develop.  Frames are independent: no collective."""
    import torch
    import torch.distributed as dist
    import materialist_b200 as mb
    from materialist_b200 import _abi
    from materialist_b200.inverse_img_w_mi import rotate_envmap
    H, W, spp = case["H"], case["W"], wl["spp"]
    scene.set_shard(0, H)
    scene.a, scene.r, scene.m = (case[k].to(dev) for k in ("a", "r", "m"))
    env = case["env"].to(dev)

    def frame(k):
        scene.set_envmap(rotate_envmap(env, float(k * world + rank)), _abi.ENV_FILE)
        return mb.render(scene, spp=spp, seed=0)

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(); torch.cuda.synchronize()

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    for k in range(args.warmup):
        frame(k)
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(args.steps):
        img = frame(100 + k)
    e1.record(); sync()
    tt = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_step = float(tt.item()) / args.steps / 1e3
    # e2e: envmap rolled on the HOST and uploaded per frame, image downloaded per frame
    henv = case["env"].pin_memory(); himg = torch.empty(H, W, 3).pin_memory()
    def frame_e2e(k):
        scene.set_envmap(torch.roll(henv, shifts=int((k * world + rank) / 360.0 * henv.shape[1]), dims=1).pin_memory().to(dev, non_blocking=True), _abi.ENV_FILE)
        himg.copy_(mb.render(scene, spp=spp, seed=0), non_blocking=True)
    for k in range(3):
        frame_e2e(k)
    sync(); e0.record()
    for k in range(args.steps):
        frame_e2e(200 + k)
    e1.record(); sync()
    te = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    t_e = float(te.item()) / args.steps / 1e3
    clk = clocks.stop() if rank == 0 else None
    if rank == 0:
        samples = H * W * spp * world
        print(json.dumps({"metric": "forward shaded samples/s (rolling-envmap relight, hierarchy rebuilt per frame)", "value": samples / t_step / 1e9,
                          "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_step * 1e3,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": {"workload": wl["desc"], "image": [H, W], "spp": spp, "envmap": [wl["He"], wl["We"]], "filter": "gaussian",
                                     "parallelism": f"one frame per GPU per step, {world} GPU(s)"},
                          "frames_per_s": world / t_step,
                          "e2e": {"value": samples / t_e / 1e9, "unit": UNIT, "h2d_bytes_per_step": henv.numel() * 4, "d2h_bytes_per_step": himg.numel() * 4,
                                  "ms_per_step": t_e * 1e3, "bytes_are": "per rank"},
                          "gpu_launches": 8 * args.steps, "gpu_launches_note": "per frame: 6 env_* (ingest + Hierarchical2D build), shade_fwd, film_develop",
                          "clocks": clk, "image_mean": float(img.mean().item())}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_real(args, wl, dev, world, rank, local):
    """C1: one step = one `mi.render(scene, spp=64, seed=i)` of render_final.py:193-196 on the shipped indoor scene, traced."""
    import numpy as np
    import torch
    import torch.distributed as dist
    import materialist_b200 as mb
    from materialist_b200 import renderop
    g = np.load(os.path.join(ROOT, "tests", "golden", "indoor_pin.npz"))
    H = W = 512; spp = wl["spp"]
    cam = mb.Camera(width=W, height=H)
    scene = mb.Scene.from_mesh(g["verts"], g["tris"], cam, device=dev, envmap=torch.from_numpy(g["env"]), use_mesh_normal=True, max_depth=4)
    scene.set_envmap(torch.from_numpy(g["env"]), mb._abi.ENV_ASSIGNED)
    if wl.get("trans"):
        scene.set_bsdf({"name": "TransBSDF", "ior": 1.2, "keep_albedo_color": False})
    p = mb.traverse(scene)
    p["shape.bsdf.a"], p["shape.bsdf.r"], p["shape.bsdf.m"] = (torch.from_numpy(g[k]).to(dev) for k in ("a", "r", "m"))
    if wl.get("trans"):
        # the scene's own edit mask and background image (output_imgs/indoor/best_results/{mask,bg}.png, shipped by the reference)
        from materialist_b200.imageio import read_bitmap
        import torch.nn.functional as NF
        gold = os.path.join(ROOT, "tests", "golden")
        mk = read_bitmap(os.path.join(gold, "indoor_mask.png"))[..., 0] != 0
        bg = torch.from_numpy(np.ascontiguousarray(read_bitmap(os.path.join(gold, "indoor_bg.png"))[..., :3]))
        bg = NF.interpolate(bg[None].permute(0, 3, 1, 2), size=(H, W), mode="bilinear", align_corners=True)[0].permute(1, 2, 0).contiguous()
        from materialist_b200.trans_edit import edit_materials
        ea, er, em = edit_materials({"albedo": p["shape.bsdf.a"], "roughness": p["shape.bsdf.r"], "metallic": p["shape.bsdf.m"],
                                     "mask": torch.from_numpy(mk).to(dev)}, False)
        p["shape.bsdf.a"], p["shape.bsdf.r"], p["shape.bsdf.m"] = ea, er, em
        p["shape.bsdf.mask"] = torch.from_numpy(mk).to(dev)
        p["shape.bsdf.bg"] = bg.to(dev)
        p["shape.bsdf.specTrans"] = 0.4
    p.update()

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(); torch.cuda.synchronize()

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    with torch.no_grad():
        for k in range(args.warmup):
            mb.render(scene, spp=spp, seed=k)
        sync()
        renderop.KERNEL_EVENTS = []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(args.steps):
            img = mb.render(scene, spp=spp, seed=(k * world + rank))
        e1.record(); sync()
        ev = renderop.KERNEL_EVENTS; renderop.KERNEL_EVENTS = None
        t_kernel = sum(a.elapsed_time(b) for _, a, b in ev) / max(len(ev), 1)
        tt = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_step = float(tt.item()) / args.steps / 1e3
        himg = torch.empty(H, W, 3).pin_memory()
        for k in range(2):
            himg.copy_(mb.render(scene, spp=spp, seed=k), non_blocking=True)
        sync(); e0.record()
        for k in range(args.steps):
            himg.copy_(mb.render(scene, spp=spp, seed=(k * world + rank)), non_blocking=True)
        e1.record(); sync()
        te = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        t_e = float(te.item()) / args.steps / 1e3
    clk = clocks.stop() if rank == 0 else None
    if rank == 0:
        samples = H * W * spp * world
        cpu = None
        if not args.no_cpu_baseline and not wl.get("trans"):
            from oracle import oracle as orc
            from test_reference_render_pin import pin_cfg
            O = orc.Oracle(); om = O.mesh_create(g["verts"], g["tris"])
            env_int, hier, d = O.env_prepare(g["env"], orc.ENV_ASSIGNED)
            rows = 16; t0 = time.time()
            O.mesh_render_fwd(pin_cfg(d, 0, 248, rows), om, g["a"], g["r"], g["m"], None, env_int, hier, d)
            dt = time.time() - t0; O.mesh_destroy(om)
            cpu = {"value": rows * W * spp / dt / 1e9, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                   "sample": f"rows [248,{248 + rows}) of the 512x512 image, 64 spp, forward render",
                   "note": "restated Mitsuba path (oracle/, C + OpenMP), not Mitsuba itself: mitsuba==3.5.2 is not installable here"}
        try:
            peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0))
        except Exception:
            peak = 6650.0
        alg = H * W * 64.0 + g["tris"].shape[0] * (96.0 + 24.0 * 4 / 3)      # film/material bytes per pixel + triangles and BVH once
        print(json.dumps({"metric": "forward shaded samples/s (relight of the shipped scene, traced)", "value": samples / t_step / 1e9, "unit": UNIT,
                          "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_step * 1e3, "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "shipped scene fixture (tests/golden/indoor_pin.npz)",
                          "config": {"workload": wl["desc"], "image": [H, W], "spp": spp, "envmap": [16, 32], "filter": "gaussian", "max_depth": 4,
                                     "triangles": int(g["tris"].shape[0]), "parallelism": f"one seed per GPU per step, {world} GPU(s)",
                                     "l2": "BVH + triangles (67 MB) are L2-resident by design; 105 MB of film partials stream per step"},
                          "renders_per_s": world / t_step,
                          "e2e": {"value": samples / t_e / 1e9, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": himg.numel() * 4,
                                  "ms_per_step": t_e * 1e3, "bytes_are": "per rank", "what": "mi.render(scene, spp, seed) + download of the image; inputs are scene state, as in render_final.py"},
                          "gpu_launches": 2 * args.steps, "gpu_launches_note": "per render: mesh_fwd, film_develop",
                          "kernel_ms": {("mesh_fwd_trans" if wl.get("trans") else "mesh_fwd"): t_kernel},
                          "roofline": {"bound": "hbm", "kernel": "mesh_fwd", "achieved": alg / (t_kernel * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                       "frac": alg / (t_kernel * 1e-3) / 1e9 / peak, "traffic": None, "alg_bytes_per_launch": alg,
                                       "note": "traced paths are bound by L2/L1 latency on dependent BVH loads (profiles/r1u), not by HBM; reported for the contract"},
                          "cpu_baseline": cpu, "clocks": clk, "image_mean": float(img.mean().item())}), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ measurement helpers
def _max_over_ranks(ms, dev, world):
    import torch
    import torch.distributed as dist
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _sync(world):
    import torch
    import torch.distributed as dist
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier(); torch.cuda.synchronize()


def _timed(fn, steps, warmup, dev, world, seed0=2000):
    """W warm-up calls, then K calls between CUDA events on the current stream, barrier + synchronize on both sides; max over ranks."""
    import torch
    for i in range(warmup):
        fn(i)
    _sync(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(seed0 + i)
    e1.record()
    _sync(world)
    return _max_over_ranks(e0.elapsed_time(e1), dev, world) / steps / 1e3


def load_ncu_counters(workload):
    """profiles/ncu_counters.json: per-kernel counters of one `ncu --set full` capture (tools/ncu_counters.py), with the commit and
    the workload it was taken on.  None when absent or taken on another workload."""
    try:
        c = json.load(open(os.path.join(ROOT, "profiles", "ncu_counters.json")))
        return c if c.get("workload") == workload else None
    except Exception:
        return None


def build_roofline(args, wl, scene, shard, dom, kavg, t_step, dev, world):
    """achieved = max(bytes_alg / BW, flops / peak) / t — both roofs reported, `bound` names the binding one.
    bytes: the algorithmic per-pixel figures of §8d.  FLOPs: EXECUTED FP32 operations per sample (2 FFMA + FADD + FMUL thread
    instructions) counted by ncu on this kernel and committed in profiles/ncu_counters.json (not an estimate); the FP32 peak is an
    FFMA loop timed in this run.  DRAM traffic and issue-slot utilisation come from the same capture."""
    import torch
    from materialist_b200 import _abi
    if not dom:
        return None, None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md 6.65 TB/s)"
    W, spp = scene.W, wl["spp"]
    npix_rank = shard.rows * W
    env_bytes = wl["He"] * (wl["We"] + 1) * 16 + scene.prepared_env()[2].total_floats * 4
    per_px = {"shade_bwd": 88.0, "shade_fwd": 96.0, "mesh_bwd": 88.0 - 32.0, "mesh_fwd": 96.0 - 32.0}.get(dom.replace("_wf", ""), 184.0)
    alg_bytes = npix_rank * per_px + env_bytes
    if scene.mesh is not None:
        alg_bytes += scene.mesh.desc.total_bytes
    t_k = kavg[dom] * 1e-3
    hbm_ach = alg_bytes / t_k / 1e9
    # FP32 peak: FFMA loop, measured now
    peak_tf = None
    try:
        buf = torch.zeros(1 << 20, device=dev)
        fl = _abi.C.c_double(0.0)
        for _ in range(2):
            _abi.check(_abi.lib.mb200_probe_ffma(_abi.ptr(buf), 4096, _abi.C.byref(fl), _abi.stream_ptr()), "probe")
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        _abi.check(_abi.lib.mb200_probe_ffma(_abi.ptr(buf), 16384, _abi.C.byref(fl), _abi.stream_ptr()), "probe")
        p1.record(); torch.cuda.synchronize()
        peak_tf = fl.value / (p0.elapsed_time(p1) * 1e-3) / 1e12
    except Exception:
        pass
    cnt = load_ncu_counters(args.workload)
    kc = (cnt or {}).get("kernels", {}).get(dom)
    fp32 = {"peak_tflops": peak_tf, "peak_source": "measured here: FFMA loop (mb200_probe_ffma)"}
    fp_frac = None
    if kc and peak_tf:
        flops = kc["flops_per_sample"] * npix_rank * spp
        ach = flops / t_k / 1e12
        fp_frac = ach / peak_tf
        fp32.update({"achieved_tflops": ach, "frac": fp_frac, "flops_per_sample_executed": kc["flops_per_sample"],
                     "thread_inst_per_sample": kc["thread_inst_per_sample"], "fp32_pipe_inst_frac_ncu": kc["fp32_pipe_inst_frac"],
                     "issue_active_frac_ncu": kc["issue_active_frac"],
                     "counters_from": {"file": "profiles/ncu_counters.json", "commit": cnt["commit"], "source": cnt.get("source")},
                     "note": "FLOPs = 2 FFMA + FADD + FMUL thread instructions executed (ncu), per sample, times the samples of this launch; "
                             "a kernel of non-contracted IEEE multiplies and adds (the bit-exact direction chain) can reach at most half of the FFMA peak"})
        # the roof this kernel actually sits under: one warp instruction per scheduler per cycle (4 schedulers per SM).  Executed warp
        # instructions per sample are ncu's count (same capture); the rate is this run's kernel time; the clock is what nvidia-smi
        # reported for this GPU.
        try:
            sms = torch.cuda.get_device_properties(dev).multi_processor_count
            mhz = float(subprocess.run(["nvidia-smi", "--query-gpu=clocks.max.sm", "--format=csv,noheader,nounits", "-i", str(dev.index or 0)],
                                       capture_output=True, text=True, timeout=10).stdout.strip().splitlines()[0])
            winst = kc["warp_inst"] * (npix_rank * spp / float(cnt["samples_per_launch"]))
            peak_i = sms * 4 * mhz * 1e6
            fp32["issue"] = {"warp_inst_per_launch": winst, "achieved_ginst_s": winst / t_k / 1e9, "peak_ginst_s": peak_i / 1e9,
                             "frac": winst / t_k / peak_i, "peak_is": f"{sms} SMs x 4 schedulers x {mhz:.0f} MHz (max SM clock)",
                             "note": "issue-slot roof: the shading kernels are bound by instruction issue, not by FP32 lanes or HBM"}
            # SURVEY 8d's algorithmic convention (1.6 k FLOP per adjoint sample, 0.8 k per forward sample) for comparison with round 1
            alg = {"shade_bwd": 1600.0, "shade_fwd": 800.0}.get(dom)
            if alg:
                fp32["frac_algorithmic_flops"] = alg * npix_rank * spp / t_k / 1e12 / peak_tf
        except Exception:
            pass
    else:
        fp32["note"] = "no ncu counters committed for this workload / kernel: FLOP rate not reported"
    hbm_frac = hbm_ach / hbm_peak
    bound_fp = fp_frac is not None and fp_frac >= hbm_frac
    roofline = {"bound": "fp32" if bound_fp else "hbm", "kernel": dom,
                "achieved": fp32.get("achieved_tflops") if bound_fp else hbm_ach, "peak": peak_tf if bound_fp else hbm_peak,
                "unit": "TFLOP/s" if bound_fp else "GB/s", "frac": fp_frac if bound_fp else hbm_frac,
                "traffic": kc["dram_bytes"] * (npix_rank * spp / float(cnt["samples_per_launch"])) if (kc and world == 1) else None,
                "hbm": {"achieved_gbs": hbm_ach, "peak_gbs": hbm_peak, "frac": hbm_frac, "alg_bytes_per_launch": alg_bytes, "peak_source": peak_src},
                "fp32": fp32, "kernel_ms": kavg[dom], "kernel_share_of_step": t_k / t_step,
                "note": "bound = the larger of the two roofline fractions (SURVEY 8d: max(bytes/BW, flops/peak)); the fused path is instruction-issue "
                        "bound at 64-256 spp (~830 FLOP/B against a ridge of ~11): the HBM fraction is ~0.3 % by construction and is reported beside it; "
                        "'bound: fp32' is the non-tensor FP32 roof (this path has no contraction to put on tensor cores)"}
    return roofline, fp32


def run_e2e_iteration(args, opt, shard, dev, world, samples_per_step):
    """SAME step as `value` — forward render, loss, adjoint render, Adam — through FusedBRDFOptimizer with the step's input image on
    the HOST: every step uploads the target image rows (pinned -> device) and reads back the step's result (the two loss scalars and
    the predicted sRGB image rows).  The material maps are the optimised parameters and stay on the device, like the weights of any
    training step."""
    import torch
    if not all(hasattr(opt, k) for k in ("gt_srgb", "pred_srgb", "sums2")):
        return None
    hgt = opt.gt_srgb.detach().cpu().pin_memory()
    hpred = torch.empty_like(hgt).pin_memory()
    hloss = torch.empty(2).pin_memory()

    def serial(seed):
        opt.gt_srgb.copy_(hgt, non_blocking=True)
        opt.step(seed)
        hpred.copy_(opt.pred_srgb, non_blocking=True)
        hloss.copy_(opt.sums2, non_blocking=True)
    t_serial = _timed(serial, args.steps, 3, dev, world)

    # the same transfers on their own streams: the upload of step i+1 runs under step i (two target buffers), the download of step
    # i's results under step i+1 (two result buffers)
    main = torch.cuda.current_stream(dev)
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    gts = [opt.gt_srgb, torch.empty_like(opt.gt_srgb)]
    ev_in = [torch.cuda.Event(), torch.cuda.Event()]; ev_used = [torch.cuda.Event(), torch.cuda.Event()]
    # results double-buffered too: step i+1 writes pred_srgb / sums2 into the other pair while step i's pair is still being read back
    preds = [opt.pred_srgb, torch.empty_like(opt.pred_srgb)]; sums = [opt.sums2, torch.zeros_like(opt.sums2)]
    ev_step = [torch.cuda.Event(), torch.cuda.Event()]; ev_out = [torch.cuda.Event(), torch.cuda.Event()]
    for e in ev_out:
        e.record(main)
    state = {"i": 0}

    def upload(slot):
        with torch.cuda.stream(s_in):
            s_in.wait_event(ev_used[slot])
            gts[slot].copy_(hgt, non_blocking=True)
            ev_in[slot].record(s_in)
    upload(0)

    def piped(seed):
        slot = state["i"] % 2; state["i"] += 1
        main.wait_event(ev_in[slot])
        opt.gt_srgb = gts[slot]; opt.pred_srgb = preds[slot]; opt.sums2 = sums[slot]
        upload(1 - slot)
        main.wait_event(ev_out[slot])                            # the results of step i-2 have left this pair of buffers
        opt.step(seed)
        ev_used[slot].record(main); ev_step[slot].record(main)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_step[slot])
            hpred.copy_(preds[slot], non_blocking=True); hloss.copy_(sums[slot], non_blocking=True)
            ev_out[slot].record(s_out)
    for i in range(3):
        piped(i)
    main.wait_event(ev_out[0]); main.wait_event(ev_out[1]); _sync(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        piped(3000 + i)
    main.wait_event(ev_out[0]); main.wait_event(ev_out[1])      # the last results are on the host when the clock stops
    e1.record(); s_in.synchronize(); s_out.synchronize(); _sync(world)
    t_pipe = _max_over_ranks(e0.elapsed_time(e1), dev, world) / args.steps / 1e3
    opt.gt_srgb = gts[0]; opt.pred_srgb = preds[0]; opt.sums2 = sums[0]
    return {"value": samples_per_step / t_pipe / 1e9, "unit": UNIT, "h2d_bytes_per_step": hgt.numel() * 4, "d2h_bytes_per_step": (hpred.numel() + 2) * 4,
            "ms_per_step": t_pipe * 1e3, "ms_per_step_copies_serialised": t_serial * 1e3, "value_copies_serialised": samples_per_step / t_serial / 1e9,
            "bytes_are": "per rank",
            "what": "the full iteration (fwd render, loss, adjoint render, Adam: the step `value` times) with its input on the host: H2D of the "
                    "target image rows every step, D2H of the loss sums and of the predicted image rows; copies on their own streams "
                    "(the figure with the copies serialised on the compute stream is listed beside it)"}


def run_e2e_operator(args, case, scene, shard, spp, dev, world, samples_per_step):
    """The operator boundary for a HOST-side optimiser: render_w_brdf forward + backward with pinned host maps in and image + material
    gradients out, every step; shard-local rows only (hostpipe).  Serial copies and copies on their own streams, both listed."""
    import torch
    import materialist_b200 as mb
    from materialist_b200.hostpipe import HostPipelinedRenderWBRDF
    H, W = scene.H, scene.W
    with scene.shard(shard.row0, shard.rows):
        pipe = HostPipelinedRenderWBRDF(scene, spp, halo_exchange=shard.halo_exchange if world > 1 else None)
        (m0, m1), (o0, o1) = pipe.map_rows, pipe.out_rows
        ha, hr, hm = (case[k][m0:m1].contiguous().pin_memory() for k in ("a", "r", "m"))
        himg = torch.empty(shard.rows, W, 3).pin_memory(); hgrad = torch.ones(shard.rows, W, 3).pin_memory()
        hga, hgr, hgm = (torch.empty(o1 - o0, W, c).pin_memory() for c in (3, 1, 1))
        da, dr_, dm = (torch.zeros(H, W, c, device=dev) for c in (3, 1, 1))

        def serial(seed):
            da[m0:m1].copy_(ha, non_blocking=True); dr_[m0:m1].copy_(hr, non_blocking=True); dm[m0:m1].copy_(hm, non_blocking=True)
            a = da.detach().requires_grad_(True); r = dr_.detach().requires_grad_(True); m = dm.detach().requires_grad_(True)
            img = mb.render(scene, spp=spp, seed=seed, albedo=a, roughness=r, metallic=m, halo_exchange=shard.halo_exchange if world > 1 else None)
            himg.copy_(img.detach(), non_blocking=True)
            img.backward(hgrad.to(dev, non_blocking=True))
            hga.copy_(a.grad[o0:o1], non_blocking=True); hgr.copy_(r.grad[o0:o1], non_blocking=True); hgm.copy_(m.grad[o0:o1], non_blocking=True)
        t_serial = _timed(serial, args.steps, 3, dev, world)
        inputs = (ha, hr, hm, hgrad)
        pipe.stage(0, *inputs)
        for i in range(3):
            pipe.step(i, i % 2, himg, hga, hgr, hgm, next_inputs=inputs)
        pipe.synchronize(); _sync(world)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            pipe.step(2000 + i, (3 + i) % 2, himg, hga, hgr, hgm, next_inputs=inputs)
        torch.cuda.current_stream().wait_event(pipe.last_out)
        e1.record(); pipe.synchronize(); _sync(world)
        t_pipe = _max_over_ranks(e0.elapsed_time(e1), dev, world) / args.steps / 1e3
    h2d = (ha.numel() + hr.numel() + hm.numel() + hgrad.numel()) * 4
    d2h = (himg.numel() + hga.numel() + hgr.numel() + hgm.numel()) * 4
    return {"value_copies_serialised": samples_per_step / t_serial / 1e9, "value_copies_overlapped": samples_per_step / t_pipe / 1e9, "unit": UNIT,
            "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "bytes_are": "per rank (own rows + 2-row film halo of the maps up; own rows down)",
            "ms_per_step_copies_serialised": t_serial * 1e3, "ms_per_step_copies_overlapped": t_pipe * 1e3,
            "what": "render_w_brdf forward + backward (no loss / optimiser: a host-side optimiser owns them) with pinned HOST buffers: H2D a/r/m + "
                    "d(loss)/d(image), D2H image + material gradients, every step"}


def run_c5_leg(args, dev, world, rank):
    """BASELINE.json configs[4] at this N: the FIXED 3840x2160 image (256 spp, 2048x1024 envmap) row-sharded over the ranks — strong
    scaling, the north-star '>= 7x at 8 GPUs on 4K G-buffers'.  Same iteration, same timing rules, fewer steps (one step is ~0.4 s of
    GPU time at N = 1)."""
    import torch
    import materialist_b200 as mb
    from materialist_b200.inverse import FusedBRDFOptimizer
    from materialist_b200.parallel import ShardContext
    wl = WORKLOADS["c5"]
    case = build_case(wl, world)
    H, W, spp = case["H"], case["W"], wl["spp"]
    shard = ShardContext(H, W, rank, world)
    scene = mb.Scene(case["pos"], case["nrm"], case["valid"], camera=case["cam"], envmap=case["env"], device=dev)
    to = lambda t: t.to(dev)
    with scene.shard(shard.row0, shard.rows):         # every rank renders its own rows of the target, then the rows are gathered
        gt_rows = mb.render(scene, spp=64, seed=999, albedo=to(case["a2"]), roughness=to(case["r2"]), metallic=to(case["m2"]))
    gt = torch.zeros(H, W, 3, device=dev)
    gt[shard.row0:shard.row0 + shard.rows] = gt_rows
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(gt)
    opt = FusedBRDFOptimizer(scene, {"albedo": to(case["a"]), "roughness": to(case["r"]), "metallic": to(case["m"])}, gt, "arm", spp=spp, shard=shard)
    steps, warmup = max(2, min(args.steps, 4)), 3
    t = _timed(lambda s: opt.step(s), steps, warmup, dev, world, seed0=1000)
    loss_last = float(opt.last["loss_mse"].item())
    exchange = "peer memory (NVLink stores + in-kernel mailboxes)" if getattr(opt, "peer", None) is not None else ("nccl" if world > 1 else None)
    if hasattr(opt, "close"):
        opt.close()
    return {"workload": wl["desc"], "scaling": "strong", "exchange": exchange, "image": [H, W], "spp": spp, "envmap": [wl["He"], wl["We"]], "n_gpus": world,
            "steps": steps, "warmup": warmup, "ms_per_step": t * 1e3, "value": H * W * spp / t / 1e9, "unit": UNIT, "iters_per_s": 1.0 / t,
            "rows_per_rank": shard.rows, "loss_mse_last": loss_last}


# ------------------------------------------------------------------------------------------------ b200 arm
def run_b200(args, wl):
    import numpy as np
    import torch
    import torch.distributed as dist
    import materialist_b200 as mb
    from materialist_b200 import _abi, renderop as mbr
    from materialist_b200.inverse import DirectBRDFOptimizer, FusedBRDFOptimizer, PosMLPBRDFOptimizer
    from materialist_b200.parallel import ShardContext

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if wl.get("real"):
        return run_real(args, wl, dev, world, rank, local)
    case = build_case(wl, world)
    H, W, spp = case["H"], case["W"], wl["spp"]
    shard = ShardContext(H, W, rank, world)
    if case["mesh"] is not None:
        scene = mb.Scene.from_mesh(case["mesh"][0], case["mesh"][1], case["cam"], device=dev, envmap=case["env"])
    else:
        scene = mb.Scene(case["pos"], case["nrm"], case["valid"], camera=case["cam"], envmap=case["env"], device=dev)
    to = lambda t: t.to(dev)
    # gt_image = render of the second material set, seed 999
    scene.set_shard(0, H)
    gt = mb.render(scene, spp=min(spp, 64), seed=999, albedo=to(case["a2"]), roughness=to(case["r2"]), metallic=to(case["m2"]))
    mat = {"albedo": to(case["a"]), "roughness": to(case["r"]), "metallic": to(case["m"])}
    if wl.get("rolling"):
        return run_rolling(args, wl, case, scene, dev, world, rank, local)
    if wl.get("pos_mlp"):
        torch.manual_seed(0)                           # same initial weights on every rank
        opt = PosMLPBRDFOptimizer(scene, mat, gt, "arm", spp=spp, shard=shard)
        with torch.no_grad():                          # lin4 is zero-initialised (mlps.py:174-176): perturb so the timed steps do real work
            opt.net.lin4.weight.normal_(0, 0.02)
        args.optimizer = "autograd+posmlp"
        args.no_e2e = True
    else:
        opt = (FusedBRDFOptimizer if args.optimizer == "fused" else DirectBRDFOptimizer)(scene, mat, gt, "arm", spp=spp, shard=shard)
    samples_per_step = H * W * spp                       # whole job (all ranks)

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    clocks = ClockSampler(local)           # sampled from the warm-up to the end of the e2e leg (all under load)
    if rank == 0:
        clocks.start()
    mbr.KERNEL_EVENTS = []                  # the warm-up runs the instrumented path too (event pool, allocator state = the timed region's)
    for i in range(args.warmup):
        opt.step(i)
    sync()
    mbr.KERNEL_EVENTS = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    e0.record()
    for i in range(args.steps):
        opt.step(1000 + i)
    e1.record()
    sync()
    t_ms = e0.elapsed_time(e1)
    kev, mbr.KERNEL_EVENTS = mbr.KERNEL_EVENTS, None
    tt = torch.tensor([t_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_step = float(tt.item()) / args.steps / 1e3
    value = samples_per_step / t_step / 1e9

    # ---- per-kernel times (CUDA events on the launching stream, inside the timed region)
    ktime = {}
    for name, a, b in kev:
        ktime.setdefault(name, []).append(a.elapsed_time(b))
    kavg = {k: sum(v) / len(v) for k, v in ktime.items()}
    shade = {k: v for k, v in kavg.items() if k != "film_weights"}
    # roofline kernel: the adjoint render kernel when it ran alone on the GPU (it is the largest kernel of the step; the forward
    # kernel's event time includes the film-weights kernel that runs beside it on a side stream, see FusedBRDFOptimizer), else the largest
    dom = "shade_bwd" if "shade_bwd" in shade else (max(shade, key=shade.get) if shade else None)
    roofline, fp32 = build_roofline(args, wl, scene, shard, dom, kavg, t_step, dev, world)

    # ---- e2e through the public API with HOST buffers
    e2e = e2e_op = None
    if not args.no_e2e and not wl.get("pos_mlp"):
        e2e = run_e2e_iteration(args, opt, shard, dev, world, samples_per_step)
        e2e_op = run_e2e_operator(args, case, scene, shard, spp, dev, world, samples_per_step)

    # ---- the north-star scaling target rides along: C5 (4K, 256 spp, 2048x1024 envmap), STRONG scaling, at this N
    c5 = None
    if args.workload == "c2" and not args.no_c5:
        try:
            c5 = run_c5_leg(args, dev, world, rank)
        except Exception as e:                                        # never lose the headline line to the rider
            c5 = {"error": repr(e)}

    clk = clocks.stop() if rank == 0 else None
    # ---- CPU baseline beside it (rank 0, N = 1 only): the oracle on a bounded row sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", args.workload,
                                "--steps", "3", "--warmup", "1"], capture_output=True, text=True, timeout=900)
            cpu = json.loads(r.stdout.strip().splitlines()[-1])["cpu_baseline"]
        except Exception as e:
            cpu = {"error": str(e)}

    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": t_step * 1e3, "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None,
               "dtype": "f32", "data": "synthetic",
               "config": {"workload": wl["desc"], "image": [H, W], "spp": spp, "envmap": [wl["He"], wl["We"]], "filter": "gaussian",
                          "max_depth": 4, "parallelism": f"rows sharded over {world} GPU(s)",
                          "l2": f"no explicit flush: each step streams {(shard.rows * W * (400 + 100 + 32 + 20 + 16 + 12 + 12 + 40)) / 1e6:.0f} MB of inputs + per-step intermediates (film tap partials 400 B/px, weight partials 100 B/px, G-buffer, maps, gradients) per rank; >126 MB L2 for C2/C5. The 0.6 MB envmap + hierarchy is L2/L1-resident by design"},
               "iters_per_s": 1.0 / t_step, "e2e": e2e, "e2e_operator": e2e_op, "c5_strong": c5, "gpu_launches": (9 if args.optimizer == "fused" else 5) * args.steps,
               "gpu_launches_note": "own kernels per step: " + ("mesh_fwd" if scene.mesh is not None else "shade_fwd") + ", film_develop, film_weights, film_adjoint, "
                                    + ("mesh_bwd" if scene.mesh is not None else "shade_bwd")
                                    + (", image_sum, loss_srgb_sums, loss_srgb_grad, adam_clamped" if args.optimizer == "fused"
                                       else " (+ ~100 torch elementwise/reduce launches for loss and Adam)"),
               "optimizer": args.optimizer,
               "kernel_ms": kavg,
               "kernel_ms_note": "CUDA events around each launch on its own stream; film_weights runs on a side stream BESIDE shade_fwd (launched behind it, "
                                 "no dependency), so those two event times include each other's work; shade_bwd runs alone",
               "roofline": roofline, "fp32": fp32, "cpu_baseline": cpu, "clocks": clk,
               "loss_mse_last": float(opt.last["loss_mse"].item()),
               "exchange": ("peer memory (NVLink stores + in-kernel mailboxes; no NCCL call in the iteration)" if getattr(opt, "peer", None) is not None
                            else ("nccl" if world > 1 else None))}
        print(json.dumps(out), flush=True)
    if hasattr(opt, "close"):
        opt.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_b200(args, wl)


if __name__ == "__main__":
    main()
