"""Pin experiment (CPU, needs /root/reference): re-render the reference's shipped scene output_imgs/<scene> with the
mesh-mode ORACLE and compare against the reference's own saved render best_results/rendered_img.exr (written by
SaveBest at inverse_img_w_mi.py:545 = linear_to_srgb(mi.render(...) * gt.mean()/pred.mean()), Mitsuba cuda_ad_rgb,
spp 64, max_depth 4, seed = np.random.randint(0, 1000)).  Because the sampler is seeded per lane from (seed, lane), a
crop of the image can be searched over all 1000 seeds cheaply: the right seed reproduces the reference's noise pattern.
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")
from oracle import oracle as orc  # noqa: E402
from materialist_b200.camera import Camera  # noqa: E402  (host-side camera maths only; no GPU needed)


def read_ply(path):
    with open(path, "rb") as f:
        header = b""
        while not header.endswith(b"end_header\n"):
            header += f.readline()
        nv = nf = 0
        for l in header.decode().splitlines():
            t = l.split()
            if t[:2] == ["element", "vertex"]: nv = int(t[2])
            if t[:2] == ["element", "face"]: nf = int(t[2])
        v = np.frombuffer(f.read(nv * 24), dtype="<f8").reshape(nv, 3)
        fd = np.frombuffer(f.read(nf * 13), dtype=np.dtype([("n", "u1"), ("i", "<u4", 3)]))
    return v.astype(np.float32), fd["i"].astype(np.int32)


def read_img(path):
    import cv2
    img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    if img.ndim == 3:
        img = img[..., [2, 1, 0]]
    return np.ascontiguousarray(img.astype(np.float32))


def make_cfg(cam, d, H, W, spp, seed, row0, rows, flags, max_depth=4, use_mesh_normal=1):
    c = orc.Cfg()
    c.H, c.W, c.spp, c.max_depth = H, W, spp, max_depth
    c.seed = seed & 0xFFFFFFFF
    c.filter = orc.FILTER_GAUSSIAN
    c.flags = flags
    c.use_mesh_normal = use_mesh_normal
    c.row0, c.rows = row0, rows
    c.view[:] = cam.view_matrix.reshape(-1).tolist()
    c.proj[:] = cam.proj_matrix.reshape(-1).tolist()
    c.cam_to_world[:] = cam.to_world.astype(np.float32).reshape(-1).tolist()
    c.tan_half_fov_x = cam.tan_half_fov_x
    c.env_u_shift = float(np.float32(0.5) / np.float32(d.res_x - 1)) if flags & orc.FLAG_ENV_HALF_TEXEL else 0.0
    return c


def load_scene(name):
    base = os.path.join("/root/reference/output_imgs", name)
    br = os.path.join(base, "best_results")
    S = {}
    S["verts"], S["tris"] = read_ply(os.path.join(base, f"{name}.ply"))
    S["a"] = read_img(os.path.join(br, "albedo.exr"))[..., :3].copy()
    r = read_img(os.path.join(br, "roughness.exr")); m = read_img(os.path.join(br, "metallic.exr"))
    S["r"] = np.ascontiguousarray((r[..., :1] if r.ndim == 3 else r[..., None]))
    S["m"] = np.ascontiguousarray((m[..., :1] if m.ndim == 3 else m[..., None]))
    S["env"] = read_img(os.path.join(br, "envmap.hdr"))
    S["ref_srgb"] = read_img(os.path.join(br, "rendered_img.exr"))[..., :3]
    S["gt"] = read_img(os.path.join(base, "gt_image.exr"))[..., :3]
    return S


def srgb_inv(x):
    """inverse of myutils linear_to_srgb"""
    return x


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="indoor")
    ap.add_argument("--row0", type=int, default=200); ap.add_argument("--rows", type=int, default=8)
    ap.add_argument("--seeds", default="0:16"); ap.add_argument("--spp", type=int, default=64)
    ap.add_argument("--flags", type=int, default=orc.FLAG_WO_WORLD_QUIRK | orc.FLAG_ROW_STRIDE_H | orc.FLAG_ENV_HALF_TEXEL)
    ap.add_argument("--max-depth", type=int, default=4); ap.add_argument("--face-normals", type=int, default=0); ap.add_argument("--variant", type=int, default=0)
    args = ap.parse_args()
    S = load_scene(args.scene)
    for k, v in S.items():
        print(k, v.shape, v.dtype, float(v.min()), float(v.max()), float(v.mean()))
    O = orc.Oracle()
    O.lib.mbo_set_variant(args.variant)
    H = W = 512
    cam = Camera(width=W, height=H)
    env_int, hier, d = O.env_prepare(S["env"], orc.ENV_ASSIGNED)
    t0 = time.time(); mesh = O.mesh_create(S["verts"], S["tris"], face_normals=args.face_normals); print("bvh build %.2fs" % (time.time() - t0))
    lo, hi = (int(x) for x in args.seeds.split(":"))
    ref = S["ref_srgb"]
    r0, r1 = args.row0, args.row0 + args.rows
    for seed in range(lo, hi):
        cfg = make_cfg(cam, d, H, W, args.spp, seed, r0, args.rows, args.flags, args.max_depth)
        t0 = time.time()
        img, st = O.mesh_render_fwd(cfg, mesh, S["a"], S["r"], S["m"], None, env_int, hier, d, want_stats=True)
        dt = time.time() - t0
        rr = ref[r0:r1]
        e_lin = np.linalg.norm(img - rr) / np.linalg.norm(rr)
        k = (img * rr).sum() / (img * img).sum()
        e_lin_s = np.linalg.norm(img * k - rr) / np.linalg.norm(rr)
        sr = np.maximum(img, 0) ** (1 / 2.2); ks = (sr * rr).sum() / (sr * sr).sum()
        e_srgb = np.linalg.norm(sr * ks - rr) / np.linalg.norm(rr)
        print(f"seed {seed}: {dt:.2f}s  mean {img.mean():.4f} ref {rr.mean():.4f}  rel-L2 linear {e_lin:.4f} scaled(k={k:.3f}) {e_lin_s:.4f}  srgb-scaled(k={ks:.3f}) {e_srgb:.4f}  stats {st.tolist()}")
