"""Synthetic inputs of the BASELINE configs: G-buffers, material maps and envmaps generated on the
CPU with explicit seeds so that the CPU oracle and the GPU path see identical bits."""
import numpy as np
import torch

from .camera import Camera


def gbuffer(H, W, camera=None, invalid_border=0):
    """Camera at the origin looking down -z (default_cam.json).  Position = point on the PIXEL-CENTRE ray at
    depth(u,v) = 20 + 4 cos(pi u) cos(pi v); normal = normalised finite-difference normal of that height field,
    flipped to face the camera.  Returns float32 numpy (pos, nrm, valid)."""
    cam = camera or Camera(width=W, height=H)
    t, aspect = cam.tan_half_fov_x, W / H
    xs = (np.arange(W, dtype=np.float64) + 0.5)
    ys = (np.arange(H, dtype=np.float64) + 0.5)
    sx, sy = np.meshgrid(xs, ys)

    def point(sx, sy):
        u, v = sx / W, sy / H
        depth = 20.0 + 4.0 * np.cos(np.pi * u) * np.cos(np.pi * v)
        # un-normalised camera-space ray with z = 1, mapped by to_world (x -> -x, z -> -z for the default camera)
        l = np.stack([(1 - 2 * sx / W) * t, (1 - 2 * sy / H) * t / aspect, np.ones_like(sx)], -1) * depth[..., None]
        return l @ cam.to_world[:3, :3].T + cam.to_world[:3, 3]

    p = point(sx, sy)
    e = 0.25
    dx = point(sx + e, sy) - point(sx - e, sy)
    dy = point(sx, sy + e) - point(sx, sy - e)
    n = np.cross(dx, dy)
    n /= np.linalg.norm(n, axis=-1, keepdims=True)
    to_cam = cam.to_world[:3, 3] - p
    flip = (n * to_cam).sum(-1) < 0
    n[flip] = -n[flip]
    valid = np.ones((H, W), dtype=bool)
    if invalid_border > 0:
        b = invalid_border
        valid[:b] = valid[-b:] = False
        valid[:, :b] = valid[:, -b:] = False
    return p.astype(np.float32), n.astype(np.float32), valid


def materials(H, W, seed_base=1):
    """albedo U[0,1]^3 (seed), roughness U[0.07,1] (seed+1), metallic U[0,1] (seed+2) — CPU torch generators."""
    def gen(seed, *shape):
        g = torch.Generator(device="cpu").manual_seed(seed)
        return torch.rand(*shape, generator=g, dtype=torch.float32)
    a = gen(seed_base, H, W, 3)
    r = gen(seed_base + 1, H, W, 1) * 0.93 + 0.07
    m = gen(seed_base + 2, H, W, 1)
    return a.contiguous(), r.contiguous(), m.contiguous()


def normal_map(nrm, seed=6, amount=0.15):
    """A perturbed unit normal map around the geometric normals (for use_mesh_normal=False cases)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    n = torch.as_tensor(nrm) + amount * torch.randn(*nrm.shape, generator=g, dtype=torch.float32)
    return torch.nn.functional.normalize(n, dim=-1).contiguous()


def envmap(He, We, seed=4, sun=2000.0):
    """0.2 + exp(N(0,1)) per texel plus one 5x5 'sun' block (envmaps/41.hdr has a 56 832 peak)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    e = 0.2 + torch.exp(torch.randn(He, We, 3, generator=g, dtype=torch.float32))
    if sun and He >= 8 and We >= 8:
        y0, x0 = He // 4, (3 * We) // 8
        e[y0:y0 + 5, x0:x0 + 5] *= sun
    return e.contiguous()


def grid_mesh(pos, valid=None):
    """Triangle mesh over a (H,W,3) position grid with the connectivity of the reference's depth-derived PLY
    (myutils/mesh_recon.py:184-258: vertex k = row*W + col; faces [k, k+W, k+1] and [k+1, k+W, k+W+1], which face the
    camera of default_cam.json).  Quads touching an invalid vertex are dropped, like the reference's zero-depth test.
    Returns (verts (H*W,3) float32, tris (nt,3) int32)."""
    pos = np.asarray(pos, dtype=np.float32)
    H, W, _ = pos.shape
    k = (np.arange(H - 1)[:, None] * W + np.arange(W - 1)[None, :]).reshape(-1)
    t0 = np.stack([k, k + W, k + 1], -1)
    t1 = np.stack([k + 1, k + W, k + W + 1], -1)
    tris = np.stack([t0, t1], 1).reshape(-1, 3)
    if valid is not None:
        ok = np.asarray(valid).reshape(-1).astype(bool)
        tris = tris[ok[tris].all(-1)]
    return np.ascontiguousarray(pos.reshape(-1, 3)), np.ascontiguousarray(tris.astype(np.int32))


def bumpy_positions(H, W, camera=None, amp=3.0, freq=3.0):
    """Positions on the pixel-centre rays of a height field with enough relief for self-occlusion and interreflection:
    depth = 20 + 4 cos(pi u) cos(pi v) - amp * max(0, sin(freq 2 pi u) sin(freq 2 pi v))."""
    cam = camera or Camera(width=W, height=H)
    t, aspect = cam.tan_half_fov_x, W / H
    sx, sy = np.meshgrid(np.arange(W, dtype=np.float64) + 0.5, np.arange(H, dtype=np.float64) + 0.5)
    u, v = sx / W, sy / H
    depth = 20.0 + 4.0 * np.cos(np.pi * u) * np.cos(np.pi * v) - amp * np.maximum(0.0, np.sin(freq * 2 * np.pi * u) * np.sin(freq * 2 * np.pi * v))
    l = np.stack([(1 - 2 * sx / W) * t, (1 - 2 * sy / H) * t / aspect, np.ones_like(sx)], -1) * depth[..., None]
    return (l @ cam.to_world[:3, :3].T + cam.to_world[:3, 3]).astype(np.float32)
