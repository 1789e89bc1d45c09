#!/usr/bin/env python
"""Makes tests/golden/indoor_ds4.npz from the reference's SHIPPED finished scene output_imgs/indoor (read-only under
/root/reference): the 512x512 G-buffer positions (vertex k <-> pixel k of indoor.ply), the optimised a/r/m maps and
envmap, and the reference's own final render, all box-downsampled 4x to 128x128 so the fixture stays small.
Used by the statistical relighting check (BASELINE config C1)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from materialist_b200.gbuffer import read_ply_vertices, read_image, load_estimated_brdf  # noqa: E402

REF = "/root/reference/output_imgs/indoor"


def centre2x2(x):      # value at the centre of each 4x4 block = mean of its central 2x2 pixels
    H, W = x.shape[:2]
    b = x.reshape(H // 4, 4, W // 4, 4, -1)
    return b[:, 1:3, :, 1:3].mean((1, 3))


def box4(x):
    H, W = x.shape[:2]
    return x.reshape(H // 4, 4, W // 4, 4, -1).mean((1, 3))


def main():
    v = read_ply_vertices(os.path.join(REF, "indoor.ply"))[:512 * 512].reshape(512, 512, 3)
    mat = load_estimated_brdf(os.path.join(REF, "best_results"))
    env = read_image(os.path.join(REF, "best_results", "envmap.hdr"))[..., :3]
    rendered = read_image(os.path.join(REF, "best_results", "rendered_img.exr"))[..., :3]      # saved as sRGB (x^(1/2.2))
    gt = read_image(os.path.join(REF, "gt_image.exr"))[..., :3]
    np.savez_compressed(os.path.join(HERE, "indoor_ds4.npz"),
                        pos=centre2x2(v).astype(np.float32), albedo=centre2x2(mat["albedo"]).astype(np.float32),
                        roughness=centre2x2(mat["roughness"]).astype(np.float32), metallic=centre2x2(mat["metallic"]).astype(np.float32),
                        envmap=env.astype(np.float32), rendered_linear=box4(np.clip(rendered, 0, None) ** 2.2).astype(np.float32),
                        gt_linear=box4(gt).astype(np.float32))
    print("indoor_ds4.npz", env.shape, float(rendered.mean()), float(gt.mean()))


if __name__ == "__main__":
    main()
