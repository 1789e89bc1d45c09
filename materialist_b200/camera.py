"""The perspective sensor and the matrices MatDiffBSDF derives from the camera JSON (myutils/mi_plugin.py:1259-1275,
:585-595).  Pure host maths (numpy / torch CPU): importable without the CUDA library."""
import json
import math

import numpy as np
import torch

# The camera of the reference's myutils/default_cam.json (the `cam_meta` every script passes): perspective, x_fov 35 deg,
# clip [0.01f, 1e4], 512 x 512 film, sensor at the origin looking down -z (to_world = diag(-1, 1, -1, 1)).  These are the
# defaults of Camera() below; a caller with another camera passes its own JSON path.


class Camera:
    """Perspective sensor + the matrices MatDiffBSDF.__init__ derives from the camera JSON
    (mi_plugin.py:1259-1275: view = inverse(to_world), persp_proj_matx(fov, W/H, near, far))."""

    def __init__(self, to_world=None, x_fov=35.0, near=0.009999999776482582, far=10000.0, width=512, height=512,
                 ref_exact_proj=None):
        if to_world is None:
            to_world = np.diag([-1.0, 1.0, -1.0, 1.0])
        self.to_world = np.asarray(to_world, dtype=np.float64).reshape(4, 4)
        self.x_fov, self.near, self.far = float(x_fov), float(near), float(far)
        self.width, self.height = int(width), int(height)
        # mi_plugin.py:585-595 uses f/aspect for x and f for y with f = 1/tan(x_fov/2): exact for square films only.
        # For W != H the intended (sensor-consistent) matrix is x: f, y: f*aspect.
        if ref_exact_proj is None:
            ref_exact_proj = self.width == self.height
        self.ref_exact_proj = bool(ref_exact_proj)

    @classmethod
    def from_json(cls, path=None, width=None, height=None):
        if path is None:                            # the reference's default camera
            return cls(width=width or 512, height=height or 512)
        meta = json.load(open(path))
        w, h = meta["film.size"]
        return cls(np.array(meta["to_world"])[0], meta["x_fov"][0], meta["near_clip"], meta["far_clip"],
                   width or w, height or h)

    @property
    def view_matrix(self):
        return torch.inverse(torch.tensor(self.to_world, dtype=torch.float32)).numpy().astype(np.float32)

    @property
    def proj_matrix(self):
        fov = torch.deg2rad(torch.tensor(self.x_fov))
        f = float(1.0 / torch.tan(fov / 2.0))
        aspect = self.width / self.height
        near, far = self.near, self.far
        fx, fy = (f / aspect, f) if self.ref_exact_proj else (f, f * aspect)
        return np.array([[fx, 0, 0, 0], [0, fy, 0, 0],
                         [0, 0, (far + near) / (near - far), (2 * far * near) / (near - far)],
                         [0, 0, -1, 0]], dtype=np.float32)

    @property
    def tan_half_fov_x(self):
        return math.tan(math.radians(self.x_fov) / 2.0)

    def pixel_ray_dirs(self, sx, sy):
        """World-space directions of the sensor rays through film positions (sx, sy) in pixel units (numpy)."""
        t, aspect = self.tan_half_fov_x, self.width / self.height
        l = np.stack([(1 - 2 * sx / self.width) * t, (1 - 2 * sy / self.height) * t / aspect, np.ones_like(sx)], -1)
        l = l / np.linalg.norm(l, axis=-1, keepdims=True)
        return l @ self.to_world[:3, :3].T
