"""materialist_b200 — B200-native differentiable envmap shading (forward + adjoint) behind the
plugin / operator surface of lez-s/Materialist.  Importing this package loads the in-tree CUDA
library (libmaterialist_b200.so); there is no CPU fallback."""
from . import _abi
from .scene import Camera, Scene, SceneParameters, TransSettings, traverse
from .renderop import render, render_envmap, render_w_brdf, default_seed_grad, sample_indices, tea32
from . import synthetic
from .mesh import Mesh, read_ply_mesh

__all__ = ["Camera", "Scene", "SceneParameters", "TransSettings", "traverse", "render", "render_envmap", "render_w_brdf",
           "default_seed_grad", "sample_indices", "tea32", "synthetic", "Mesh", "read_ply_mesh"]
