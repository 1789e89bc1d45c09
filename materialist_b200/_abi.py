"""ctypes binding of libmaterialist_b200.so — the C-ABI declared in include/materialist_b200.h.

The library is built in-tree by materialist_b200/csrc/Makefile (nvcc, sm_100a).  There is NO CPU
fallback: if the shared library is missing the import of this module raises, and every compute call
raises if the tensors are not CUDA tensors.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MB200_LIB") or os.path.join(_HERE, "libmaterialist_b200.so")   # MB200_LIB: tuning builds only

MAX_LEVELS = 24
FILM_TAPS = 25

OK, EINVAL, ERANGE, ELAUNCH, EUNSUPPORTED, EIO = 0, -1, -2, -3, -4, -5
IMG_SRGB = 1
FLAG_WO_WORLD_QUIRK, FLAG_ROW_STRIDE_H, FLAG_ENV_HALF_TEXEL, FLAG_AD_WEIGHTS = 1, 2, 4, 8
FILTER_BOX, FILTER_GAUSSIAN = 0, 1
ENV_ASSIGNED, ENV_FILE = 0, 1
POSMLP_TCGEN05, POSMLP_FFMA = 0, 1


class Cfg(C.Structure):
    _fields_ = [("H", C.c_int32), ("W", C.c_int32), ("spp", C.c_int32), ("max_depth", C.c_int32),
                ("seed", C.c_uint32), ("filter", C.c_int32), ("flags", C.c_int32), ("use_mesh_normal", C.c_int32),
                ("row0", C.c_int32), ("rows", C.c_int32),
                ("view", C.c_float * 16), ("proj", C.c_float * 16), ("cam_to_world", C.c_float * 16),
                ("tan_half_fov_x", C.c_float), ("env_u_shift", C.c_float)]


class HierDesc(C.Structure):
    _fields_ = [("res_x", C.c_int32), ("res_y", C.c_int32), ("n_levels", C.c_int32),
                ("lvl_off", C.c_int32 * MAX_LEVELS), ("lvl_w", C.c_int32 * MAX_LEVELS), ("lvl_h", C.c_int32 * MAX_LEVELS),
                ("total_floats", C.c_int32)]


MESH_MAX_LEVELS = 16


class MeshDesc(C.Structure):
    _fields_ = [("nv", C.c_int32), ("nt", C.c_int32), ("face_normals", C.c_int32),
                ("n_slots", C.c_int32), ("n_levels", C.c_int32), ("n_groups", C.c_int32),
                ("lvl_nodes", C.c_int32 * MESH_MAX_LEVELS), ("lvl_group_off", C.c_int32 * MESH_MAX_LEVELS),
                ("off_header", C.c_int64), ("off_tv", C.c_int64), ("off_tn", C.c_int64), ("off_nodes", C.c_int64),
                ("total_bytes", C.c_int64)]


class Trans(C.Structure):
    """mb200_trans: TransBSDF parameters (device pointers)."""
    _fields_ = [("ior", C.c_float), ("spec_trans", C.c_float), ("refract_distance", C.c_float), ("reserved", C.c_int32),
                ("bg", C.c_void_p), ("mask", C.c_void_p)]


class PosMLPDesc(C.Structure):
    _fields_ = [("n_color", C.c_int32), ("n_out", C.c_int32), ("hidden", C.c_int32), ("n_freq", C.c_int32),
                ("output_type", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("impl", C.c_int32), ("row0", C.c_int32)]


class AdamSeg(C.Structure):
    _fields_ = [("p", C.c_void_p), ("mat", C.c_void_p), ("g", C.c_void_p), ("ori", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p),
                ("n", C.c_int64), ("lo", C.c_float), ("hi", C.c_float), ("aux_coeff", C.c_float)]


MAX_PEERS, PEER_MAX_PUSH, PEER_HALO, PEER_MAP = 16, 8, 0, 1


class Peer(C.Structure):
    """mb200_peer: this rank, the world, the iteration's sequence number and every rank's mailbox as mapped into this process."""
    _fields_ = [("rank", C.c_int32), ("world", C.c_int32), ("seq", C.c_uint32), ("reserved", C.c_uint32), ("box", C.c_void_p * MAX_PEERS)]


class PushSeg(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("n_float4", C.c_int64)]


class MB200Error(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the CUDA extension is not built. Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C materialist_b200/csrc`. There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, sz = C.c_void_p, C.c_int, C.c_int64, C.c_size_t
    pc, ph, pm, pd, pt = C.POINTER(Cfg), C.POINTER(HierDesc), C.POINTER(PosMLPDesc), C.POINTER(MeshDesc), C.POINTER(Trans)
    sig = {
        "mb200_strerror": (C.c_char_p, [i32]),
        "mb200_last_cuda_error": (C.c_char_p, []),
        "mb200_version": (i32, []),
        "mb200_env_internal_width": (i32, [i32, i32]),
        "mb200_hier_describe": (i32, [i32, i32, ph]),
        "mb200_env_scratch_bytes": (sz, [i32, i32]),
        "mb200_fwd_partial_rows": (i32, [pc, C.POINTER(C.c_int)]),
        "mb200_partial_stride": (i32, [i32]),
        "mb200_env_prepare": (i32, [vp, i32, i32, i32, vp, vp, ph, vp, vp]),
        "mb200_env_grad_slabs": (i32, [i32, i32, i32]),
        "mb200_env_grad_finish": (i32, [vp, i32, i32, i32, i32, vp, vp]),
        "mb200_shade_fwd": (i32, [pc] + [vp] * 8 + [ph, vp, vp]),
        "mb200_film_develop": (i32, [pc, vp, vp, vp]),
        "mb200_film_weights": (i32, [pc, vp, vp]),
        "mb200_film_adjoint": (i32, [pc, vp, vp, vp, vp]),
        "mb200_bwd_wpart_rows": (i32, [pc, C.POINTER(C.c_int)]),
        "mb200_bwd_gadj_rows": (i32, [pc, C.POINTER(C.c_int)]),
        "mb200_shade_bwd": (i32, [pc] + [vp] * 8 + [ph] + [vp] * 6 + [i32, vp]),
        "mb200_debug_sample_indices": (i32, [pc, vp, vp, vp, ph, vp, vp]),
        "mb200_debug_sample_record": (i32, [pc, vp, vp, vp, vp, vp, vp, vp, vp, ph, vp, vp, vp]),
        "mb200_mesh_describe": (i32, [i32, i32, i32, pd]),
        "mb200_mesh_scratch_bytes": (sz, [i32, i32, i32]),
        "mb200_mesh_build": (i32, [vp, vp, pd, vp, vp, vp]),
        "mb200_mesh_shade_fwd": (i32, [pc, pd] + [vp] * 7 + [ph, vp, vp]),
        "mb200_mesh_shade_bwd": (i32, [pc, pd] + [vp] * 7 + [ph] + [vp] * 6 + [i32, vp]),
        "mb200_mesh_intersect": (i32, [pd, vp, vp, vp, vp, i32, i32, vp, vp, vp]),
        "mb200_mesh_primary": (i32, [pc, pd, vp, C.c_float, C.c_float, vp, vp, vp, vp, vp]),
        "mb200_bsdf_eval_pdf": (i32, [pc, i64] + [vp] * 11),
        "mb200_bsdf_sample": (i32, [pc, i64] + [vp] * 13),
        "mb200_bsdf_eval_grad": (i32, [pc, i64] + [vp] * 14),
        "mb200_trans_shade_fwd": (i32, [pc, pt] + [vp] * 8 + [ph, vp, vp]),
        "mb200_trans_mesh_shade_fwd": (i32, [pc, pt, pd] + [vp] * 7 + [ph, vp, vp]),
        "mb200_mesh_fwd_wf_scratch_bytes": (sz, [pc]),
        "mb200_mesh_shade_fwd_wf": (i32, [pc, pt, pd] + [vp] * 7 + [ph, vp, vp, sz, vp, vp]),
        "mb200_mesh_primary_index_bytes": (sz, [pc, pd]),
        "mb200_mesh_primary_index_build": (i32, [pc, pd, vp, vp, vp]),
        "mb200_mesh_bwd_wf_scratch_bytes": (sz, [pc]),
        "mb200_mesh_shade_bwd_wf": (i32, [pc, pd] + [vp] * 7 + [ph] + [vp] * 6 + [i32, vp, sz, vp, vp]),
        "mb200_trans_eval_pdf": (i32, [pc, pt, i64] + [vp] * 11),
        "mb200_trans_sample": (i32, [pc, pt, i64] + [vp] * 13),
        "mb200_trans_refracted_texel": (i32, [pc, pt, i64] + [vp] * 6),
        "mb200_debug_exact_math": (i32, [i32, vp, vp, i64, vp, vp, vp]),
        "mb200_image_info": (i32, [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
        "mb200_image_read": (i32, [C.c_char_p, vp, i32, i32, i32]),
        "mb200_image_write": (i32, [C.c_char_p, vp, i32, i32, i32]),
        "mb200_image_write_ex": (i32, [C.c_char_p, vp, i32, i32, i32, i32]),
        "mb200_posmlp_param_count": (i64, [pm]),
        "mb200_posmlp_cache_bytes": (sz, [pm, i64]),
        "mb200_posmlp_workspace_bytes": (sz, [pm]),
        "mb200_posmlp_fwd": (i32, [pm, vp, vp, i64, vp, vp, vp, vp]),
        "mb200_posmlp_bwd": (i32, [pm, vp, vp, i64, vp, vp, vp, vp, vp, vp]),
        "mb200_cdf_build": (i32, [vp, i32, i32, vp, vp, vp]),
        "mb200_cdf_sample": (i32, [vp, vp, i32, i32, vp, i64, vp, vp, vp, vp, vp]),
        "mb200_sh_project": (i32, [vp, i32, i32, vp, i64, vp, vp]),
        "mb200_sh_reconstruct": (i32, [vp, i32, i32, i32, vp, vp]),
        "mb200_reduce_scratch_bytes": (sz, []),
        "mb200_image_sum": (i32, [vp, i64, vp, vp, vp]),
        "mb200_loss_srgb_sums": (i32, [vp, vp, i64, vp, vp, vp, vp, vp]),
        "mb200_loss_srgb_grad": (i32, [vp, vp, i64, vp, vp, i64, vp, vp]),
        "mb200_adam_clamped": (i32, [C.POINTER(AdamSeg), i32, C.c_float, C.c_float, C.c_float, C.c_float, i32, vp]),
        "mb200_probe_ffma": (i32, [vp, i32, C.POINTER(C.c_double), vp]),
        "mb200_peer_box_bytes": (sz, []),
        "mb200_peer_alloc": (i32, [sz, C.POINTER(C.c_void_p), vp]),
        "mb200_peer_open": (i32, [vp, C.POINTER(C.c_void_p)]),
        "mb200_peer_close": (i32, [vp]),
        "mb200_peer_free": (i32, [vp]),
        "mb200_image_sum_peer": (i32, [vp, i64, vp, vp, C.POINTER(Peer), vp]),
        "mb200_loss_srgb_sums_peer": (i32, [vp, vp, i64, vp, vp, vp, vp, C.POINTER(Peer), vp]),
        "mb200_loss_srgb_grad_peer": (i32, [vp, vp, i64, vp, vp, i64, vp, C.POINTER(Peer), vp]),
        "mb200_peer_push": (i32, [C.POINTER(Peer), i32, C.POINTER(PushSeg), i32, i32, i32, vp, vp]),
        "mb200_peer_wait": (i32, [C.POINTER(Peer), i32, i32, i32, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)           # AttributeError here = header / library mismatch
        fn.restype = res
        fn.argtypes = args
    return lib, tuple(sig)


lib, EXPORTS = _load()


def check(rc, what=""):
    if rc != OK:
        msg = lib.mb200_strerror(rc).decode()
        if rc == ELAUNCH:
            msg += " — " + lib.mb200_last_cuda_error().decode()
        if rc == EINVAL:
            raise ValueError(f"{what}: {msg}")
        raise MB200Error(f"{what}: {msg} (rc={rc})")


def ptr(t, dtype=None):
    """Device pointer of a contiguous CUDA tensor (None -> NULL). Raises on CPU tensors: no fallback.
    dtype: when given, the tensor must have exactly this dtype (the kernels reinterpret nothing)."""
    if t is None:
        return None
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"expected a torch.Tensor, got {type(t)}")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"expected a {dtype} tensor, got {t.dtype}")
    if not t.is_cuda:
        raise ValueError("materialist_b200 runs on CUDA tensors only (no CPU fallback); got a CPU tensor")
    if not t.is_contiguous():
        raise ValueError("tensor must be contiguous")
    return C.c_void_p(t.data_ptr())


def fptr(t):
    """ptr() for the float32 buffers of the render / optimiser entry points (user tensors arrive here: a float64 or half map
    would otherwise be reinterpreted bit-wise by a float* kernel)."""
    return ptr(t, torch.float32)


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def hier_describe(res_x, res_y):
    d = HierDesc()
    check(lib.mb200_hier_describe(res_x, res_y, C.byref(d)), "mb200_hier_describe")
    return d
