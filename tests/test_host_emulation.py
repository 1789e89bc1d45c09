"""The kernels' own per-sample device source (materialist_b200/csrc/mb200_device.cuh, mb200_shade.cuh) compiled for the HOST
(tests/host_emul/: g++ -ffp-contract=off + stand-ins for the CUDA intrinsics) against the oracle, lane by lane, without a GPU:

* every integer decision and the BITS of both sampled directions are equal (hierarchy cell, texel, lobe, envmap cell of the
  emitter sample AND of the BSDF-sampled direction, emitter direction, BSDF-sampled direction) — the device source mirrors the
  oracle operation for operation on the whole direction chain (north_star: "CDF/sample indices bit-exact");
* the per-lane radiance agrees to float rounding.

The -m gpu twin (test_gpu_render_parity.py::test_sample_record_bit_exact) repeats this on the real nvcc build."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from helpers import Case
from oracle import oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("emul") / "libmb_emul.so")
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    if not os.path.exists(os.path.join(cuda_inc, "cuda_runtime.h")):
        pytest.skip("CUDA headers not found")
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off", "-Wno-attributes", "-w",
                    "-I" + cuda_inc, "-I" + os.path.join(ROOT, "materialist_b200", "csrc"), "-o", out,
                    os.path.join(HERE, "host_emul", "emul.cpp")], check=True)
    return C.CDLL(out)


def _p(x, t=C.c_float):
    return None if x is None else x.ctypes.data_as(C.POINTER(t))


def _both(emul, O, c, seed, ad=False):
    env_int, hier, d = O.env_prepare(c.env, c.env_mode)
    cfg = c.cfg(d, seed, extra_flags=orc.FLAG_AD_WEIGHTS if ad else 0)
    n_opt = None if c.use_mesh_normal else c.n
    ref, ref_L = O.sample_record(cfg, c.gpos, c.gnrm, c.a, c.r, c.m, n_opt, env_int, hier, d, want_radiance=True)
    env4 = np.ascontiguousarray(np.concatenate([env_int, np.zeros(env_int.shape[:2] + (1,), np.float32)], -1))
    S = cfg.rows * cfg.W * cfg.spp
    got = np.zeros((S, 12), np.int32); got_L = np.zeros((S, 3), np.float32)
    rc = emul.emul_sample_record(C.byref(cfg), _p(c.gpos), _p(c.gnrm), _p(c.a), _p(c.r), _p(c.m), _p(n_opt), _p(env4), _p(hier), C.byref(d),
                                 _p(got, C.c_int32), _p(got_L))
    assert rc == 0
    return ref, ref_L, got, got_L


@pytest.mark.parametrize("kw", [
    dict(H=32, W=32, spp=16, He=16, We=32),
    dict(H=24, W=40, spp=8, He=33, We=70, env_mode=orc.ENV_FILE, invalid_border=3),
    dict(H=32, W=32, spp=8, He=128, We=256, use_mesh_normal=False),
    dict(H=16, W=16, spp=32, He=64, We=128, sun=50000.0, env_mode=orc.ENV_FILE),
])
@pytest.mark.parametrize("ad", [False, True])
def test_device_source_decisions_equal_oracle(emul, oracle32, kw, ad):
    c = Case(**kw)
    ref, ref_L, got, got_L = _both(emul, oracle32, c, seed=11, ad=ad)
    names = ("hier off.x", "hier off.y", "texel", "lobe", "emitter cell", "bsdf-direction cell",
             "d_em.x", "d_em.y", "d_em.z", "d_bs.x", "d_bs.y", "d_bs.z")
    for k, nm in enumerate(names):
        bad = np.flatnonzero(ref[:, k] != got[:, k])
        assert bad.size == 0, (nm, bad.size, bad[:5], ref[bad[:5], k], got[bad[:5], k])
    # radiance per lane: everything downstream of the exact chain is smooth (host build: libm instead of the fast device functions)
    err = np.abs(got_L - ref_L).max(-1) / np.maximum(np.abs(ref_L).max(-1), 1e-3)
    assert np.percentile(err, 99.9) < 2e-5 and err.max() < 1e-3, (np.percentile(err, 99.9), err.max())
