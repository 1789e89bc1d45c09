"""The kernels' build of include/mb200_exact_math.h (nvcc) against the oracle's (gcc), bit for bit, and the two instruction-level
shortcuts of the kernels against the IEEE intrinsics they stand in for:
  * mbx_rsqrt: lean Newton step on the unscaled argument == __frsqrt_rn — EXHAUSTIVE over the 2^24 arguments of two adjacent
    binades (every mantissa x both exponent parities), plus other exponents sampled;
  * xdiv_pos / xsqrt_pos (branch-free fast paths of the hierarchy descent, with their deferred fallback) == __fdiv_rn / __fsqrt_rn."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def _dev(op, x, y=None, two=False):
    from materialist_b200 import _abi
    tx = torch.from_numpy(np.ascontiguousarray(x, np.float32)).cuda()
    ty = None if y is None else torch.from_numpy(np.ascontiguousarray(y, np.float32)).cuda()
    o0 = torch.empty_like(tx); o1 = torch.empty_like(tx) if two else None
    _abi.check(_abi.lib.mb200_debug_exact_math(op, _abi.ptr(tx), _abi.ptr(ty), tx.numel(), _abi.ptr(o0), _abi.ptr(o1), _abi.stream_ptr()), "exact_math")
    return (o0.cpu().numpy(), o1.cpu().numpy()) if two else o0.cpu().numpy()


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_device_and_oracle_builds_are_bit_identical():
    lib = orc.Oracle().lib
    rs = np.random.RandomState(0)
    n = 1 << 20
    x = np.concatenate([rs.uniform(0, 2, n), rs.uniform(-3, 3, n // 4)]).astype(np.float32)
    hs = np.empty_like(x); hc = np.empty_like(x)
    lib.mbo_exact_sincospi(_p(x), C.c_size_t(x.size), _p(hs), _p(hc))
    ds, dc = _dev(0, x, two=True)
    assert np.array_equal(_bits(hs), _bits(ds)) and np.array_equal(_bits(hc), _bits(dc))
    yy = rs.randn(n).astype(np.float32); xx = rs.randn(n).astype(np.float32)
    yy[:8] = [0, 0, 1, -1, 0, -0.0, 1e-30, -1e-30]; xx[:8] = [1, -1, 0, 0, 0, -1, 1e30, -1e30]
    h = np.empty_like(xx); lib.mbo_exact_atan2(_p(yy), _p(xx), C.c_size_t(xx.size), _p(h))
    assert np.array_equal(_bits(h), _bits(_dev(1, yy, xx)))
    x = np.concatenate([rs.uniform(-1, 1, n), [-1, -0.5, 0, 0.5, 1]]).astype(np.float32)
    h = np.empty_like(x); lib.mbo_exact_acos(_p(x), C.c_size_t(x.size), _p(h))
    assert np.array_equal(_bits(h), _bits(_dev(2, x)))
    x = np.abs(x)
    h = np.empty_like(x); lib.mbo_exact_asin01(_p(x), C.c_size_t(x.size), _p(h))
    assert np.array_equal(_bits(h), _bits(_dev(3, x)))
    x = np.concatenate([rs.uniform(0.01, 16, n), np.exp(rs.uniform(-80, 80, n // 4))]).astype(np.float32)
    h = np.empty_like(x); lib.mbo_exact_rsqrt(_p(x), C.c_size_t(x.size), _p(h))
    assert np.array_equal(_bits(h), _bits(_dev(4, x)))          # the oracle's exact-arithmetic definition == the device's result


def test_lean_rsqrt_equals_intrinsic_exhaustively():
    for e in (126, 60, 180):                                    # biased exponents e, e+1: [0.5, 2), and far binades inside [2^-60, 2^60]
        bits = (np.arange(1 << 24, dtype=np.uint32) + np.uint32(e << 23))
        lean, intr = _dev(4, bits.view(np.float32), two=True)
        assert np.array_equal(_bits(lean), _bits(intr)), e
    edge = np.array([0.0, -0.0, 1e-45, 1e-38, 2.0 ** -61, 2.0 ** -60, 2.0 ** 60, 2.0 ** 61, 3e38, np.inf, -1.0, np.nan], np.float32)
    lean, intr = _dev(4, edge, two=True)
    assert np.array_equal(_bits(lean), _bits(intr))


def test_branch_free_div_sqrt_equal_ieee():
    rs = np.random.RandomState(3)
    n = 1 << 22
    b = np.exp(rs.uniform(-20, 3, n)).astype(np.float32)
    a = (b * rs.uniform(0, 1.5, n)).astype(np.float32)
    a[:6] = [0, 1e-30, 1e-40, 1, 1, 0]; b[:6] = [1, 1, 1, 0, 1e-30, 0]         # zero / tiny / zero-denominator: the fallback
    fast, ieee = _dev(5, a, b, two=True)
    assert np.array_equal(_bits(fast), _bits(ieee))
    x = np.concatenate([np.exp(rs.uniform(-30, 10, n)), [0.0, 1e-30, 1e-42, 4.0]]).astype(np.float32)
    fast, ieee = _dev(6, x, two=True)
    assert np.array_equal(_bits(fast), _bits(ieee))
    bits = (np.arange(1 << 24, dtype=np.uint32) + np.uint32(126 << 23))        # sqrt: every mantissa of [0.5, 2)
    fast, ieee = _dev(6, bits.view(np.float32), two=True)
    assert np.array_equal(_bits(fast), _bits(ieee))
