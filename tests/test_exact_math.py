"""include/mb200_exact_math.h as gcc compiles it (oracle/mb_oracle_aux.c front-ends) against float64 libm: the shared reproducible
sincospi / atan2 / acos / asin / rsqrt that put the kernels and the oracle on one bit-exact direction chain.  Bars: <= 2 ulp for
the transcendental functions, correctly rounded (0.5 ulp) for rsqrt."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle as orc


@pytest.fixture(scope="module")
def lib():
    return orc.Oracle().lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _ulps(got, want64):
    want32 = want64.astype(np.float32)
    ulp = np.spacing(np.abs(want32)).astype(np.float64)
    ulp = np.maximum(ulp, np.float64(np.finfo(np.float32).tiny))
    return np.abs(got.astype(np.float64) - want64) / ulp


def test_sincospi(lib):
    rs = np.random.RandomState(0)
    x = np.concatenate([rs.uniform(0, 2, 200000), rs.uniform(-4, 4, 50000), np.arange(-8, 9) * 0.25]).astype(np.float32)
    s = np.empty_like(x); c = np.empty_like(x)
    lib.mbo_exact_sincospi(_p(x), C.c_size_t(x.size), _p(s), _p(c))
    ws, wc = np.sin(np.pi * x.astype(np.float64)), np.cos(np.pi * x.astype(np.float64))
    # near the zeros of sin(pi x) / cos(pi x) the float64 reference itself carries pi's rounding: bound the absolute error there
    assert np.all((_ulps(s, ws) <= 2) | (np.abs(s - ws) < 3e-8)) and np.all((_ulps(c, wc) <= 2) | (np.abs(c - wc) < 3e-8))
    k = np.arange(-8, 9, dtype=np.float32)
    s = np.empty_like(k); c = np.empty_like(k)
    lib.mbo_exact_sincospi(_p(k), C.c_size_t(k.size), _p(s), _p(c))
    assert np.all(s == 0) and np.all(np.abs(c) == 1)          # exact at the integers


def test_atan2_acos_asin(lib):
    rs = np.random.RandomState(1)
    y = rs.randn(300000).astype(np.float32); x = rs.randn(300000).astype(np.float32)
    o = np.empty_like(x)
    lib.mbo_exact_atan2(_p(y), _p(x), C.c_size_t(x.size), _p(o))
    assert _ulps(o, np.arctan2(y.astype(np.float64), x.astype(np.float64))).max() <= 2
    x = np.concatenate([rs.uniform(-1, 1, 300000), [-1.0, -0.5, 0.0, 0.5, 1.0]]).astype(np.float32)
    o = np.empty_like(x)
    lib.mbo_exact_acos(_p(x), C.c_size_t(x.size), _p(o))
    assert _ulps(o, np.arccos(x.astype(np.float64))).max() <= 2
    x = np.concatenate([rs.uniform(0, 1, 300000), [0.0, 0.5, 1.0]]).astype(np.float32)
    o = np.empty_like(x)
    lib.mbo_exact_asin01(_p(x), C.c_size_t(x.size), _p(o))
    assert _ulps(o, np.arcsin(x.astype(np.float64))).max() <= 2


def test_rsqrt_correctly_rounded(lib):
    rs = np.random.RandomState(2)
    x = np.concatenate([rs.uniform(0.25, 4, 400000), np.exp(rs.uniform(-60, 60, 100000))]).astype(np.float32)
    o = np.empty_like(x)
    lib.mbo_exact_rsqrt(_p(x), C.c_size_t(x.size), _p(o))
    want = 1.0 / np.sqrt(x.astype(np.longdouble))
    err = np.abs(o.astype(np.longdouble) - want) / np.spacing(o).astype(np.longdouble)
    assert err.max() <= 0.5 + 1e-9
