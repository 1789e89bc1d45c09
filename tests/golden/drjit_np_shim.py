"""numpy float32 stand-ins for the subset of `drjit` / `mitsuba` that myutils/mi_plugin.py touches in
MatDiffBSDF / TransBSDF (mi_plugin.py:217-283, 645-671, 1229-1770).

Purpose: mitsuba==3.5.2 / drjit==0.4.6 are not installable here, so the reference's Dr.Jit-typed BSDF code could not be
run to produce golden vectors.  With these stand-ins the REFERENCE'S OWN SOURCE LINES execute unmodified (imported from
/root/reference by make_golden.py) on seeded lanes; only the array primitives underneath are emulated:

  * every array is float32 / int32 / bool, python scalars are rounded to float32 first (Dr.Jit's promotion rule);
  * `a[mask] = b` is a masked assignment over full-width arrays and `a[mask]` returns the full-width array (the lanes a
    Dr.Jit kernel computes and then discards) — never a compaction;
  * `x ** n` with an integer n is repeated multiplication (dr.power), with a float exponent it is powf;
  * `mi.Frame3f(n)` is Mitsuba's `coordinate_system` (Duff et al. 2017), `Matrix4f @ Vector4f` accumulates column by
    column with fused multiply-adds as `dr::Matrix::operator*` does.  These two are upstream (not reference) code and
    are restated here from the published sources.

Test infrastructure only: nothing in the product imports this file.
"""
import math
import os
import sys
import types

import numpy as np

# MB_SHIM_F64=1: every "float32" below becomes float64 — used by make_bsdf_grad_golden.py to take finite differences of the
# reference's BSDF source in double precision (the adjoint pin); the default is Dr.Jit's float32.
F32 = np.float64 if os.environ.get("MB_SHIM_F64") == "1" else np.float32


def _raw(x):
    if isinstance(x, A):
        return x.v
    if isinstance(x, (bool, np.bool_)):
        return np.bool_(x)
    if isinstance(x, (int, np.integer)):
        return x
    if isinstance(x, (float, np.floating)):
        return F32(x)
    if hasattr(x, "detach"):                       # torch tensor
        return x.detach().cpu().numpy()
    return np.asarray(x)


def _wrap(v):
    v = np.asarray(v)
    if v.dtype == np.bool_:
        return Bool(v)
    if np.issubdtype(v.dtype, np.integer):
        return Int(v)
    return Float(v)


def _fma(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(F32)   # exact product, one rounding


class A:
    """1-D (or broadcast scalar) array."""
    dtype = F32

    def __init__(self, v=0, *rest):
        if rest:
            raise TypeError("scalar array type takes one argument")
        self.v = np.asarray(_raw(v)).astype(self.dtype)

    def __len__(self):
        return int(self.v.shape[0]) if self.v.ndim else 1

    def _b(self, o, f):
        if isinstance(o, (Vec,)):
            return NotImplemented
        a, b = self.v, _raw(o)
        if a.dtype == F32 and not isinstance(b, np.ndarray):
            b = F32(b)
        elif isinstance(b, np.ndarray) and b.dtype == np.float64:
            b = b.astype(F32)
        with np.errstate(all="ignore"):
            return _wrap(f(a, b))

    def __add__(self, o): return self._b(o, lambda a, b: a + b)
    def __radd__(self, o): return self._b(o, lambda a, b: b + a)
    def __sub__(self, o): return self._b(o, lambda a, b: a - b)
    def __rsub__(self, o): return self._b(o, lambda a, b: b - a)
    def __mul__(self, o): return self._b(o, lambda a, b: a * b)
    def __rmul__(self, o): return self._b(o, lambda a, b: b * a)
    def __truediv__(self, o): return self._b(o, lambda a, b: a / b)
    def __rtruediv__(self, o): return self._b(o, lambda a, b: b / a)
    def __neg__(self): return _wrap(-self.v)
    def __pos__(self): return self
    def __gt__(self, o): return self._b(o, lambda a, b: a > b)
    def __ge__(self, o): return self._b(o, lambda a, b: a >= b)
    def __lt__(self, o): return self._b(o, lambda a, b: a < b)
    def __le__(self, o): return self._b(o, lambda a, b: a <= b)
    def __and__(self, o): return _wrap(self.v & _raw(o))
    def __or__(self, o): return _wrap(self.v | _raw(o))
    def __invert__(self): return _wrap(~self.v)
    def __mod__(self, o): return self._b(o, lambda a, b: a % b)
    def __floordiv__(self, o): return self._b(o, lambda a, b: a // b)

    def __pow__(self, e):
        if isinstance(e, (int, np.integer)):
            r = None; base = self.v; n = int(e)                 # dr.power(x, int): square-and-multiply
            while n:
                if n & 1:
                    r = base if r is None else (r * base).astype(F32)
                base = (base * base).astype(F32); n >>= 1
            return _wrap(r if r is not None else np.ones_like(self.v))
        with np.errstate(all="ignore"):
            return _wrap(np.power(self.v, F32(e)).astype(F32))

    def __getitem__(self, k):
        if isinstance(k, A):
            return self                                         # masked read: the full-width lanes
        return _wrap(self.v[k])

    def __setitem__(self, k, val):
        if not isinstance(k, A):
            raise TypeError("only masked assignment is emulated")
        self.v = np.where(k.v, np.broadcast_to(_raw(val), np.broadcast(k.v, self.v).shape), self.v).astype(self.dtype)

    def numpy(self):
        return self.v


class Float(A): dtype = F32
class Int(A): dtype = np.int32
class UInt32(A): dtype = np.uint32
class Bool(A): dtype = np.bool_


class Vec:
    """fixed-size vector of Float components (mi.Vector3f / Normal3f / Vector2f / Vector4f)."""
    n = 3

    def __init__(self, *a):
        if len(a) == 0:
            a = (0.0,)
        if len(a) == 1:
            x = a[0]
            if isinstance(x, Vec):
                self.c = [Float(c.v.copy()) for c in x.c]
            elif isinstance(x, np.ndarray) and x.ndim == 2:
                self.c = [Float(x[:, i]) for i in range(self.n)]
            else:
                self.c = [Float(x) for _ in range(self.n)]
        else:
            assert len(a) == self.n
            self.c = [c if isinstance(c, Float) else Float(c) for c in a]

    def __len__(self):
        return self.n

    def _b(self, o, f):
        oc = o.c if isinstance(o, Vec) else [o] * self.n
        return type(self)(*[f(a, b) for a, b in zip(self.c, oc)])

    def __add__(self, o): return self._b(o, lambda a, b: a + b)
    def __radd__(self, o): return self._b(o, lambda a, b: b + a)
    def __sub__(self, o): return self._b(o, lambda a, b: a - b)
    def __rsub__(self, o): return self._b(o, lambda a, b: b - a)
    def __mul__(self, o): return self._b(o, lambda a, b: a * b)
    def __rmul__(self, o): return self._b(o, lambda a, b: b * a)
    def __truediv__(self, o): return self._b(o, lambda a, b: a / b)
    def __neg__(self): return type(self)(*[-a for a in self.c])
    def __pow__(self, e): return type(self)(*[a ** e for a in self.c])
    def __gt__(self, o): return MaskVec(*[a > b for a, b in zip(self.c, o.c if isinstance(o, Vec) else [o] * self.n)])
    def __lt__(self, o): return MaskVec(*[a < b for a, b in zip(self.c, o.c if isinstance(o, Vec) else [o] * self.n)])

    def __getitem__(self, k):
        if isinstance(k, A):
            return self
        return self.c[k]

    def __setitem__(self, k, val):
        if isinstance(k, A):
            vc = val.c if isinstance(val, Vec) else [val] * self.n
            for a, b in zip(self.c, vc):
                a[k] = b
        else:
            self.c[k] = val if isinstance(val, Float) else Float(val)

    x = property(lambda s: s.c[0]); y = property(lambda s: s.c[1]); z = property(lambda s: s.c[2])

    def numpy(self):
        L = max(len(c) if c.v.ndim else 1 for c in self.c)
        return np.stack([np.broadcast_to(c.v, (L,)) for c in self.c], -1).astype(F32)


class MaskVec:
    def __init__(self, *c): self.c = list(c)


class Vector3f(Vec): n = 3
class Normal3f(Vec): n = 3
class Vector2f(Vec): n = 2
class Vector4f(Vec): n = 4


class Matrix4f:
    def __init__(self, m):
        self.m = np.asarray(_raw(m), F32).reshape(4, 4)

    def __matmul__(self, v):
        rows = []
        for i in range(4):                                     # sum = col0*v0; sum = fmadd(col_j, v_j, sum)
            acc = (self.m[i, 0] * v.c[0].v).astype(F32)
            for j in range(1, 4):
                acc = _fma(np.broadcast_to(self.m[i, j], np.shape(v.c[j].v)), v.c[j].v, np.broadcast_to(acc, np.broadcast(acc, v.c[j].v).shape))
            rows.append(Float(acc))
        return Vector4f(*rows)


class Frame3f:
    """mitsuba Frame3f(n): coordinate_system(n) of Duff et al. (include/mitsuba/core/vector.h)."""

    def __init__(self, n):
        nx, ny, nz = (c.v for c in n.c)
        sign = np.copysign(F32(1), nz).astype(F32)
        a = (F32(-1) / (sign + nz)).astype(F32)
        b = (nx * ny * a).astype(F32)
        mulsign = lambda x: (x * sign).astype(F32)
        self.s = Vector3f(Float(mulsign((nx * nx).astype(F32) * a) + F32(1)), Float(mulsign(b)), Float(-mulsign(nx)))
        self.t = Vector3f(Float(b), Float(_fma(ny, (ny * a).astype(F32), sign)), Float(-ny))
        self.n = n

    def to_world(self, v):
        # dr::fmadd(n, v.z, fmadd(t, v.y, s * v.x))
        out = []
        for s, t, n in zip(self.s.c, self.t.c, self.n.c):
            L = np.broadcast(s.v, v.c[0].v).shape
            r = (np.broadcast_to(s.v, L) * v.c[0].v).astype(F32)
            r = _fma(np.broadcast_to(t.v, L), np.broadcast_to(v.c[1].v, L), r)
            r = _fma(np.broadcast_to(n.v, L), np.broadcast_to(v.c[2].v, L), r)
            out.append(Float(r))
        return Vector3f(*out)


class TensorXf:
    def __init__(self, data=0.0, shape=None):
        if isinstance(data, TensorXf):
            self.a = data.a.copy()
        elif shape is not None:
            self.a = np.full(shape, data, F32)
        else:
            d = _raw(data)
            self.a = np.ascontiguousarray(d) if d.dtype == np.bool_ else np.ascontiguousarray(d, F32)

    shape = property(lambda s: s.a.shape)

    @property
    def array(self):
        return _wrap(self.a.reshape(-1))

    def __ge__(self, o): t = TensorXf.__new__(TensorXf); t.a = self.a >= o; return t
    def __gt__(self, o): t = TensorXf.__new__(TensorXf); t.a = self.a > o; return t


def _unary(f):
    def g(x):
        if isinstance(x, Vec):
            return type(x)(*[g(c) for c in x.c])
        with np.errstate(all="ignore"):
            return _wrap(f(np.asarray(_raw(x), F32)).astype(F32))
    return g


def _select(m, a, b):
    if isinstance(m, MaskVec):
        T = type(a) if isinstance(a, Vec) else type(b)
        ac = a.c if isinstance(a, Vec) else [a] * T.n
        bc = b.c if isinstance(b, Vec) else [b] * T.n
        return T(*[_select(mm, x, y) for mm, x, y in zip(m.c, ac, bc)])
    if isinstance(a, Vec) or isinstance(b, Vec):
        T = type(a) if isinstance(a, Vec) else type(b)
        ac = a.c if isinstance(a, Vec) else [a] * T.n
        bc = b.c if isinstance(b, Vec) else [b] * T.n
        return T(*[_select(m, x, y) for x, y in zip(ac, bc)])
    ra, rb = _raw(a), _raw(b)
    if isinstance(ra, float) or (isinstance(ra, np.ndarray) and ra.dtype == np.float64):
        ra = F32(ra)
    if isinstance(rb, float) or (isinstance(rb, np.ndarray) and rb.dtype == np.float64):
        rb = F32(rb)
    return _wrap(np.where(_raw(m), ra, rb))


def _dot(a, b):
    # dr::dot: fmadd chain  x*x' -> fmadd(y,y',.) -> fmadd(z,z',.)
    L = np.broadcast(*[c.v for c in a.c], *[c.v for c in b.c]).shape
    acc = (np.broadcast_to(a.c[0].v, L) * np.broadcast_to(b.c[0].v, L)).astype(F32)
    for i in range(1, a.n):
        acc = _fma(np.broadcast_to(a.c[i].v, L), np.broadcast_to(b.c[i].v, L), acc)
    return Float(acc)


def _normalize(v):
    with np.errstate(all="ignore"):
        inv = Float((F32(1) / np.sqrt(_dot(v, v).v)).astype(F32))     # dr::normalize = v * rsqrt(squared_norm)
    return v * inv


def _gather(T, arr, idx, active=True):
    i = idx.v.astype(np.int64)
    flat = arr.v
    if issubclass(T, Vec):
        return T(*[_wrap(flat[T.n * i + c]) for c in range(T.n)])
    return T(flat[i])


def _isnan(x):
    if isinstance(x, Vec):
        return MaskVec(*[Bool(np.isnan(c.v)) for c in x.c])
    return Bool(np.isnan(x.v))


def _clamp(x, lo, hi):
    if isinstance(x, Vec):
        return type(x)(*[_clamp(c, lo, hi) for c in x.c])
    return _wrap(np.minimum(np.maximum(x.v, F32(lo)), F32(hi)).astype(F32))


def _binary(f):
    def g(a, b):
        ra, rb = _raw(a), _raw(b)
        ra = F32(ra) if not isinstance(ra, np.ndarray) else ra.astype(F32)
        rb = F32(rb) if not isinstance(rb, np.ndarray) else rb.astype(F32)
        return Float(f(ra, rb).astype(F32))
    return g


class _Flags(int):
    def __or__(self, o): return _Flags(int(self) | int(o))
    def __pos__(self): return int(self)


class _Props(dict):
    def has_property(self, k): return k in self


class BSDFSample3f:
    pass


def install():
    """Registers the stand-ins as `drjit` and `mitsuba` in sys.modules (call BEFORE importing myutils.mi_plugin)."""
    dr = types.ModuleType("drjit")
    dr.dot = _dot; dr.normalize = _normalize; dr.select = _select; dr.gather = _gather; dr.isnan = _isnan; dr.clamp = _clamp
    dr.maximum = _binary(np.maximum); dr.minimum = _binary(np.minimum)
    dr.floor = _unary(np.floor); dr.sin = _unary(np.sin); dr.cos = _unary(np.cos); dr.asin = _unary(np.arcsin)
    dr.sqrt = _unary(np.sqrt); dr.safe_sqrt = _unary(lambda x: np.sqrt(np.maximum(x, F32(0))))
    dr.wrap_ad = lambda **k: (lambda f: f)
    dr.set_flag = lambda *a, **k: None
    dr.JitFlag = types.SimpleNamespace(VCallRecord=0, LoopRecord=1)
    mi = types.ModuleType("mitsuba")
    mi.set_variant = lambda *a, **k: None
    mi.register_bsdf = lambda *a, **k: None
    mi.BSDF = type("BSDF", (), {"__init__": lambda self, props=None: None})
    mi.BSDFFlags = types.SimpleNamespace(SpatiallyVarying=_Flags(1 << 16), DiffuseReflection=_Flags(1 << 2), FrontSide=_Flags(1 << 17),
                                         BackSide=_Flags(1 << 18), GlossyReflection=_Flags(1 << 4))
    mi.ParamFlags = types.SimpleNamespace(Differentiable=0, NonDifferentiable=1)
    mi.Float = Float; mi.Int = Int; mi.UInt32 = UInt32; mi.Bool = Bool
    mi.Vector3f = Vector3f; mi.Normal3f = Normal3f; mi.Vector2f = Vector2f; mi.Vector4f = Vector4f; mi.Point3f = Vector3f
    mi.Matrix4f = Matrix4f; mi.Frame3f = Frame3f; mi.TensorXf = TensorXf; mi.BSDFSample3f = BSDFSample3f
    mi.Properties = _Props
    sys.modules["drjit"] = dr; sys.modules["mitsuba"] = mi
    return dr, mi


class FakeSI:
    """SurfaceInteraction3f stand-in: world-space p, geometric normal n, shading frame sh (Frame3f), local wi."""

    def __init__(self, p, n, wi_world):
        self.p = Vector3f(np.asarray(p, F32)); self.n = Normal3f(np.asarray(n, F32))
        self.sh_frame = Frame3f(self.n)
        w = Vector3f(np.asarray(wi_world, F32))
        self.wi = Vector3f(_dot(w, self.sh_frame.s), _dot(w, self.sh_frame.t), _dot(w, self.sh_frame.n))

    def to_world(self, v):
        return self.sh_frame.to_world(v)

    def to_local(self, w):
        return Vector3f(_dot(w, self.sh_frame.s), _dot(w, self.sh_frame.t), _dot(w, self.sh_frame.n))
