"""CPU: the C-ABI library loads and exports every symbol include/materialist_b200.h declares, with matching
argument counts in the ctypes binding; host-only helpers behave; compute entry points reject bad arguments
(MB200_EINVAL) before touching the GPU."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "materialist_b200.h")


def prototypes():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    out = {}
    for m in re.finditer(r"\b(?:int|int64_t|size_t|const char\*)\s+(mb200_\w+)\s*\(([^;{]*?)\)\s*;", src, re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else len(args.split(","))
    return out


def test_library_exports_every_declared_symbol():
    from materialist_b200 import _abi
    protos = prototypes()
    assert len(protos) >= 29
    nm = subprocess.run(["nm", "-D", "--defined-only", _abi.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (mb200_\w+)", nm))
    assert set(protos) <= exported, set(protos) - exported
    for name, nargs in protos.items():
        fn = getattr(_abi.lib, name)
        assert fn.argtypes is not None and len(fn.argtypes) == nargs, (name, nargs, len(fn.argtypes or []))


def test_host_helpers_and_struct_layout():
    from materialist_b200 import _abi
    from oracle import oracle as orc
    assert C.sizeof(_abi.Cfg) == C.sizeof(orc.Cfg) == 10 * 4 + 48 * 4 + 8
    assert C.sizeof(_abi.HierDesc) == C.sizeof(orc.HierDesc)
    assert _abi.lib.mb200_strerror(0) == b"ok" and b"invalid" in _abi.lib.mb200_strerror(-1)
    assert _abi.lib.mb200_env_internal_width(32, _abi.ENV_FILE) == 33
    assert _abi.lib.mb200_env_internal_width(32, _abi.ENV_ASSIGNED) == 32
    O = orc.Oracle()
    for rx, ry in ((32, 16), (33, 16), (256, 128), (2049, 1024), (2, 2), (5, 9)):
        d = _abi.hier_describe(rx, ry); e = O.hier_describe(rx, ry)
        assert (d.n_levels, d.total_floats, list(d.lvl_off), list(d.lvl_w), list(d.lvl_h)) == \
               (e.n_levels, e.total_floats, list(e.lvl_off), list(e.lvl_w), list(e.lvl_h))
    with pytest.raises(ValueError):
        _abi.hier_describe(1, 5)
    c = _abi.Cfg(); c.H, c.W, c.row0, c.rows, c.filter = 100, 50, 10, 20, _abi.FILTER_GAUSSIAN
    first = C.c_int(-1)
    assert _abi.lib.mb200_fwd_partial_rows(C.byref(c), C.byref(first)) == 24 and first.value == 8
    assert _abi.lib.mb200_bwd_wpart_rows(C.byref(c), C.byref(first)) == 28 and first.value == 6
    c.row0, c.rows = 0, 100
    assert _abi.lib.mb200_fwd_partial_rows(C.byref(c), C.byref(first)) == 100 and first.value == 0
    assert _abi.lib.mb200_partial_stride(_abi.FILTER_GAUSSIAN) == 100 and _abi.lib.mb200_partial_stride(_abi.FILTER_BOX) == 4


def test_compute_entry_points_reject_null_arguments():
    from materialist_b200 import _abi
    c = _abi.Cfg(); d = _abi.HierDesc()
    assert _abi.lib.mb200_shade_fwd(C.byref(c), *([None] * 8), C.byref(d), None, None) == _abi.EINVAL
    assert _abi.lib.mb200_shade_bwd(C.byref(c), *([None] * 8), C.byref(d), *([None] * 6), 1, None) == _abi.EINVAL
    assert _abi.lib.mb200_env_prepare(None, 16, 32, 0, None, None, C.byref(d), None, None) == _abi.EINVAL
    assert _abi.lib.mb200_film_develop(C.byref(c), None, None, None) == _abi.EINVAL
    assert _abi.lib.mb200_bsdf_eval_pdf(C.byref(c), 4, *([None] * 11)) == _abi.EINVAL


def test_cpu_tensors_are_refused():
    import torch
    from materialist_b200 import _abi
    with pytest.raises(ValueError):
        _abi.ptr(torch.zeros(4))
