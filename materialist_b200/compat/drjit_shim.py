"""`drjit` stand-in (see materialist_b200.compat): the reference only uses dr.wrap_ad and dr.set_flag at script level."""
import types


class _JitFlag:
    VCallRecord = "VCallRecord"
    LoopRecord = "LoopRecord"


def module():
    m = types.ModuleType("drjit")
    m.__doc__ = __doc__
    m.JitFlag = _JitFlag
    m.flags = {}

    def set_flag(flag, value):
        m.flags[flag] = value
    m.set_flag = set_flag

    def wrap_ad(source="torch", target="drjit"):
        """inverse_img_w_mi.py:59,:69 — torch tensors in, torch tensors out with a working backward: the operator already is a
        torch.autograd.Function, so the decorator has nothing to convert."""
        if source != "torch":
            raise ValueError("only source='torch' is supported")
        return lambda fn: fn
    m.wrap_ad = wrap_ad
    return m
