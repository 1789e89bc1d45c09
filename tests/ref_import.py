"""TEST INFRASTRUCTURE: import the reference's OWN modules (unmodified, from MATERIALIST_REF=/root/reference) on top of the
materialist_b200.compat stand-ins for mitsuba / drjit, with inert stubs for the third-party packages this image lacks and for the
reference's out-of-scope pre-processing (MaterialNet, mesh reconstruction).  Used by tests/test_reference_caller.py."""
import math
import os
import sys
import types

import numpy as np

REF = os.environ.get("MATERIALIST_REF", "/root/reference")


class _Any:
    def __init__(self, *a, **k): pass
    def __call__(self, *a, **k): return _Any()
    def __getattr__(self, k): return _Any()
    def __iter__(self): return iter(())


def _stub(name):
    m = types.ModuleType(name)

    def ga(n):
        if n.startswith("__"):
            raise AttributeError(n)
        return _Any()
    m.__getattr__ = ga
    m.monkey_patch = lambda *a, **k: None
    sys.modules[name] = m
    return m


def import_reference_caller():
    """Returns the reference's inverse_img_w_mi module, imported under the compat stand-ins (cwd-independent)."""
    import materialist_b200.compat as compat
    compat.install()
    for name in ("open3d", "lovely_tensors", "matplotlib", "matplotlib.pyplot", "imageio", "huggingface_hub", "pytorch_lightning",
                 "Material_net", "Material_net.dpt", "myutils.mesh_recon"):
        if name in ("Material_net", "Material_net.dpt", "myutils.mesh_recon"):
            _stub(name)                      # the reference's one-shot pre-processing: out of scope, never called here
            continue
        try:
            __import__(name)
        except Exception:
            _stub(name)
    if "matplotlib" in sys.modules and not hasattr(sys.modules["matplotlib"], "__path__"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    try:
        import torchvision.utils  # noqa: F401
    except Exception:
        _stub("torchvision"); _stub("torchvision.utils")
    if not hasattr(np, "math"):
        np.math = math
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import inverse_img_w_mi as ref
    return ref
