"""On-disk image formats of the path through the library's own readers / writers (csrc/mb200_io.cu): the stand-in for
`mi.Bitmap(path)` and `mi.util.write_bitmap(path, img)` (myutils/misc.py:99-111, myutils/mi_plugin.py:701-739,
render_final.py:182-202).  Radiance .hdr and OpenEXR (scanline; NONE / ZIPS / ZIP / PIZ) -> float32 (H, W, C) numpy arrays
in R,G,B(,A) order, linear file values."""
import ctypes as C

import numpy as np

from . import _abi


def image_info(path):
    H, W, Cn = C.c_int(), C.c_int(), C.c_int()
    _abi.check(_abi.lib.mb200_image_info(str(path).encode(), C.byref(H), C.byref(W), C.byref(Cn)), f"mb200_image_info({path})")
    return H.value, W.value, Cn.value


def read_bitmap(path):
    """`np.array(mi.Bitmap(path))`: (H, W, 3|4) or (H, W) for single-channel files, float32."""
    H, W, Cn = image_info(path)
    out = np.empty((H, W, Cn), np.float32)
    _abi.check(_abi.lib.mb200_image_read(str(path).encode(), out.ctypes.data_as(C.c_void_p), H, W, Cn), f"mb200_image_read({path})")
    return out[..., 0] if Cn == 1 else out


def write_bitmap(path, img, srgb=True):
    """`mi.util.write_bitmap(path, img)` for .hdr (RGBE, 3 channels), .exr (ZIP, float32; 1, 3 or 4 channels) and 8-bit .png.
    PNG: like Mitsuba, the colour channels go through the sRGB transfer curve (alpha stays linear); `srgb=False` stores the clamped
    values as they are (data files: masks, backgrounds that must read back unchanged).  .hdr / .exr are linear float formats."""
    if hasattr(img, "detach"):
        img = img.detach().cpu().numpy()
    a = np.ascontiguousarray(img, dtype=np.float32)
    if a.ndim == 2:
        a = a[..., None]
    if a.ndim != 3:
        raise ValueError("image must be (H, W) or (H, W, C)")
    H, W, Cn = a.shape
    _abi.check(_abi.lib.mb200_image_write_ex(str(path).encode(), a.ctypes.data_as(C.c_void_p), H, W, Cn, _abi.IMG_SRGB if srgb else 0),
               f"mb200_image_write({path})")
