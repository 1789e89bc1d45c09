"""CPU (host code, no GPU): the library's Radiance .hdr / OpenEXR readers and writers (csrc/mb200_io.cu) —
the stand-in for mi.Bitmap / mi.util.write_bitmap (misc.py:99-111, mi_plugin.py:701-739).

Checker: OpenCV's independent decoders / encoders (bit-exact: the formats are lossless for float data; RGBE quantisation is
defined by the format), on
  * two files written by the REFERENCE's own Mitsuba run and shipped with it (an RGBE envmap, a PIZ-compressed FLOAT EXR),
    pinned by the sha256 of their decoded pixels (computed with OpenCV when the fixture was made), and
  * files OpenCV writes at test time: every supported compression x pixel type, odd sizes (wavelet / block edge cases),
    constant images (Huffman run-length symbol), 1 / 3 / 4 channels.
"""
import hashlib
import os

import numpy as np
import pytest

os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")
cv2 = pytest.importorskip("cv2")

from materialist_b200 import imageio as io  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def cv_read(path):
    a = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    assert a is not None, path
    if a.ndim == 3:
        a = a[..., [2, 1, 0] + ([3] if a.shape[2] == 4 else [])]
    return np.ascontiguousarray(a, np.float32)


def cv_write(path, img, flags=()):
    a = img
    if a.ndim == 3:
        a = a[..., [2, 1, 0] + ([3] if a.shape[2] == 4 else [])]
    assert cv2.imwrite(path, np.ascontiguousarray(a), list(flags))


@pytest.mark.parametrize("name,shape,sha", [
    ("indoor_envmap_mitsuba.hdr", (16, 32, 3), "ec8e016e5b47cc8134876271a912eb4d216229d7f111fc039a6ea9180aaefa70"),
    ("jinjya_roughness_mitsuba_piz.exr", (512, 512), "40df7f48670a32b8aea49c85e5f3d179aa4fdb5148e612fe24d36c7e5ef9727c"),
])
def test_files_written_by_the_reference(name, shape, sha):
    path = os.path.join(GOLD, name)
    a = io.read_bitmap(path)
    assert a.shape == shape and a.dtype == np.float32
    assert hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest() == sha
    assert np.array_equal(a, cv_read(path))
    assert io.image_info(path) == (shape[0], shape[1], 1 if len(shape) == 2 else shape[2])


def _images():
    rs = np.random.RandomState(5)
    yield "hdr-range", (np.exp(rs.randn(37, 53, 3) * 3)).astype(np.float32)           # 1e-6 .. 1e5
    yield "smooth", np.stack(np.meshgrid(np.linspace(0, 1, 64), np.linspace(0, 2, 48)), -1).astype(np.float32)[..., [0, 1, 0]] + 0.25
    yield "constant", np.full((33, 17, 3), 0.5, np.float32)
    yield "one-channel", rs.rand(40, 31).astype(np.float32)
    yield "rgba", rs.rand(19, 70, 4).astype(np.float32)
    yield "tiny", rs.rand(1, 1, 3).astype(np.float32)
    yield "negative", rs.randn(35, 35, 3).astype(np.float32)


EXR_COMP = {"none": cv2.IMWRITE_EXR_COMPRESSION_NO, "zips": cv2.IMWRITE_EXR_COMPRESSION_ZIPS, "zip": cv2.IMWRITE_EXR_COMPRESSION_ZIP,
            "piz": cv2.IMWRITE_EXR_COMPRESSION_PIZ}


@pytest.mark.parametrize("comp", list(EXR_COMP))
@pytest.mark.parametrize("half", [False, True])
def test_exr_reader_matches_opencv(tmp_path, comp, half):
    for tag, img in _images():
        p = str(tmp_path / f"{tag}_{comp}_{int(half)}.exr")
        cv_write(p, img, (cv2.IMWRITE_EXR_COMPRESSION, EXR_COMP[comp], cv2.IMWRITE_EXR_TYPE, cv2.IMWRITE_EXR_TYPE_HALF if half else cv2.IMWRITE_EXR_TYPE_FLOAT))
        ref = cv_read(p)
        got = io.read_bitmap(p)
        assert got.shape == ref.shape, (tag, got.shape, ref.shape)
        assert np.array_equal(got, ref), (tag, comp, half)
        if not half:
            assert np.array_equal(got, img)                                      # FLOAT files are lossless


def test_exr_writer_read_back_by_opencv_and_by_us(tmp_path):
    for tag, img in _images():
        p = str(tmp_path / f"w_{tag}.exr")
        io.write_bitmap(p, img)
        assert np.array_equal(cv_read(p), img), tag
        assert np.array_equal(io.read_bitmap(p), img), tag
    import torch
    t = torch.rand(8, 9, 3)
    io.write_bitmap(str(tmp_path / "t.exr"), t)
    assert np.array_equal(io.read_bitmap(str(tmp_path / "t.exr")), t.numpy())


def test_hdr_reader_and_writer_match_opencv(tmp_path):
    rs = np.random.RandomState(9)
    for tag, img in (("env", (0.2 + np.exp(rs.randn(16, 32, 3))).astype(np.float32)),
                     ("sun", np.concatenate([np.full((8, 64, 3), 0.01, np.float32), np.full((8, 64, 3), 56832.0, np.float32)], 0)),   # runs -> RLE
                     ("narrow", rs.rand(5, 7, 3).astype(np.float32)),                                   # width < 8: flat scanlines
                     ("zeros", np.zeros((4, 16, 3), np.float32))):
        p_cv, p_us = str(tmp_path / f"cv_{tag}.hdr"), str(tmp_path / f"us_{tag}.hdr")
        cv_write(p_cv, img)
        assert np.array_equal(io.read_bitmap(p_cv), cv_read(p_cv)), tag          # our reader on OpenCV's file
        io.write_bitmap(p_us, img)
        a, b = io.read_bitmap(p_us), cv_read(p_us)                                 # our file: both readers agree ...
        assert np.array_equal(a, b), tag
        assert np.array_equal(a, cv_read(p_cv)), tag                               # ... and the RGBE quantisation equals OpenCV's encoder
        ok = img.max(-1) > 1e-30
        assert np.all(np.abs(a - img)[ok] <= img.max(-1, keepdims=True).repeat(3, -1)[ok] / 128 + 1e-30)   # 8-bit shared-exponent mantissa


def test_png_reader_and_writer_match_opencv(tmp_path):
    """bg.png / mask.png of the editing scripts (mi_plugin.py:717-731, plt.imread -> floats in [0, 1])."""
    rs = np.random.RandomState(4)
    cases = [("gray8", rs.randint(0, 256, (21, 34)).astype(np.uint8)), ("rgb8", rs.randint(0, 256, (33, 17, 3)).astype(np.uint8)),
             ("rgba8", rs.randint(0, 256, (9, 40, 4)).astype(np.uint8)), ("gray16", rs.randint(0, 65536, (12, 13)).astype(np.uint16)),
             ("rgb16", rs.randint(0, 65536, (14, 15, 3)).astype(np.uint16)), ("smooth", (np.add.outer(np.arange(64), np.arange(48)) % 256).astype(np.uint8))]
    for tag, img in cases:
        p = str(tmp_path / f"{tag}.png")
        a = img if img.ndim == 2 else img[..., [2, 1, 0] + ([3] if img.shape[2] == 4 else [])]
        assert cv2.imwrite(p, np.ascontiguousarray(a))                           # OpenCV picks its own row filters
        got = io.read_bitmap(p)
        ref = img.astype(np.float32) / (255.0 if img.dtype == np.uint8 else 65535.0)
        assert got.shape == ref.shape and np.array_equal(got, ref), tag
    for C in (1, 3, 4):
        img = rs.rand(19, 23, C).astype(np.float32) if C > 1 else rs.rand(19, 23).astype(np.float32)
        p = str(tmp_path / f"w{C}.png")
        io.write_bitmap(p, img, srgb=False)                                     # data file: raw values
        back = cv2.imread(p, cv2.IMREAD_UNCHANGED)
        if back.ndim == 3:
            back = back[..., [2, 1, 0] + ([3] if back.shape[2] == 4 else [])]
        assert np.array_equal(back, np.floor(np.clip(img, 0, 1) * 255.0 + 0.5).astype(np.uint8))
        assert np.array_equal(io.read_bitmap(p), back.astype(np.float32) / 255.0)
        io.write_bitmap(p, img)                                                 # default, as mi.util.write_bitmap: sRGB transfer, alpha linear
        back = cv2.imread(p, cv2.IMREAD_UNCHANGED)
        if back.ndim == 3:
            back = back[..., [2, 1, 0] + ([3] if back.shape[2] == 4 else [])]
        x = np.clip(img, 0, 1).astype(np.float64)
        want = np.where(x <= 0.0031308, 12.92 * x, 1.055 * x ** (1 / 2.4) - 0.055)
        if C == 4:
            want[..., 3] = x[..., 3]
        assert np.abs(back.astype(np.float64) - want * 255.0).max() <= 0.5 + 1e-3


def test_errors():
    with pytest.raises(ValueError):
        io.read_bitmap("/nonexistent/file.exr")
    with pytest.raises(Exception):
        io.write_bitmap("/tmp/x.jpg", np.zeros((2, 2, 3), np.float32))            # only .hdr / .exr / .png
    with pytest.raises(ValueError):
        io.write_bitmap("/tmp/x.hdr", np.zeros((2, 2), np.float32))               # RGBE needs 3 channels


def test_load_estimated_brdf_goes_through_the_native_readers(tmp_path):
    """gbuffer.load_estimated_brdf (mi_plugin.py:701-739) on a best_results-style folder written with write_bitmap."""
    from materialist_b200 import gbuffer
    rs = np.random.RandomState(1)
    a, r, m, n = rs.rand(12, 12, 3).astype(np.float32), rs.rand(12, 12).astype(np.float32), rs.rand(12, 12).astype(np.float32), rs.rand(12, 12, 3).astype(np.float32)
    for name, img in (("albedo", a), ("roughness", r), ("metallic", m), ("normal", n)):
        io.write_bitmap(str(tmp_path / f"{name}.exr"), img)
    bg = rs.rand(7, 9, 4).astype(np.float32); mk = (rs.rand(12, 12, 4) > 0.5).astype(np.float32); env = (rs.rand(4, 8, 3) + 0.5).astype(np.float32)
    io.write_bitmap(str(tmp_path / "bg.png"), bg, srgb=False); io.write_bitmap(str(tmp_path / "mask.png"), mk, srgb=False)
    io.write_bitmap(str(tmp_path / "envmap.hdr"), env)
    mat = gbuffer.load_estimated_brdf(str(tmp_path))
    assert mat["bg"].shape == (12, 12, 3) and mat["mask"].dtype == np.bool_ and np.array_equal(mat["mask"], mk[..., 0] > 0.5)
    assert mat["envmap"].shape == (4, 8, 3) and np.allclose(mat["envmap"], env, rtol=1 / 64)
    assert np.array_equal(mat["albedo"], a) and np.allclose(mat["roughness"][..., 0], r * 0.95 + 0.05) and np.array_equal(mat["metallic"][..., 0], m)
    assert np.array_equal(mat["normal"], n)
