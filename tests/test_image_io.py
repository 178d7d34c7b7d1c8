"""CPU tests of the host's image output: PNG (sRGB-E, 8 bit, written bottom-up like
src/image.rs:138-166), EXR (half RGB in rec709-E, src/image.rs:167-195) and the Blender base64 stream."""
import base64
import ctypes as C
import os

import numpy as np
import pytest

from psychopath_b200 import abi

os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")

M = np.array([[2.6896551724, -1.2758620690, -0.4137931034], [-1.0221081721, 1.9782866167, 0.0438215555],
              [0.0612244898, -0.2244897959, 1.1632653061]], dtype=np.float32)


def _lib():
    h = abi.load_host()
    h.psyh_write_png.argtypes = [C.c_char_p, C.c_void_p, C.c_uint32, C.c_uint32]
    h.psyh_write_exr.argtypes = [C.c_char_p, C.c_void_p, C.c_uint32, C.c_uint32]
    h.psyh_bucket_base64.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_char_p, C.c_size_t]
    h.psyh_bucket_base64.restype = C.c_size_t
    return h


def _image(w=37, h=21):
    rng = np.random.default_rng(3)
    img = rng.random((h, w, 3)).astype(np.float32) * 0.8
    img[0, :, :] = 0.0          # top row black
    img[-1, :, 1] = 0.9         # bottom row greenish
    return img


def test_png_roundtrip(tmp_path):
    import cv2
    h = _lib()
    img = _image()
    p = str(tmp_path / "a.png")
    assert h.psyh_write_png(p.encode(), img.ctypes.data, img.shape[1], img.shape[0]) == 0
    got = cv2.imread(p, cv2.IMREAD_UNCHANGED)
    assert got.shape == (img.shape[0], img.shape[1], 4) and (got[..., 3] == 255).all()
    rgb = np.clip(img @ M.T, 0, None)
    srgb = np.where(rgb < 0.0031308, rgb * 12.92, 1.055 * np.power(np.maximum(rgb, 1e-12), 1 / 2.4) - 0.055)
    # upright: image row 0 (the reddish top row) is the first row of the file.  image.rs:147 reverses the rows only
    # because png_encode_mini takes them bottom-to-top and reverses them again.
    want = (np.clip(srgb, 0, 1) * 255).astype(np.uint8)
    assert np.abs(got[..., 2::-1][..., :3].astype(int) - want.astype(int)).max() <= 1


@pytest.mark.parametrize("compression", ["zip", "none"])
@pytest.mark.parametrize("size", [(37, 23), (64, 16), (5, 40)])
def test_exr_roundtrip(tmp_path, monkeypatch, compression, size):
    """Half RGB in rec709-E, ZIP-compressed in 16-line blocks (default) or plain scanlines; read back by OpenCV's
    OpenEXR.  Heights below, at and above a multiple of the block size."""
    import cv2
    h = _lib()
    img = _image(*size)
    monkeypatch.setenv("PSY_EXR_COMPRESSION", compression)
    p = str(tmp_path / "a.exr")
    assert h.psyh_write_exr(p.encode(), img.ctypes.data, img.shape[1], img.shape[0]) == 0
    got = cv2.imread(p, cv2.IMREAD_UNCHANGED)
    assert got is not None and got.shape[:2] == img.shape[:2]
    want = (img @ M.T).astype(np.float16).astype(np.float32)
    assert np.allclose(got[..., ::-1], want, atol=2e-3)


def test_exr_zip_is_smaller_on_a_smooth_image(tmp_path, monkeypatch):
    import os
    h = _lib()
    yy, xx = np.mgrid[0:128, 0:256].astype(np.float32)
    img = np.ascontiguousarray(np.stack([xx / 256, yy / 128, (xx + yy) / 384], axis=-1).astype(np.float32))
    sizes = {}
    for c in ("zip", "none"):
        monkeypatch.setenv("PSY_EXR_COMPRESSION", c)
        p = str(tmp_path / f"{c}.exr")
        assert h.psyh_write_exr(p.encode(), img.ctypes.data, 256, 128) == 0
        sizes[c] = os.path.getsize(p)
    assert sizes["zip"] < 0.5 * sizes["none"]


def test_bucket_base64_stream():
    h = _lib()
    img = _image(5, 3)
    need = h.psyh_bucket_base64(img.ctypes.data, 5, 3, None, 0)
    buf = C.create_string_buffer(need)
    n = h.psyh_bucket_base64(img.ctypes.data, 5, 3, buf, need)
    data = np.frombuffer(base64.b64decode(buf.raw[:n]), dtype=np.float32).reshape(3, 5, 4)
    assert (data[..., 3] == 1.0).all()
    assert np.allclose(data[..., :3], img @ M.T, atol=1e-6)
