"""The C++ PNG codec (kc_png.cu; replaces the `image` crate at read_slot_image
src/shared.rs:218-261 and the Write node src/node/write.rs:5-21) against Pillow: every
fixture of the reference's test data, every colour type the reference can take, palette,
sub-byte depths, tRNS, Adam7, and encode -> decode round trips.  Host-only: runs without a GPU."""
import ctypes as C
import glob
import io
import os
import struct
import zlib

import numpy as np
import pytest
from PIL import Image

from kanter_core_b200._lib import TexProError, call

DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "data")


def _take(p, w, h, ch):
    a = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), (h.value, w.value, ch.value)).copy()
    call("kc_free", p)
    return a


def dec_file(path):
    p, w, h, ch = C.c_void_p(), C.c_uint32(), C.c_uint32(), C.c_uint32()
    call("kc_png_decode_file", path.encode(), C.byref(p), C.byref(w), C.byref(h), C.byref(ch))
    return _take(p, w, h, ch)


def dec_mem(data):
    p, w, h, ch = C.c_void_p(), C.c_uint32(), C.c_uint32(), C.c_uint32()
    buf = (C.c_uint8 * len(data)).from_buffer_copy(data)
    call("kc_png_decode", buf, len(data), C.byref(p), C.byref(w), C.byref(h), C.byref(ch))
    return _take(p, w, h, ch)


def enc(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    h, w, ch = a.shape
    p, n = C.c_void_p(), C.c_size_t()
    call("kc_png_encode", a.ctypes.data, w, h, ch, C.byref(p), C.byref(n))
    data = C.string_at(p, n.value)
    call("kc_free", p)
    return data


def pil(a, mode):
    return Image.fromarray(a[:, :, 0] if a.shape[2] == 1 else a, mode)


def as3(img):
    b = np.asarray(img)
    return b[:, :, None] if b.ndim == 2 else b


def test_every_reference_fixture_decodes_like_pillow():
    files = sorted(glob.glob(os.path.join(DATA, "**", "*.png"), recursive=True))
    assert len(files) >= 30
    for f in files:
        assert np.array_equal(dec_file(f), as3(Image.open(f))), f


@pytest.mark.parametrize("mode,ch", [("L", 1), ("LA", 2), ("RGB", 3), ("RGBA", 4)])
@pytest.mark.parametrize("shape", [(1, 1), (7, 13), (64, 33)])
def test_colour_types(mode, ch, shape):
    a = np.random.default_rng(ch).integers(0, 256, shape + (ch,), dtype=np.uint8)
    bio = io.BytesIO()
    pil(a, mode).save(bio, "PNG")
    assert np.array_equal(dec_mem(bio.getvalue()), a)
    ours = enc(a)                                   # our encoder: we read it back, and so does Pillow
    assert np.array_equal(dec_mem(ours), a)
    assert np.array_equal(as3(Image.open(io.BytesIO(ours))), a)


def test_palette_and_sub_byte_depths():
    r = np.random.default_rng(9)
    rgb = r.integers(0, 256, (20, 31, 3), dtype=np.uint8)
    for colours in (2, 4, 16, 200):                 # 1-, 2-, 4- and 8-bit palettes
        pim = Image.fromarray(rgb, "RGB").quantize(colours)
        bio = io.BytesIO()
        pim.save(bio, "PNG")
        assert np.array_equal(dec_mem(bio.getvalue()), np.asarray(pim.convert("RGB"))), colours
    one = Image.fromarray((r.integers(0, 2, (9, 21)) * 255).astype(np.uint8), "L").convert("1")
    bio = io.BytesIO()
    one.save(bio, "PNG")
    assert np.array_equal(dec_mem(bio.getvalue())[:, :, 0], np.asarray(one.convert("L")))
    pa = Image.fromarray(rgb, "RGB").quantize(8)
    bio = io.BytesIO()
    pa.save(bio, "PNG", transparency=3)             # tRNS on a palette image -> RGBA8
    assert np.array_equal(dec_mem(bio.getvalue()), np.asarray(Image.open(io.BytesIO(bio.getvalue())).convert("RGBA")))


def _chunk(t, body):
    return struct.pack(">I", len(body)) + t + body + struct.pack(">I", zlib.crc32(t + body) & 0xffffffff)


def _adam7(a):
    """Interlaced PNG of an (h, w, c) uint8 array, filter type 0 everywhere."""
    h, w, c = a.shape
    color = {1: 0, 2: 4, 3: 2, 4: 6}[c]
    raw = b""
    for x0, y0, dx, dy in ((0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)):
        sub = a[y0::dy, x0::dx]
        if sub.size == 0:
            continue
        for row in sub:
            raw += b"\x00" + row.tobytes()
    ihdr = struct.pack(">IIBBBBB", w, h, 8, color, 0, 0, 1)
    return b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", ihdr) + _chunk(b"IDAT", zlib.compress(raw)) + _chunk(b"IEND", b"")


@pytest.mark.parametrize("shape", [(1, 1, 3), (5, 3, 4), (17, 23, 1), (64, 64, 2)])
def test_adam7(shape):
    a = np.random.default_rng(3).integers(0, 256, shape, dtype=np.uint8)
    data = _adam7(a)
    assert np.array_equal(as3(Image.open(io.BytesIO(data))), a)   # the fixture itself is a valid PNG
    assert np.array_equal(dec_mem(data), a)


def test_rejects_what_the_reference_cannot_take(tmp_path):
    a16 = (np.random.default_rng(1).integers(0, 65536, (4, 4))).astype(np.uint16)
    bio = io.BytesIO()
    Image.fromarray(a16).save(bio, "PNG")
    for bad in (bio.getvalue(), b"not a png at all", b""):
        with pytest.raises(TexProError) as e:
            dec_mem(bad if bad else b"\x00")
        assert e.value.kind == "Image"
    good = enc(np.zeros((3, 3, 4), np.uint8))
    corrupt = bytearray(good)
    corrupt[40] ^= 0xff                              # flips a byte inside IDAT: CRC mismatch
    with pytest.raises(TexProError):
        dec_mem(bytes(corrupt))
    with pytest.raises(TexProError) as e:
        dec_file(str(tmp_path / "missing.png"))
    assert e.value.kind == "Io"


def test_header_that_promises_terabytes_is_rejected_before_allocating():
    good = bytearray(enc(np.zeros((3, 3, 4), np.uint8)))
    good[16:24] = struct.pack(">II", 1 << 30, 1 << 30)                      # IHDR width, height
    good[29:33] = struct.pack(">I", zlib.crc32(bytes(good[12:29])) & 0xffffffff)  # keep the IHDR checksum valid
    with pytest.raises(TexProError) as e:
        dec_mem(bytes(good))
    assert e.value.kind == "Image"


@pytest.mark.parametrize("w,h", [(1 << 31, 1 << 31), (1 << 31, 1), (1, 1 << 31), ((1 << 16) + 1, 3), (1 << 16, 1 << 16), (0xffffffff, 0xffffffff)])
def test_header_sizes_that_wrap_64_bit_arithmetic_are_refused(w, h):
    """w = h = 2^31 made h * (stride + 1) wrap to 2 GiB and the pixel buffer to 0 bytes (advisor finding, round 1):
    the header is now bounded before any arithmetic, and a 9 MB stream of zeros behind it must not be walked."""
    good = bytearray(enc(np.zeros((3, 3, 4), np.uint8)))
    good[16:24] = struct.pack(">II", w, h)
    good[29:33] = struct.pack(">I", zlib.crc32(bytes(good[12:29])) & 0xffffffff)
    with pytest.raises(TexProError) as e:
        dec_mem(bytes(good))
    assert e.value.kind == "Image"
    # the same header in front of a stream that really inflates to 2 GiB of zeros (what the wrapped total asked for)
    if (w, h) == (1 << 31, 1 << 31):
        z = zlib.compressobj(9)
        body = b"".join(z.compress(bytes(1 << 24)) for _ in range(128)) + z.flush()
        ihdr = struct.pack(">II", w, h) + bytes([8, 6, 0, 0, 0])

        def chunk(t, b):
            return struct.pack(">I", len(b)) + t + b + struct.pack(">I", zlib.crc32(t + b) & 0xffffffff)
        bomb = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", ihdr) + chunk(b"IDAT", body) + chunk(b"IEND", b"")
        with pytest.raises(TexProError) as e:
            dec_mem(bomb)
        assert e.value.kind == "Image"


def test_encode_file(tmp_path):
    a = np.random.default_rng(4).integers(0, 256, (12, 10, 4), dtype=np.uint8)
    path = str(tmp_path / "out.png")
    call("kc_png_encode_file", path.encode(), a.ctypes.data, 10, 12, 4)
    assert np.array_equal(np.asarray(Image.open(path)), a)
    assert np.array_equal(dec_file(path), a)
