"""Mutation fuzz of the PNG decoder (kc_png.cu): corrupted, truncated and re-checksummed
files must come back as an error or as pixels, never as a crash or an out-of-bounds access.
Meant to be run against the AddressSanitizer build (no GPU needed: the codec is host code):

  scripts/asan_gpu.sh build
  KANTER_B200_LIB=$PWD/kanter_core_b200/build/asan/libkanter_b200.so \
  LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 \
  python scripts/probes/png_mutation_fuzz.py [iterations]
"""
import ctypes as C
import glob
import io
import os
import struct
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from kanter_core_b200._lib import lib  # noqa: E402

from PIL import Image  # noqa: E402


def decode(data):
    buf = (C.c_uint8 * len(data)).from_buffer_copy(data)
    out = C.c_void_p()
    w, h, ch = C.c_uint32(), C.c_uint32(), C.c_uint32()
    rc = lib.kc_png_decode(buf, len(data), C.byref(out), C.byref(w), C.byref(h), C.byref(ch))
    if rc == 0:
        n = w.value * h.value * ch.value
        if n:
            _ = bytes((C.c_uint8 * n).from_address(out.value))   # touch every byte handed back
        lib.kc_free(out)
    return rc


def chunks(data):
    pos, out = 8, []
    while pos + 12 <= len(data):
        (n,) = struct.unpack(">I", data[pos:pos + 4])
        out.append((pos, data[pos + 4:pos + 8], n))
        pos += 12 + n
    return out


def fix_crcs(data):
    b = bytearray(data)
    for pos, typ, n in chunks(data):
        if pos + 12 + n <= len(b):
            b[pos + 8 + n:pos + 12 + n] = struct.pack(">I", zlib.crc32(bytes(b[pos + 4:pos + 8 + n])) & 0xffffffff)
    return bytes(b)


def seeds():
    out = []
    for p in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "data", "*.png")))[:6]:
        out.append(open(p, "rb").read())
    r = np.random.default_rng(0)
    for mode, ch in (("L", 1), ("LA", 2), ("RGB", 3), ("RGBA", 4), ("P", 1), ("1", 1), ("I;16", 1)):
        a = r.integers(0, 256, size=(13, 17, ch) if ch > 1 else (13, 17), dtype=np.uint8)
        if mode == "I;16":
            im = Image.fromarray(r.integers(0, 65536, size=(13, 17), dtype=np.uint16))
        elif mode == "1":
            im = Image.fromarray((a > 127).astype(np.uint8) * 255).convert("1")
        elif mode == "P":
            im = Image.fromarray(a).convert("P")
        else:
            im = Image.fromarray(a, mode)
        for interlace in (False,):
            f = io.BytesIO()
            im.save(f, "PNG")
            out.append(f.getvalue())
    return out


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    r = np.random.default_rng(1)
    base = seeds()
    ok = err = 0
    for it in range(iters):
        d = bytearray(base[int(r.integers(len(base)))])
        how = int(r.integers(6))
        if how == 0:                      # flip a few bytes anywhere
            for _ in range(int(r.integers(1, 6))):
                d[int(r.integers(len(d)))] = int(r.integers(256))
        elif how == 1:                    # truncate
            d = d[:int(r.integers(len(d)))]
        elif how == 2:                    # corrupt the header fields, keep the checksums valid
            for _ in range(int(r.integers(1, 4))):
                d[16 + int(r.integers(13))] = int(r.integers(256))
            d = bytearray(fix_crcs(bytes(d)))
        elif how == 3:                    # corrupt chunk payloads, keep the checksums valid
            for _ in range(int(r.integers(1, 8))):
                d[int(r.integers(33, len(d)))] = int(r.integers(256))
            d = bytearray(fix_crcs(bytes(d)))
        elif how == 4:                    # lie about a chunk length
            cs = chunks(bytes(d))
            pos = cs[int(r.integers(len(cs)))][0]
            d[pos:pos + 4] = struct.pack(">I", int(r.integers(1 << 32)) if r.random() < 0.5 else int(r.integers(64)))
        else:                             # huge dimensions with a valid checksum; powers of two (2^31 x 2^31 wraps
            #                               h * (stride + 1) to 2 GiB in 64 bits) and sides just past the decoder's cap
            if r.random() < 0.5:
                wh = (1 << int(r.integers(12, 33)), 1 << int(r.integers(12, 33)))
                wh = tuple(min(v, 0xffffffff) + int(r.integers(-1, 2)) * (v < (1 << 32)) for v in wh)
            else:
                wh = (int(r.integers(1, 1 << 31)), int(r.integers(1, 1 << 31)))
            d[16:24] = struct.pack(">II", max(0, wh[0]) & 0xffffffff, max(0, wh[1]) & 0xffffffff)
            d = bytearray(fix_crcs(bytes(d)))
        rc = decode(bytes(d))
        ok += rc == 0
        err += rc != 0
    print("png mutation fuzz: %d inputs, %d decoded, %d rejected, no crash" % (iters, ok, err))


if __name__ == "__main__":
    main()
