"""Per-kernel device time of the long-window resize SRC^2 -> DST^2 (default 8192 -> 1024; events around every launch):
    python scripts/probes/downsample_split.py [fast|exact] [SRC] [DST]"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import kanter_core_b200 as kc
from kanter_core_b200._lib import call
from kanter_core_b200.api import ResizeFilter, Size

mode = sys.argv[1] if len(sys.argv) > 1 else "fast"
SRC = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
DST = int(sys.argv[3]) if len(sys.argv) > 3 else SRC // 8
from kanter_core_b200 import _lib
tp = kc.TextureProcessor(math_mode=_lib.MATH_FAST if mode == "fast" else _lib.MATH_EXACT)
rng = np.random.default_rng(1)
img = kc.SlotImage.from_planes(tp, [rng.random((SRC, SRC), dtype=np.float32)])
flush = kc.SlotImage.from_planes(tp, [np.zeros((8192, 8192), np.float32)])
for _ in range(3):
    out = kc.resize(tp, img, Size(DST, DST), ResizeFilter.Lanczos3)
tp.synchronize()
ctx = tp._ctx._h
ms, n = C.c_double(), C.c_uint64()
call("kc_context_set_timing", ctx, 1)
call("kc_context_timing_read", ctx, -1, C.byref(ms), C.byref(n))
reps = 10
for _ in range(reps):
    out = kc.resize(tp, img, Size(DST, DST), ResizeFilter.Lanczos3)
tp.synchronize()
for kind, name in {4: "vertical march", 5: "horizontal pass"}.items():
    call("kc_context_timing_read", ctx, kind, C.byref(ms), C.byref(n))
    print("%-16s %.4f ms  (%d launches)" % (name, ms.value / reps, n.value))
