#!/usr/bin/env python
"""Host<->device copy rates through the library's own upload/download entry points
(pinned buffers from kc_host_alloc), step by step, to see what bounds bench.py's e2e."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kanter_core_b200 as kc

S = 4096
tp = kc.TextureProcessor.new(math_mode=kc.MATH_FAST)
host = [kc.pinned_empty((S, S)) for _ in range(8)]
for h in host:
    h[...] = 0.5
out = kc.pinned_empty((S, S, 4), np.uint8)
for it in range(8):
    t0 = time.perf_counter()
    a = kc.SlotImage.from_planes(tp, host[:4], sync=False)
    b = kc.SlotImage.from_planes(tp, host[4:], sync=False)
    t1 = time.perf_counter()
    tp.synchronize()
    t2 = time.perf_counter()
    m = kc.mix(tp, kc.MixType.Multiply, a, b)
    import ctypes as C
    from kanter_core_b200._lib import call
    call("kc_image_to_u8", tp._ctx._h, C.byref(m._im), 0, out.ctypes.data)
    t3 = time.perf_counter()
    del a, b, m
    print("iter %d: enqueue uploads %.2f ms, H2D 512 MiB done after %.2f ms (%.1f GB/s), fused mul+to_u8 + D2H 64 MiB %.2f ms" % (
        it, (t1 - t0) * 1e3, (t2 - t0) * 1e3, 0.536870912 / (t2 - t0), (t3 - t2) * 1e3))
tp.close()
