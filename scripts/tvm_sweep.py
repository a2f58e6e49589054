#!/usr/bin/env python
"""Sweep of the fused elementwise kernel's launch configuration (tile size V, resident CTAs,
pipeline stages) against the number of source planes of a segment, on 4096^2 planes:
out = s0 + s1 + ... + s(ns-1), algorithmic bytes (ns + 1) * 4 per pixel.  The table this
prints is what pick_config (kc_kernels.cu) and the planner's source cap (kc_fusion.cu) are
tuned from.   python scripts/tvm_sweep.py [--out profiles/tvm_sweep_rNN.json] [--pow]
"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kanter_core_b200 as kc  # noqa: E402
from kanter_core_b200 import MixType  # noqa: E402
from kanter_core_b200._lib import call, kc_image  # noqa: E402


def main():
    tp = kc.TextureProcessor.new(math_mode=kc.MATH_FAST)
    ctx = tp._ctx._h
    call("kc_debug_set_tuning", b"jit", -1)     # this sweep is about the interpreter kernel's configurations
    S = 4096
    r = np.random.default_rng(5)
    planes = [kc.SlotImage.from_planes(tp, [r.random((S, S), dtype=np.float32)]) for _ in range(8)]
    op2 = MixType.Pow if "--pow" in sys.argv else MixType.Multiply

    def build(ns):
        acc = planes[0]
        for k in range(1, ns):
            acc = kc.mix(tp, MixType.Add if k % 2 else op2, acc, planes[k])
        return acc

    def timed(ns, reps=10):
        ms, n = C.c_double(), C.c_uint64()
        for it in range(reps + 2):
            if it == 2:
                call("kc_context_set_timing", ctx, 1)
                call("kc_context_timing_read", ctx, -1, C.byref(ms), C.byref(n))
            img = build(ns)
            call("kc_image_materialize", ctx, C.byref(img._im), 0)
            del img
        call("kc_context_timing_read", ctx, 0, C.byref(ms), C.byref(n))
        call("kc_context_set_timing", ctx, 0)
        v, c, s = C.c_int32(), C.c_int32(), C.c_int32()
        call("kc_debug_last_tile_config", C.byref(v), C.byref(c), C.byref(s))
        return ms.value / max(1, n.value), n.value / reps, (v.value, c.value, s.value)

    rows = []
    for ns in range(2, 9):
        best = None
        for v in (4, 2, 1):
            for ctas in (3, 2, 1):
                for st in (2, 3, 4):
                    call("kc_debug_set_tuning", b"tile_v", v)
                    call("kc_debug_set_tuning", b"ctas", ctas)
                    call("kc_debug_set_tuning", b"stages", st)
                    t, launches, used = timed(ns)
                    if used != (v, ctas, st) or launches != 1:
                        continue            # did not fit shared memory: the library fell back
                    gbs = (ns + 1) * 4 * S * S / (t / 1e3) / 1e9
                    rows.append({"ns": ns, "v": v, "ctas": ctas, "stages": st, "ms": t, "GBs": gbs})
                    if best is None or gbs > best["GBs"]:
                        best = rows[-1]
        for k in (b"tile_v", b"ctas", b"stages"):
            call("kc_debug_set_tuning", k, 0)
        t, launches, used = timed(ns)
        gbs = (ns + 1) * 4 * S * S / (t / 1e3) / 1e9
        print("ns=%d  default %s %.4f ms %5.0f GB/s | best v=%d ctas=%d stages=%d %.4f ms %5.0f GB/s" %
              (ns, used, t, gbs, best["v"], best["ctas"], best["stages"], best["ms"], best["GBs"]), flush=True)
        by = {}
        for row in rows:
            if row["ns"] == ns:
                k = (row["v"], row["ctas"])
                if k not in by or row["GBs"] > by[k]["GBs"]:
                    by[k] = row
        print("      " + "  ".join("v%dc%d:%4.0f(s%d)" % (k[0], k[1], r_["GBs"], r_["stages"]) for k, r_ in sorted(by.items(), reverse=True)), flush=True)
    if "--out" in sys.argv:
        json.dump(rows, open(sys.argv[sys.argv.index("--out") + 1], "w"), indent=0)
    tp.close()


if __name__ == "__main__":
    main()
