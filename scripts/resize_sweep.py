#!/usr/bin/env python
"""Variants of the fused upsample kernel (kc_resize.cu), timed with CUDA events around each launch:
the round-1 cp.async/STG strip kernel against the tensor-map (TMA load + TMA store) kernel at every
(rows per group, rows per accumulator chunk, CTAs per SM) it is compiled for.

    python scripts/resize_sweep.py [--out profiles/resize_sweep_r02.json] [--math fast|exact] [--reps 20]
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import kanter_core_b200 as kc  # noqa: E402
from kanter_core_b200 import ResizeFilter  # noqa: E402
from kanter_core_b200._lib import call, kc_image  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--math", default="fast", choices=["fast", "exact"])
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--src", type=int, default=1024)
    ap.add_argument("--dst", type=int, default=8192)
    ap.add_argument("--all", action="store_true", help="every compiled variant (default: the shortlist)")
    args = ap.parse_args()
    tp = kc.TextureProcessor.new(math_mode=kc.MATH_FAST if args.math == "fast" else kc.MATH_EXACT)
    ctx = tp._ctx._h
    peak = 6531.6
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = float(json.load(open(pk))["hbm_gbs"])
    S, D = args.src, args.dst
    L = kc.SlotImage.from_planes(tp, [np.random.default_rng(4).random((S, S), dtype=np.float32)])

    def set_knobs(**kw):
        for k in ("resize_tma", "resize_g", "resize_rc", "resize_minb", "resize_threads", "resize_store"):
            call("kc_debug_set_tuning", k.encode(), int(kw.get(k, 0)))

    def time_filter(filt):
        def rs():
            o = kc_image()
            call("kc_resize", ctx, C.byref(L._im), D, D, int(filt), C.byref(o))
            return kc.SlotImage(tp._ctx, o)
        keep = None
        for _ in range(3):
            keep = rs()
        tp.synchronize()
        ms, n = C.c_double(), C.c_uint64()
        call("kc_context_set_timing", ctx, 1)
        call("kc_context_timing_read", ctx, -1, C.byref(ms), C.byref(n))
        for _ in range(args.reps):
            keep = rs()
        call("kc_context_timing_read", ctx, -1, C.byref(ms), C.byref(n))
        call("kc_context_set_timing", ctx, 0)
        del keep
        return ms.value / args.reps

    variants = [("strip_r01_128thr", dict(resize_tma=-1)), ("strip_r01_64thr", dict(resize_tma=-1, resize_threads=64))]
    if args.all:
        for g in (8, 16, 32):
            for rc in (4, 8, 16):
                if rc > g or (g == 32 and rc == 16):
                    continue
                for mb in ((6, 8) if g == 8 else ((4, 6) if (rc == 4 and g == 16) else (4,))):
                    for st in (-1, 1):
                        variants.append(("tma_%s_G%d_RC%d_MINB%d" % ("warp" if st > 0 else "block", g, rc, mb), dict(resize_tma=1, resize_g=g, resize_rc=rc, resize_minb=mb, resize_store=st)))
    else:
        for g, rc, mb, st in ((16, 4, 4, 1), (16, 8, 4, 1), (32, 4, 4, 1), (32, 8, 4, 1), (32, 4, 4, -1), (32, 8, 4, -1), (16, 8, 4, -1)):
            variants.append(("tma_%s_G%d_RC%d_MINB%d" % ("warp" if st > 0 else "block", g, rc, mb), dict(resize_tma=1, resize_g=g, resize_rc=rc, resize_minb=mb, resize_store=st)))
    rows = []
    alg = D * D * 4 + S * S * 4
    for name, kw in variants:
        set_knobs(**kw)
        row = {"variant": name}
        for filt in (ResizeFilter.Lanczos3, ResizeFilter.Gaussian, ResizeFilter.CatmullRom, ResizeFilter.Triangle, ResizeFilter.Nearest):
            ms = time_filter(filt)
            row[filt.name.lower()] = {"ms": ms, "GB/s": alg / (ms / 1e3) / 1e9, "frac_of_measured_peak": alg / (ms / 1e3) / 1e9 / peak}
        rows.append(row)
        print("%-22s " % name + "  ".join("%s %.4f ms (%.1f%%)" % (k[:4], v["ms"], 100 * v["frac_of_measured_peak"]) for k, v in row.items() if k != "variant"), file=sys.stderr)
    set_knobs()
    txt = json.dumps({"workload": "resize %d^2 -> %d^2, one plane, %s" % (S, D, args.math), "peak_GBs": peak, "rows": rows}, indent=1)
    if args.out:
        with open(args.out, "w") as f:
            f.write(txt)
    print(txt)
    tp.close()


if __name__ == "__main__":
    main()
