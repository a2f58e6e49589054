#!/usr/bin/env python
"""Per-kernel roofline table for every row of SURVEY.md section 8(a)/(d):
each node kernel timed alone with CUDA events on the launching stream
(kc_context_set_timing), inputs larger than L2, algorithmic bytes / time against
the measured HBM peak.  Writes a JSON list to stdout (and --out).

    python scripts/kernel_bench.py [--math fast|exact] [--reps 20] [--out profiles/kernels_r01.json]
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
from kanter_core_b200 import MixType, Node, NodeType, ResizeFilter, ResizePolicy, Size, SlotId  # noqa: E402
from kanter_core_b200._lib import call, kc_image  # noqa: E402


def peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return float(json.load(open(p))["hbm_gbs"]) if os.path.exists(p) else 6650.0


def rnd(seed, h, w):
    return np.random.default_rng(seed).random((h, w), dtype=np.float32)


class Timer:
    def __init__(self, tp):
        self.tp, self.ctx = tp, tp._ctx._h

    def run(self, fn, reps, warm=3):
        for _ in range(warm):
            fn()
        kc.jit_wait()          # a hot tape is compiled in the background: measure the kernel that serves it from then on
        fn()
        self.tp.synchronize()
        ms, n = C.c_double(), C.c_uint64()
        call("kc_context_set_timing", self.ctx, 1)
        call("kc_context_timing_read", self.ctx, -1, C.byref(ms), C.byref(n))  # reset
        for _ in range(reps):
            fn()
        call("kc_context_timing_read", self.ctx, -1, C.byref(ms), C.byref(n))
        call("kc_context_set_timing", self.ctx, 0)
        return ms.value / reps, n.value / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--math", default="fast", choices=["fast", "exact"])
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--out", default=None)
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    tp = kc.TextureProcessor.new(math_mode=kc.MATH_FAST if args.math == "fast" else kc.MATH_EXACT)
    T = Timer(tp)
    PEAK = peak()
    rows = []

    def record(name, ref, px, bytes_per_px, fn, reps=None):
        if args.only and args.only not in name:
            return
        ms, launches = T.run(fn, reps or args.reps)
        gbs = px * bytes_per_px / (ms / 1e3) / 1e9
        rows.append({"kernel": name, "reference": ref, "math": args.math, "pixels": px, "bytes_per_px": bytes_per_px,
                     "ms": ms, "launches": launches, "GB/s": gbs, "frac_of_measured_peak": gbs / PEAK,
                     "frac_of_8TBs": gbs / 8000.0, "Mpixel/s": px / ms / 1e3})
        print("%-44s %8.4f ms  %7.0f GB/s  %5.1f%% of measured  (%d launch/iter)" % (name, ms, gbs, 100 * gbs / PEAK, launches), file=sys.stderr)

    S = 4096
    A = kc.SlotImage.from_planes(tp, [rnd(1 + c, S, S) for c in range(4)])
    B = kc.SlotImage.from_planes(tp, [rnd(5 + c, S, S) for c in range(4)])
    Ag = kc.SlotImage.from_planes(tp, [rnd(11, S, S)])
    Bg = kc.SlotImage.from_planes(tp, [rnd(12, S, S)])

    def op_mix(op, l, r):
        def f():
            out = kc_image()
            call("kc_mix", tp._ctx._h, int(op), C.byref(l._im), C.byref(r._im), C.byref(out))
            img = kc.SlotImage(tp._ctx, out)
            # one fused launch for the lazy planes (alpha stays a constant descriptor here: 36 B/px)
            call("kc_image_materialize", tp._ctx._h, C.byref(img._im), 0)
        return f

    for op in MixType:
        record("mix_gray_%s_4096" % op.name.lower(), "src/node/mix.rs:136-192", S * S, 12, op_mix(op, Ag, Bg))
    for op in (MixType.Add, MixType.Multiply, MixType.Pow):
        record("mix_rgba_%s_4096" % op.name.lower(), "src/node/mix.rs:194-302", S * S, 36, op_mix(op, A, B))

    def rgb2gray():
        out = kc_image()
        call("kc_image_as_type", tp._ctx._h, C.byref(A._im), 0, C.byref(out))
        img = kc.SlotImage(tp._ctx, out)
        p = C.c_void_p()
        call("kc_plane_device_ptr", img.plane_handles()[0], C.byref(p))
    record("rgba_to_gray_4096", "src/slot_image.rs:242-253", S * S, 16, rgb2gray)

    dev_u8 = C.c_void_p()
    pl = C.c_void_p()
    call("kc_plane_create", tp._ctx._h, S, S, C.byref(pl))  # S*S*4 bytes of device memory for the RGBA8 result
    call("kc_plane_device_ptr", pl, C.byref(dev_u8))
    record("to_u8_rgba_4096", "src/slot_image.rs:146-170", S * S, 20,
           lambda: call("kc_image_to_u8_device", tp._ctx._h, C.byref(A._im), 0, dev_u8))
    record("to_u8_gray_4096", "src/slot_image.rs:146-157", S * S, 8,
           lambda: call("kc_image_to_u8_device", tp._ctx._h, C.byref(Ag._im), 0, dev_u8))
    record("to_u8_srgb_rgba_4096", "src/slot_image.rs:172-207", S * S, 20,
           lambda: call("kc_image_to_u8_device", tp._ctx._h, C.byref(A._im), 1, dev_u8))

    # configs[1]: fused Pow(Multiply(A,B),B) -> materialised RGBA
    lg = tp.new_live_graph()
    lg.embed_slot_data_with_id(kc.SlotData.new(0, 0, A), 0)
    lg.embed_slot_data_with_id(kc.SlotData.new(0, 0, B), 1)
    a = lg.add_node(Node.new(NodeType.Embed(0)))
    b = lg.add_node(Node.new(NodeType.Embed(1)))
    mul = lg.add_node(Node.new(NodeType.Mix(MixType.Multiply)))
    pw = lg.add_node(Node.new(NodeType.Mix(MixType.Pow)))
    out = lg.add_node(Node.new(NodeType.OutputRgba("out")))
    for (o, i, s) in [(a, mul, 0), (b, mul, 1), (mul, pw, 0), (b, pw, 1), (pw, out, 0)]:
        lg.connect(o, i, SlotId(0), SlotId(s))

    def cfg2():
        lg.replace_embedded(A, 0)
        lg.request(out)
    record("config2_fused_mul_pow_rgba_4096", "BASELINE.json configs[1]", S * S, 40, cfg2)

    # unfused (reference-shaped: one kernel per node, intermediates in HBM)
    tp.set_fuse(False)
    record("config2_UNFUSED_mul_pow_rgba_4096", "BASELINE.json configs[1], reference execution shape", S * S, 80, cfg2)
    tp.set_fuse(True)
    del lg

    # configs[2]: HeightToNormal 8192^2
    H = 8192
    hg = kc.SlotImage.from_planes(tp, [rnd(3, H, H)])

    def h2n():
        o = kc_image()
        call("kc_height_to_normal", tp._ctx._h, C.byref(hg._im), C.byref(o))
        kc.SlotImage(tp._ctx, o)
    record("height_to_normal_8192", "src/node/height_to_normal.rs:16-77", H * H, 16, h2n, reps=10)
    del hg

    # configs[3]: resize 1024^2 -> 8192^2, per plane (4 + 256 MiB) ; RGBA = 4 planes
    L = kc.SlotImage.from_planes(tp, [rnd(4, 1024, 1024)])
    for filt in (ResizeFilter.Lanczos3, ResizeFilter.Gaussian, ResizeFilter.CatmullRom, ResizeFilter.Triangle, ResizeFilter.Nearest):
        def rs(filt=filt):
            o = kc_image()
            call("kc_resize", tp._ctx._h, C.byref(L._im), 8192, 8192, int(filt), C.byref(o))
            kc.SlotImage(tp._ctx, o)
        record("resize_%s_1024_to_8192_plane" % filt.name.lower(), "src/shared.rs:155-201 (image 0.24.0 imageops::resize)",
               8192 * 8192, 4 + 4 / 64.0, rs, reps=10)
    L4 = kc.SlotImage.from_planes(tp, [rnd(40 + c, 1024, 1024) for c in range(4)])
    for filt in (ResizeFilter.Lanczos3, ResizeFilter.Gaussian):
        def rs4(filt=filt):
            o = kc_image()
            call("kc_resize", tp._ctx._h, C.byref(L4._im), 8192, 8192, int(filt), C.byref(o))
            kc.SlotImage(tp._ctx, o)
        record("resize_%s_1024_to_8192_rgba_node" % filt.name.lower(), "src/shared.rs:141-216, four planes in one launch",
               4 * 8192 * 8192, 4 + 4 / 64.0, rs4, reps=10)
    del L4

    # a constant plane somebody insists on seeing as pixels (kc_fill_kernel)
    def fill():
        cst = kc.SlotImage.from_value(tp, Size(8192, 8192), 0.25, False)        # a fresh descriptor: materialising turns it into pixels for good
        call("kc_image_materialize", tp._ctx._h, C.byref(cst._im), 1)
    record("fill_constant_plane_8192", "src/slot_image.rs:28-64 (vec![v; n])", 8192 * 8192, 4, fill, reps=10)

    big = kc.SlotImage.from_planes(tp, [rnd(6, 8192, 8192)])

    def down():
        o = kc_image()
        call("kc_resize", tp._ctx._h, C.byref(big._im), 1024, 1024, int(ResizeFilter.Lanczos3), C.byref(o))
        kc.SlotImage(tp._ctx, o)
    record("resize_lanczos3_8192_to_1024_plane", "src/shared.rs:155-201", 8192 * 8192, 4 + 4 / 64.0, down, reps=5)

    # u8 -> f32 planes (Image node upload path), 4096^2 RGBA8: kernel only
    u8 = np.random.default_rng(7).integers(0, 256, (S, S, 4), dtype=np.uint8)

    def from_u8():
        kc.SlotImage.from_u8(tp, u8)
    if not args.only or "from_u8" in args.only:
        for _ in range(2):
            from_u8()
        ms, n = C.c_double(), C.c_uint64()
        call("kc_context_set_timing", tp._ctx._h, 1)
        call("kc_context_timing_read", tp._ctx._h, -1, C.byref(ms), C.byref(n))
        for _ in range(5):
            from_u8()
        call("kc_context_timing_read", tp._ctx._h, 2, C.byref(ms), C.byref(n))
        call("kc_context_set_timing", tp._ctx._h, 0)
        t = ms.value / max(1, n.value)
        gbs = S * S * 20 / (t / 1e3) / 1e9
        rows.append({"kernel": "from_u8_rgba_4096", "reference": "src/shared.rs:16-56", "math": args.math, "pixels": S * S,
                     "bytes_per_px": 20, "ms": t, "launches": 1, "GB/s": gbs, "frac_of_measured_peak": gbs / PEAK,
                     "frac_of_8TBs": gbs / 8000.0, "Mpixel/s": S * S / t / 1e3})
        print("%-44s %8.4f ms  %7.0f GB/s  %5.1f%% of measured" % ("from_u8_rgba_4096", t, gbs, 100 * gbs / PEAK), file=sys.stderr)

    txt = json.dumps({"peak_GBs": PEAK, "rows": rows}, indent=1)
    if args.out:
        with open(args.out, "w") as f:
            f.write(txt)
    print(txt)
    tp.close()


if __name__ == "__main__":
    main()
