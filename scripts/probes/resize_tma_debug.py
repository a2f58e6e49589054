"""One resize through the tensor-map kernel; argv: filter sw sh dw dh [exact|fast] [g rc minb].  Used under compute-sanitizer."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import kanter_core_b200 as kc  # noqa: E402
import oracle  # noqa: E402
from kanter_core_b200._lib import call, kc_image  # noqa: E402

filt, sw, sh, dw, dh = [int(x) for x in sys.argv[1:6]]
mode = sys.argv[6] if len(sys.argv) > 6 else "exact"
knobs = [int(x) for x in sys.argv[7:10]] if len(sys.argv) > 9 else [0, 0, 0]
tp = kc.TextureProcessor.new(math_mode=kc.MATH_EXACT if mode == "exact" else kc.MATH_FAST)
call("kc_debug_set_tuning", b"resize_tma", 1)
for k, v in zip(("resize_g", "resize_rc", "resize_minb"), knobs):
    call("kc_debug_set_tuning", k.encode(), v)
p = np.random.default_rng(1).random((sh, sw), dtype=np.float32)
img = kc.SlotImage.from_planes(tp, [p])
out = kc_image()
call("kc_resize", tp._ctx._h, C.byref(img._im), dw, dh, filt, C.byref(out))
tp.synchronize()
got = kc.SlotImage(tp._ctx, out).planes()[0]
want = oracle.resize_plane(p, dw, dh, filt)
print("filter", filt, (sw, sh), "->", (dw, dh), mode, knobs, "bit-exact" if np.array_equal(got.view(np.uint32), want.view(np.uint32)) else "max err %g" % np.abs(got - want).max())
