"""Leave the process while a background specialisation is still compiling: must exit cleanly (rc 0)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import kanter_core_b200 as kc
from kanter_core_b200 import MixType

tp = kc.TextureProcessor.new(math_mode=kc.MATH_FAST)
r = np.random.default_rng(0)
imgs = [kc.SlotImage.from_planes(tp, [r.random((1024, 1024), dtype=np.float32)]) for _ in range(3)]
for _ in range(3):
    x = kc.mix(tp, MixType.Multiply, imgs[0], imgs[1])
    x = kc.mix(tp, MixType.Subtract, x, imgs[2])
    x = kc.mix(tp, MixType.Divide, x, imgs[1])
    x = kc.mix(tp, MixType.Add, x, imgs[0])
    x.planes()
t0 = time.perf_counter()
print("leaving with a compile in flight")
