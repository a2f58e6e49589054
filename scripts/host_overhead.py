#!/usr/bin/env python
"""Host-side cost of one configs[1] evaluation step (graph bookkeeping + fusion planning +
launch), measured against the device time: tells whether the GPU queue stays fed."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kanter_core_b200 as kc
from kanter_core_b200 import MixType, Node, NodeType, SlotId

S = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
tp = kc.TextureProcessor.new(math_mode=kc.MATH_FAST)
r = np.random.default_rng(0)
A = kc.SlotImage.from_planes(tp, [r.random((S, S), dtype=np.float32) for _ in range(4)])
B = kc.SlotImage.from_planes(tp, [r.random((S, S), dtype=np.float32) for _ in range(4)])
lg = tp.new_live_graph()
lg.embed_slot_data_with_id(kc.SlotData.new(0, 0, A), 0)
lg.embed_slot_data_with_id(kc.SlotData.new(0, 0, B), 1)
a = lg.add_node(Node.new(NodeType.Embed(0))); b = lg.add_node(Node.new(NodeType.Embed(1)))
mul = lg.add_node(Node.new(NodeType.Mix(MixType.Multiply))); pw = lg.add_node(Node.new(NodeType.Mix(MixType.Pow)))
out = lg.add_node(Node.new(NodeType.OutputRgba("out")))
for (o, i, s) in [(a, mul, 0), (b, mul, 1), (mul, pw, 0), (b, pw, 1), (pw, out, 0)]:
    lg.connect(o, i, SlotId(0), SlotId(s))

def step():
    lg.replace_embedded(A, 0); lg.replace_embedded(B, 1); lg.request(out)

for _ in range(20): step()
tp.synchronize()
N = 300
t0 = time.perf_counter()
for _ in range(N): step()
t1 = time.perf_counter()
tp.synchronize()
t2 = time.perf_counter()
print("size %d: host enqueue %.1f us/step, total %.1f us/step" % (S, (t1 - t0) / N * 1e6, (t2 - t0) / N * 1e6))
tp.close()
