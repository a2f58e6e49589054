"""Where the time of one small-graph evaluation goes with evaluation replay on: Python calls, dirty propagation, the replay
itself, the device.  The 32-node graph of configs[4] at 256^2 (argv[1])."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import kanter_core_b200 as kc  # noqa: E402
from tests import graphs  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 256
tp = kc.TextureProcessor.new()
g, out = graphs.config5_graph(size)
inputs = graphs.config5_inputs(77, size)
for replay in (False, True):
    lg = tp.new_live_graph()
    lg.set_node_graph(g)
    lg.set_replay(replay)
    imgs = [kc.SlotImage.from_planes(tp, planes) for planes in inputs]
    for eid, img in enumerate(imgs):
        lg.embed_slot_data_with_id(kc.SlotData.new(0, 0, img), eid)

    def emb():
        for eid, img in enumerate(imgs):
            lg.replace_embedded(img, eid)

    for _ in range(20):
        emb(); lg.request(out)
    tp.synchronize()
    N = 500
    t_emb = t_req = 0.0
    t0 = time.perf_counter()
    for _ in range(N):
        a = time.perf_counter(); emb(); b = time.perf_counter(); lg.request(out); c = time.perf_counter()
        t_emb += b - a; t_req += c - b
    t1 = time.perf_counter()
    tp.synchronize()
    t2 = time.perf_counter()
    # device alone: synchronise after every evaluation
    t3 = time.perf_counter()
    for _ in range(100):
        emb(); lg.request(out); tp.synchronize()
    t4 = time.perf_counter()
    print("replay=%s  per evaluation: 3 x replace_embedded %.1f us, request %.1f us, host loop %.1f us, with final sync %.1f us; "
          "evaluate + synchronize each time %.1f us  %s" % (replay, t_emb / N * 1e6, t_req / N * 1e6, (t1 - t0) / N * 1e6, (t2 - t0) / N * 1e6,
                                                          (t4 - t3) / 100 * 1e6, lg.replay_stats()))
tp.close()
