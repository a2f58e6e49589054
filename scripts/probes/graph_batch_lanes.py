"""configs[4] (64 x 32-node graphs at 4096^2, evaluation replay) as a function of the concurrent-section lanes and of the
number of distinct live graphs that share the batch:  python scripts/probes/graph_batch_lanes.py [lanes:distinct ...]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import kanter_core_b200 as kc
from kanter_core_b200 import _lib
import bench_workloads as bw

peak = 6531.6
try:
    peak = json.load(open(os.path.join(os.path.dirname(bw.__file__), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
tp = kc.TextureProcessor(math_mode=_lib.MATH_FAST)
env = bw.Env(kc, tp, 0, 1, lambda: None, lambda x: x, peak)
for spec in sys.argv[1:] or ["1:2", "2:2", "2:4", "4:4"]:
    # lanes:distinct[:key=value,...]  -- the rest are kc_debug_set_tuning knobs for this run
    parts = spec.split(":")
    lanes, distinct = int(parts[0]), int(parts[1])
    knobs = dict(kv.split("=") for kv in parts[2].split(",")) if len(parts) > 2 else {}
    for k in ("ctas", "stages", "tile_v", "smem_cap_kb", "smem_cap_exact_kb"):
        _lib.call("kc_debug_set_tuning", k.encode(), int(knobs.get(k, 0)))
    r = bw.wl_graph_batch(env, distinct=distinct, lanes=lanes)
    print(spec, "lanes %d distinct %d: %.4f ms/graph  frac %.3f  parity %s  %s" % (
        lanes, distinct, r.get("ms_per_graph_per_gpu", -1), r.get("roofline", {}).get("frac", -1), r.get("parity", {}).get("ok"),
        r.get("unavailable", "")), flush=True)
    kc.jit_wait(1)
    _lib.call("kc_context_trim", tp._ctx._h)
