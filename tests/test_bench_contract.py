"""The reference arm of bench.py (`--impl reference`: the CPU restatement of the reference's engine on
the host cores) needs no GPU, so its side of the contract is checked here: one JSON line with the
agreed keys from rank 0, nothing but exit code 0 from the other ranks of a torchrun launch."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(env_extra):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "4", "--warmup", "2"],
                          cwd=ROOT, env=env, capture_output=True, text=True, timeout=300)


def test_reference_arm_prints_one_contract_line():
    p = run({})
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "graph_eval_mpixel_per_s" and d["unit"] == "Mpixel/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["value"] > 0
    assert d["steps"] == 4 and d["warmup"] == 2          # --steps / --warmup are honoured as given (same_steps with our arm)
    assert d["config"]["workload"].startswith("configs[1]")
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["sample"] and cb["value"] == d["value"]
    assert cb["cores"] == len(os.sched_getaffinity(0))   # every core the process may run on, no cap
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_reference_arm_other_ranks_exit_quietly():
    p = run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2", "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": "29599"})
    assert p.returncode == 0
    assert not [ln for ln in p.stdout.splitlines() if ln.startswith("{")]


import pytest  # noqa: E402


@pytest.mark.gpu
def test_own_arm_prints_one_contract_line():
    """bench.py on the GPU (short run): every key the contract names, the launch count it claims,
    the byte counts of the end-to-end path as the library counted them."""
    p = subprocess.run([sys.executable, "bench.py", "--steps", "5", "--warmup", "3"], cwd=ROOT, capture_output=True,
                       text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert "impl" not in d and d["metric"] == "graph_eval_mpixel_per_s" and d["steps"] == 5 and d["warmup"] >= 3
    assert d["scaling"] == "weak" and d["vs_baseline"] is None and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert d["config"]["workload"].startswith("configs[1]") and "model" not in d["config"]
    assert d["gpu_launches"] == 5                                   # one fused kernel per step
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and 0.3 < r["frac"] <= 1.05 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] is None or r["traffic"] > 0
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == 6 * 4096 * 4096 * 4 and e["d2h_bytes_per_step"] == 4096 * 4096 * 4
    assert 0 < e["value"] < d["value"]
    if "unavailable" not in (e.get("u8_inputs") or {"unavailable": 1}):      # informational block: may opt out, never wrong
        assert e["u8_inputs"]["h2d_bytes_per_step"] == 2 * 4096 * 4096 * 4
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    # the other BASELINE configs ride in the same line, each with ms, bytes, a roofline fraction and a parity verdict
    w = d["workloads"]
    assert set(w) == {"height_to_normal_8192", "resize_1024_to_8192_rgba", "graphs32_batch64_4096", "mix_rgba_8192", "graph32_latency_256"}
    lat = w["graph32_latency_256"]
    assert lat["parity"]["ok"] is True and 0 < lat["replay_us_per_evaluation"] < lat["ordinary_us_per_evaluation"], lat
    for mode in ("fast", "exact"):
        e = w["height_to_normal_8192"][mode]
        assert e["ms"] > 0 and 0.2 < e["roofline"]["frac"] <= 1.05 and e["parity"]["ok"] is True, e
    assert w["height_to_normal_8192"]["exact"]["parity"]["bit_exact"] is True
    for filt in ("lanczos3", "gaussian"):
        e = w["resize_1024_to_8192_rgba"][filt]
        assert e["ms"] > 0 and 0.2 < e["roofline"]["frac"] <= 1.05 and e["parity"]["ok"] is True, e
    g = w["graphs32_batch64_4096"]
    assert g["graphs"] == 64 and g["kernels_per_graph"] < 32 and 0.2 < g["roofline"]["frac"] <= 1.05
    assert g["parity"]["ok"] is True and g["parity"]["samples_outside_1e-5rel_1e-6abs"] == 0, g["parity"]
    m = w["mix_rgba_8192"]
    assert m["pixels"] == 8192 * 8192 and 0.3 < m["roofline"]["frac"] <= 1.05 and m["parity"]["ok"] is True, m
