#!/usr/bin/env python
"""bench.py -- graph-evaluation throughput of the B200 backend (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--math fast|exact] [--impl ours|reference]

A step = one evaluation of BASELINE.json configs[1]: Pow(Multiply(A, B), B) on two
synthetic 4096x4096 RGBA f32 images -> OutputRgba, i.e. one pass of the per-pixel hot
path (process_node + Mix kernels, src/node/node_type.rs:213, src/node/mix.rs:51) over one
batch of input.  `value` is measured with the inputs resident in HBM; `e2e` goes through
the public API from pinned host buffers to RGBA8 bytes on the host every step.

N > 1: one process per GPU (torchrun); every rank evaluates its own graph on its own
inputs (independent texture graphs, SURVEY.md section 8e: no data-path collective); the
only communication is the barrier and the max-over-ranks of the device time.

The same JSON line carries a `workloads` block with the other BASELINE.json configurations
(bench_workloads.py): HeightToNormal 8192^2 (FAST + EXACT; strips with a peer-memory halo at
N > 1), Resize Lanczos3/Gaussian 1024^2 -> 8192^2 RGBA (row strips at N > 1), the batch of 64
32-node graphs at 4096^2 split over the ranks, and one 8192^2 RGBA Mix graph -- each with its
ms, algorithmic bytes, roofline fraction and a parity sample against the CPU oracle.

--impl reference times the CPU restatement of the reference engine (oracle/, the Rust
crate cannot be built here) on every host core the process may use, --steps/--warmup as given.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SIZE = 4096
MPIX = SIZE * SIZE / 1e6
METRIC = "graph_eval_mpixel_per_s"
UNIT = "Mpixel/s"
WORKLOAD = "configs[1]: Pow(Multiply(A,B),B), two synthetic %dx%d RGBA f32 images -> OutputRgba" % (SIZE, SIZE)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def make_inputs(seed):
    r = np.random.default_rng(seed)
    return [r.random((SIZE, SIZE), dtype=np.float32) for _ in range(4)]


# ---------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle's engine on the host cores
# ---------------------------------------------------------------------------
def oracle_graph(A, B):
    import oracle
    g = oracle.Graph()
    g.embed(0, A)
    g.embed(1, B)
    g.add_node(1, 6, embed_id=0)          # Embed A
    g.add_node(2, 6, embed_id=1)          # Embed B
    g.add_node(3, 9, mix_type=2)          # Mix Multiply
    g.add_node(4, 9, mix_type=4)          # Mix Pow
    g.add_node(5, 3, name="out")          # OutputRgba
    g.add_edge(1, 3, 0, 0)
    g.add_edge(2, 3, 0, 1)
    g.add_edge(3, 4, 0, 0)
    g.add_edge(2, 4, 0, 1)
    g.add_edge(4, 5, 0, 0)
    return g


def cpu_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


CPU_BAND_ROWS = 1024     # the bounded sample: a 1024-row band (1/4) of the 4096^2 images per graph copy


def run_cpu(steps, warmup):
    """Copies of the graph evaluated concurrently, one thread per ready node (the
    reference engine's only parallelism, src/engine.rs:288, src/process_pack.rs:27);
    every node's pixel loop is single-threaded as in the reference.  Each step evaluates
    one copy per host thread on a CPU_BAND_ROWS-row band of the inputs (the graph is purely
    per-pixel, so a band costs exactly its share of the image): ~0.15 s per step, so that
    --steps 200 still ends within a minute, and 8 GB of planes at 32 threads."""
    rows = CPU_BAND_ROWS
    A = [p[:rows] for p in make_inputs(1)]
    B = [p[:rows] for p in make_inputs(2)]
    g = oracle_graph(A, B)
    threads = cpu_threads()              # every core the process may run on (sched_getaffinity)
    copies = threads
    band_mpix = rows * SIZE / 1e6
    for _ in range(warmup):
        g.eval_batch_seconds(copies, threads)
    total = 0.0
    for _ in range(steps):
        total += g.eval_batch_seconds(copies, threads)
    value = copies * steps * band_mpix / total
    return value, total / steps * 1e3, threads, ("%d steps x %d concurrent copies of the graph on a %d-row band (%d/%d) of the %dx%d images, one thread per copy, all %d cores of sched_getaffinity"
                                                 % (steps, copies, rows, rows, SIZE, SIZE, SIZE, threads))


def reference_arm(args, rank):
    if rank != 0:
        return
    steps, warm = max(1, args.steps), max(0, args.warmup)     # the driver's --steps / --warmup, as given
    value, ms, threads, sample = run_cpu(steps, warm)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "engine": "CPU restatement of the reference engine (oracle/): Rust toolchain absent, crate not buildable here"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.path = "/tmp/kc_clocks_%d_%d.csv" % (os.getpid(), device)
        self.proc = None
        try:
            self.f = open(self.path, "w", buffering=1)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def has_samples(self):
        try:
            return os.path.getsize(self.path) > 0
        except OSError:
            return False

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for ln in open(self.path):
                p = [x.strip() for x in ln.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1])); mx.append(float(p[2]))
                except ValueError:
                    continue
                for i, nm in enumerate(names):
                    if p[5 + i].lower().startswith("active"):
                        reasons.add(nm)
            os.remove(self.path)
        except Exception:
            pass
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["sm_max_mhz"] = float(max(mx))
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


# ---------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------
def ours(args, rank, world, local_rank):
    import ctypes as C

    import kanter_core_b200 as kc
    from kanter_core_b200 import MixType, Node, NodeType, SlotId
    from kanter_core_b200._lib import call

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        import datetime
        # a rank that dies must not leave the others waiting for the default ten minutes
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(seconds=240))

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    math_mode = kc.MATH_FAST if args.math == "fast" else kc.MATH_EXACT
    # NUMA: this rank's host thread moves next to its GPU (where the cpuset allows) BEFORE any pinned buffer is
    # allocated or filled; the pinned buffers themselves are placed on the GPU's node by the library (kc_numa.cu)
    numa = {"device_node": None, "thread_node_before": None, "thread_node": None, "nodes": None, "thread_moved": False}
    try:
        dn, tn, nn, moved = C.c_int32(-1), C.c_int32(-1), C.c_int32(0), C.c_int32(0)
        call("kc_numa_info", local_rank, C.byref(dn), C.byref(tn), C.byref(nn))
        numa["thread_node_before"] = tn.value
        if not args.no_numa:
            call("kc_bind_thread_near_device", local_rank, C.byref(moved))
        call("kc_numa_info", local_rank, C.byref(dn), C.byref(tn), C.byref(nn))
        numa.update(device_node=dn.value, thread_node=tn.value, nodes=nn.value, thread_moved=bool(moved.value))
    except Exception as ex:  # noqa: BLE001 - placement is best effort
        numa["error"] = repr(ex)[:120]
    tp = kc.TextureProcessor.new(device=local_rank, math_mode=math_mode)
    ctx = tp._ctx._h

    # pinned host copies of the inputs (rank-specific seeds: independent graphs)
    hostA = [kc.pinned_empty((SIZE, SIZE)) for _ in range(4)]
    hostB = [kc.pinned_empty((SIZE, SIZE)) for _ in range(4)]
    for dst, src in zip(hostA, make_inputs(1 + 2 * rank)):
        dst[...] = src
    for dst, src in zip(hostB, make_inputs(2 + 2 * rank)):
        dst[...] = src
    host_out = kc.pinned_empty((SIZE, SIZE, 4), np.uint8)

    imgA = kc.SlotImage.from_planes(tp, hostA)
    imgB = kc.SlotImage.from_planes(tp, hostB)

    lg = tp.new_live_graph()
    lg.embed_slot_data_with_id(kc.SlotData.new(0, 0, imgA), 0)
    lg.embed_slot_data_with_id(kc.SlotData.new(0, 0, imgB), 1)
    a = lg.add_node(Node.new(NodeType.Embed(0)))
    b = lg.add_node(Node.new(NodeType.Embed(1)))
    mul = lg.add_node(Node.new(NodeType.Mix(MixType.Multiply)))
    pw = lg.add_node(Node.new(NodeType.Mix(MixType.Pow)))
    out = lg.add_node(Node.new(NodeType.OutputRgba("out")))
    lg.connect(a, mul, SlotId(0), SlotId(0))
    lg.connect(b, mul, SlotId(0), SlotId(1))
    lg.connect(mul, pw, SlotId(0), SlotId(0))
    lg.connect(b, pw, SlotId(0), SlotId(1))
    lg.connect(pw, out, SlotId(0), SlotId(0))

    def step_resident():
        # new inputs arrive (same planes, already in HBM): everything downstream is dirty again
        lg.replace_embedded(imgA, 0)
        lg.replace_embedded(imgB, 1)
        lg.request(out)

    # ---- device-resident throughput --------------------------------------------------------
    sampler = ClockSampler(local_rank)
    for _ in range(max(args.warmup, 3)):
        step_resident()
    # the tape became hot during these steps and its specialised kernel is being compiled in the
    # background (no stall in a caller's loop): the warm-up ends when that kernel is in use
    kc.jit_wait()
    step_resident()
    tp.synchronize()
    stats = lg.last_run_stats()
    # the timed region is milliseconds long: keep the GPU under the very same load for about
    # half a second around it while nvidia-smi samples (every 20 ms), so the clocks line
    # describes the device under this kernel stream and not an idle one
    t_wait = time.perf_counter()
    while time.perf_counter() - t_wait < (0.6 if sampler.has_samples() else 5.0):
        for _ in range(50):
            step_resident()
        tp.synchronize()
        if sampler.has_samples() and time.perf_counter() - t_wait > 0.6:
            break

    ev0, ev1 = C.c_void_p(), C.c_void_p()
    call("kc_event_create", C.byref(ev0))
    call("kc_event_create", C.byref(ev1))
    k0 = tp.stats()["kernel_launches"]
    call("kc_context_set_timing", ctx, 1)
    barrier()
    tp.synchronize()
    call("kc_event_record", ctx, ev0)
    for _ in range(args.steps):
        step_resident()
    call("kc_event_record", ctx, ev1)
    tp.synchronize()
    barrier()
    ms = C.c_float()
    call("kc_event_elapsed_ms", ev0, ev1, C.byref(ms))
    clocks = sampler.stop()   # samples cover the load loop that precedes the timed region and the region itself
    kms, kn = C.c_double(), C.c_uint64()
    call("kc_context_timing_read", ctx, 0, C.byref(kms), C.byref(kn))   # the fused tape kernel
    call("kc_context_set_timing", ctx, 0)
    launches = tp.stats()["kernel_launches"] - k0
    total_ms = max_over_ranks(float(ms.value))
    value = world * args.steps * MPIX / (total_ms / 1e3)

    # ---- parity spot check of what was just timed (rank 0, cheap sample) -------------------
    parity = None
    if rank == 0:
        import oracle
        got = lg.slot_data(out, SlotId(0)).image.planes()
        rows = slice(0, 64)
        for c in range(3):
            want = oracle.mix_plane(4, oracle.mix_plane(2, hostA[c][rows], hostB[c][rows]), hostB[c][rows]).astype(np.float64)
            g = got[c][rows].astype(np.float64)
            assert (np.abs(g - want) <= 1e-6 + 1e-5 * np.abs(want)).all(), "bench output differs from the oracle"
        assert np.array_equal(got[3][rows], np.ones((64, SIZE), np.float32))
        parity = "checked vs CPU oracle on 64 rows x 3 channels: within 1e-5 rel / 1e-6 abs" + (" (bit-exact mode)" if args.math == "exact" else "")

    # ---- end to end through the public API: pinned host f32 planes -> RGBA8 bytes on host ----
    # Every step uploads its 8 input planes from pinned host memory and reads its RGBA8 result
    # back to the host.  Steps are software-pipelined one deep, the way a caller streaming
    # textures would drive the API: step i is enqueued (read_rgba(sync=False)) before the host
    # waits for step i-1's bytes, so the upload of step i (upload stream) overlaps the download
    # of step i-1 (download stream).  Two result buffers alternate.
    host_outs = [host_out, kc.pinned_empty((SIZE, SIZE, 4), np.uint8)]
    done = [C.c_void_p(), C.c_void_p()]
    for e in done:
        call("kc_event_create", C.byref(e))

    def enqueue_e2e(i):
        # deferred: each plane is copied to the GPU (upload stream) when the evaluation first reads it.
        # Mix never reads its operands' alpha (src/node/mix.rs:194-302 writes A = 1.0), so 6 of the 8
        # input planes cross PCIe; h2d/d2h below are the bytes the library counted, not an estimate.
        ia = kc.SlotImage.from_planes(tp, hostA, sync=False, deferred=True)
        ib = kc.SlotImage.from_planes(tp, hostB, sync=False, deferred=True)
        lg.replace_embedded(ia, 0)
        lg.replace_embedded(ib, 1)
        lg.read_rgba(out, SlotId(0), kc.Size(SIZE, SIZE), out=host_outs[i % 2], sync=False)
        call("kc_event_record_download", ctx, done[i % 2])

    def run_e2e(n, stamps=None, overlap=True):
        """overlap: step i is enqueued before the host waits for step i-1 (its upload runs under step i-1's download: full
        duplex).  Not overlapped: step i-1's bytes are on the host before step i is enqueued (one direction at a time)."""
        for i in range(n):
            if not overlap and i > 0:
                call("kc_event_synchronize", done[(i - 1) % 2])
                if stamps is not None:
                    stamps.append(time.perf_counter())
            enqueue_e2e(i)
            if overlap and i > 0:
                call("kc_event_synchronize", done[(i - 1) % 2])      # step i-1's bytes are on the host
                if stamps is not None:
                    stamps.append(time.perf_counter())
        call("kc_event_synchronize", done[(n - 1) % 2])
        if stamps is not None:
            stamps.append(time.perf_counter())

    e2e_steps = max(1, min(args.steps, 20))
    run_e2e(3)
    kc.jit_wait()        # the fused RGBA8-export tape of this path, same as above
    run_e2e(2)
    # Which schedule does THIS box favour?  One GPU's PCIe link is full duplex, and overlapping wins; eight ranks share one host
    # fabric that moves less in total when both directions are busy (the `pcie` probe below: 241 GB/s host->device alone,
    # 163 GB/s with both directions at 8 ranks), and there one direction at a time wins.  Measured, not assumed: four steps
    # of each, max over ranks, the faster one is used for the timed region (every rank takes the same decision).
    trial = {}
    for name, ov in (("overlapped", True), ("one_direction_at_a_time", False)):
        barrier()
        tp.synchronize()
        tt = time.perf_counter()
        run_e2e(4, overlap=ov)
        tp.synchronize()
        trial[name] = max_over_ranks(time.perf_counter() - tt) / 4 * 1e3
    e2e_overlap = trial["overlapped"] <= trial["one_direction_at_a_time"]
    barrier()
    tp.synchronize()
    stamps = []
    x0 = tp.transfer_stats()
    t0 = time.perf_counter()
    run_e2e(e2e_steps, stamps, overlap=e2e_overlap)
    tp.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    x1 = tp.transfer_stats()
    e2e_value = world * e2e_steps * MPIX / e2e_s
    e2e_step_ms = [round((b - a) * 1e3, 3) for a, b in zip([t0] + stamps[:-1], stamps)]
    # the result of the last pipelined step against the resident run's planes (same inputs)
    if rank == 0:
        last = host_outs[(e2e_steps - 1) % 2]
        ref8 = lg.buffer_rgba(out, SlotId(0))
        assert np.array_equal(last, ref8), "pipelined e2e result differs from the synchronous export"
    h2d = (x1["h2d_bytes"] - x0["h2d_bytes"]) // e2e_steps
    d2h = (x1["d2h_bytes"] - x0["d2h_bytes"]) // e2e_steps

    # ---- the same end-to-end path with 8-bit inputs: what an Image node hands over (src/shared.rs:16-56) ----
    # Informational (the headline e2e above is the f32 contract): two RGBA8 images up (128 MiB instead of 384 MiB),
    # u8 -> f32 planes on the device, the same fused kernel, RGBA8 down.
    e2e_u8, u8_s, u8_fail, y0, y1 = None, 0.0, 0.0, None, None
    try:   # no collective inside: a rank that fails here must not leave the others waiting at one
        r8 = np.random.default_rng(1234 + rank)
        u8A, u8B = kc.pinned_empty((SIZE, SIZE, 4), np.uint8), kc.pinned_empty((SIZE, SIZE, 4), np.uint8)
        u8A[...] = r8.integers(0, 256, size=u8A.shape, dtype=np.uint8)
        u8B[...] = r8.integers(0, 256, size=u8B.shape, dtype=np.uint8)

        def run_u8(n):
            for i in range(n):
                lg.replace_embedded(kc.SlotImage.from_u8(tp, u8A, sync=False), 0)
                lg.replace_embedded(kc.SlotImage.from_u8(tp, u8B, sync=False), 1)
                lg.read_rgba(out, SlotId(0), kc.Size(SIZE, SIZE), out=host_outs[i % 2], sync=False)
                call("kc_event_record_download", ctx, done[i % 2])
                if i > 0:
                    call("kc_event_synchronize", done[(i - 1) % 2])
            call("kc_event_synchronize", done[(n - 1) % 2])

        run_u8(3)
        kc.jit_wait()
        run_u8(2)
        tp.synchronize()
        y0 = tp.transfer_stats()
        t0 = time.perf_counter()
        run_u8(e2e_steps)
        tp.synchronize()
        u8_s = time.perf_counter() - t0
        y1 = tp.transfer_stats()
        if rank == 0:
            a32, b32 = u8A[:8].astype(np.float32) / np.float32(255.0), u8B[:8].astype(np.float32) / np.float32(255.0)
            last = host_outs[(e2e_steps - 1) % 2][:8].astype(np.int32)
            want = np.power((a32 * b32).astype(np.float32), b32).astype(np.float32)
            want8 = np.minimum(np.clip(want, 0, 1) * np.float32(255.0), 255).astype(np.int32)
            want8[..., 3] = 255
            assert np.abs(last - want8).max() <= 1, "u8 end-to-end result differs from numpy"
        kc.free_pinned(u8A)
        kc.free_pinned(u8B)
    except Exception as ex:  # noqa: BLE001 - informational block: never costs the line
        u8_fail, e2e_u8 = 1.0, {"unavailable": repr(ex)[:200]}
    u8_fail = max_over_ranks(u8_fail)
    u8_s = max_over_ranks(u8_s)
    if u8_fail == 0.0:
        e2e_u8 = {"value": world * e2e_steps * MPIX / u8_s, "unit": UNIT,
                  "h2d_bytes_per_step": (y1["h2d_bytes"] - y0["h2d_bytes"]) // e2e_steps,
                  "d2h_bytes_per_step": (y1["d2h_bytes"] - y0["d2h_bytes"]) // e2e_steps, "steps": e2e_steps,
                  "path": "2 pinned RGBA8 images per step -> u8/255 planes on the device -> fused mul/pow/to_u8 kernel -> RGBA8 on pinned host; pipelined one deep"}
    elif e2e_u8 is None:
        e2e_u8 = {"unavailable": "another rank failed"}

    # ---- the PCIe ceiling of this box, all ranks at once: plain cudaMemcpyAsync out of / into the same kind of pinned memory ----
    pcie = None
    try:
        probe = kc.pinned_empty((64 << 20,), np.uint8)                    # 64 MiB: one plane
        probe[...] = 1
        res = {}
        for name, direction in (("h2d", 0), ("d2h", 1), ("both", 2)):
            barrier()
            a, b = C.c_double(), C.c_double()
            call("kc_context_pcie_probe", ctx, probe.ctypes.data, probe.nbytes, 12, direction, C.byref(a), C.byref(b))
            res[name] = (a.value, b.value)
        kc.free_pinned(probe)
        mine = [res["h2d"][0], res["d2h"][1], res["both"][0], res["both"][1]]
    except Exception as ex:  # noqa: BLE001
        mine = [0.0, 0.0, 0.0, 0.0]
        numa["probe_error"] = repr(ex)[:160]
    if dist is not None:
        allv = [None] * world
        dist.all_gather_object(allv, (mine, numa))
    else:
        allv = [(mine, numa)]
    if rank == 0:
        per = [v[0] for v in allv]
        pcie = {"probe": "12 x 64 MiB plain cudaMemcpyAsync per rank from/to NUMA-placed pinned memory, every rank at once (barrier), CUDA events",
                "per_rank_h2d_GBs": [round(p[0], 1) for p in per], "per_rank_d2h_GBs": [round(p[1], 1) for p in per],
                "aggregate_h2d_GBs": round(sum(p[0] for p in per), 1), "aggregate_d2h_GBs": round(sum(p[1] for p in per), 1),
                "aggregate_duplex_GBs": round(sum(p[2] + p[3] for p in per), 1),
                "e2e_bytes_per_step_per_rank": int(h2d + d2h),
                "e2e_aggregate_GBs": round(world * (h2d + d2h) * e2e_steps / e2e_s / 1e9, 1),
                "numa": [v[1] for v in allv]}
        # the step is host->device heavy (6 f32 planes up, one RGBA8 image down): its floor is the upload alone at the probe's rate
        if pcie["aggregate_h2d_GBs"] > 0:
            pcie["e2e_h2d_GBs"] = round(world * h2d * e2e_steps / e2e_s / 1e9, 1)
            pcie["e2e_fraction_of_h2d_ceiling"] = round(pcie["e2e_h2d_GBs"] / pcie["aggregate_h2d_GBs"], 3)

    peak, peak_src = peaks()
    workloads = None
    if not args.no_workloads:
        import bench_workloads as bw

        def gather(obj):
            outl = [None] * world
            dist.all_gather_object(outl, obj)
            return outl
        env = bw.Env(kc, tp, rank, world, barrier, max_over_ranks, peak, gather if dist is not None else None)
        try:
            workloads = bw.run_all(env, max(3, min(args.steps, 10)))
        except Exception as ex:  # noqa: BLE001 - the headline above is measured already: it must still be printed
            workloads = {"error": repr(ex)[:300]}
        tp.set_math_mode(math_mode)
    alg_bytes = stats["algorithmic_bytes"] / max(1, stats["kernels"]) if stats["kernels"] else 0
    avg_kernel_ms = kms.value / max(1, kn.value)
    achieved = alg_bytes / (avg_kernel_ms / 1e3) / 1e9 if avg_kernel_ms > 0 else 0.0
    # DRAM bytes per launch of the dominant kernel: from the newest ncu --set full capture of this same command
    # (profiles/traffic_rNN.json names the capture file it was read from); a profiler cannot run inside the bench
    traffic, traffic_src = None, None
    import glob
    for tpath in sorted(glob.glob(os.path.join(ROOT, "profiles", "traffic_r*.json")), reverse=True):
        try:
            tj = json.load(open(tpath))
            ent = tj.get("fused_kernel_%s_specialised" % args.math) or tj.get("tape_kernel_%s" % args.math, {})
            if ent.get("dram_bytes_per_launch"):
                traffic = ent["dram_bytes_per_launch"]
                traffic_src = "%s (%s)" % (os.path.relpath(tpath, ROOT), ent.get("capture", tj.get("capture", "ncu --set full")))
                break
        except Exception:
            continue

    tv, tc_, ts_ = C.c_int32(), C.c_int32(), C.c_int32()
    call("kc_debug_last_tile_config", C.byref(tv), C.byref(tc_), C.byref(ts_))
    mode_name = "EXACT" if args.math == "exact" else "FAST"
    kernel_name = ("kc_jit_entry = kc_tile_vm_body<%s> specialised for this tape with NVRTC" % mode_name) if tv.value < 0 else ("kc_tile_vm_kernel<%s> (tape interpreter)" % mode_name)
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "graphs_per_step_per_gpu": 1, "math_mode": args.math,
                       "parallelism": "independent graphs per GPU, no collective" if world > 1 else "single GPU",
                       "l2": "inputs (512 MiB) + outputs (256 MiB) per step exceed the 126 MB L2; no flush needed",
                       "parity": parity},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "step_ms_min_median_max": [min(e2e_step_ms), float(np.median(e2e_step_ms)), max(e2e_step_ms)],
                    "schedule": "overlapped (step i enqueued before step i-1's bytes are awaited)" if e2e_overlap else "one direction at a time (step i-1's bytes on the host before step i is enqueued)",
                    "schedule_trial_ms_per_step": {k: round(v, 3) for k, v in trial.items()}, "u8_inputs": e2e_u8, "path": "8 pinned host f32 planes per step -> deferred upload of the 6 planes the graph reads (upload stream) -> fused mul/pow/to_u8 kernel -> RGBA8 on pinned host (read_rgba, download stream); steps pipelined one deep unless the trial says otherwise (`schedule`); bytes as counted by the library"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "kernel": kernel_name,
                         "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": avg_kernel_ms,
                         "launches_timed": int(kn.value), "peak_source": peak_src,
                         "frac_of_nominal_8TBs": achieved / 8000.0},
        }
        if pcie is not None:
            line["pcie"] = pcie
        if workloads is not None:
            line["workloads"] = workloads
        if world == 1 and not args.no_cpu:
            v, _ms, threads, sample = run_cpu(3, 1)      # one warm-up (first touch of 8 GB of planes), three timed steps
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample}
        print(json.dumps(line), flush=True)

    for arr in hostA + hostB + host_outs:
        kc.free_pinned(arr)
    tp.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--math", default="fast", choices=["fast", "exact"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-numa", action="store_true", help="leave the host thread where the launcher put it (A/B of the NUMA placement; KC_NO_NUMA=1 also disables the memory placement)")
    ap.add_argument("--no-workloads", action="store_true", help="headline only: skip the `workloads` block (configs[2..4], 8192^2)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank)
        return
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under it, as the driver would
        port = 29500 + os.getpid() % 1000
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
