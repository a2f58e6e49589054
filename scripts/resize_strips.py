#!/usr/bin/env python
"""BASELINE.json configs[3]: Resize Lanczos3 / Gaussian 1024^2 -> 8192^2 RGBA, tiled over the GPUs
of one box by OUTPUT ROWS: every rank holds the whole 16 MiB source and
computes rows [y0, y1) of the 1 GiB result with kc_resize_rows.  No inter-GPU traffic at all.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29512 scripts/resize_strips.py [--src 1024] [--dst 8192] [--steps 20] [--math fast|exact]
    (or plainly `python scripts/resize_strips.py` for one GPU)

Rank 0 prints one JSON line per filter; `value` is the whole image's Mpixel/s (strong scaling).
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", type=int, default=1024)
    ap.add_argument("--dst", type=int, default=8192)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--math", default="fast", choices=["fast", "exact"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    import kanter_core_b200 as kc
    from kanter_core_b200 import ResizeFilter
    from kanter_core_b200 import dist as kdist
    from kanter_core_b200._lib import call, kc_image

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    tp = kc.TextureProcessor.new(device=local, math_mode=kc.MATH_FAST if args.math == "fast" else kc.MATH_EXACT)
    ctx = tp._ctx._h
    S, D = args.src, args.dst
    r = np.random.default_rng(4)
    planes = [r.random((S, S), dtype=np.float32) for _ in range(4)]       # the same on every rank (seeded)
    img = kc.SlotImage.from_planes(tp, planes)
    y0, y1 = kdist.strip_rows(D, rank, world)

    def step(filt):
        out = kc_image()
        call("kc_resize_rows", ctx, C.byref(img._im), D, D, int(filt), y0, y1 - y0, C.byref(out))
        return kc.SlotImage(tp._ctx, out)

    for filt in (ResizeFilter.Lanczos3, ResizeFilter.Gaussian):
        res = None
        for _ in range(3):
            res = step(filt)
        tp.synchronize()
        parity = None
        if rank == 0:
            import oracle
            rows = min(16, y1 - y0)
            want = oracle.resize_plane(planes[0], D, D, int(filt))[y0:y0 + rows]      # a few seconds of CPU at 8192^2
            got = res.planes()[0][:rows]
            if args.math == "exact":
                assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), "strip differs from the oracle"
                parity = "first %d rows of plane 0 bit-identical to the CPU oracle" % rows
            else:
                assert (np.abs(got.astype(np.float64) - want) <= 1e-6 + 1e-5 * np.abs(want)).all()
                parity = "first %d rows of plane 0 within 1e-5 rel / 1e-6 abs of the CPU oracle" % rows
        ev0, ev1 = C.c_void_p(), C.c_void_p()
        call("kc_event_create", C.byref(ev0))
        call("kc_event_create", C.byref(ev1))
        if dist is not None:
            dist.barrier()
        tp.synchronize()
        call("kc_event_record", ctx, ev0)
        for _ in range(args.steps):
            res = step(filt)
        call("kc_event_record", ctx, ev1)
        tp.synchronize()
        ms = C.c_float()
        call("kc_event_elapsed_ms", ev0, ev1, C.byref(ms))
        t = float(ms.value) / args.steps
        if dist is not None:
            import torch
            tt = torch.tensor([t], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t = float(tt.item())
        if rank == 0:
            alg = 4 * (S * S * 4 + D * D * 4)           # whole job: source read once per GPU is 16 MiB, negligible
            print(json.dumps({"workload": "configs[3]: Resize %s %dx%d -> %dx%d RGBA, %d row strip(s), source replicated" % (filt.name, S, S, D, D, world),
                              "metric": "graph_eval_mpixel_per_s", "unit": "Mpixel/s", "n_gpus": world, "scaling": "strong",
                              "value": D * D / 1e6 / (t / 1e3), "ms_per_step_max_over_ranks": t,
                              "algorithmic_GBs_whole_job": alg / (t / 1e3) / 1e9, "math_mode": args.math, "parity": parity}), flush=True)
    tp.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
