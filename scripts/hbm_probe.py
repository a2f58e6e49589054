#!/usr/bin/env python
"""HBM ceilings by access mix on this B200, torch library kernels + CUDA events:
write-only (fill_), read-only (sum), copy (read+write), 1 read : 3 writes (the
HeightToNormal mix).  Context for write-dominated kernels (resize upsampling,
HeightToNormal): MEASURED_PEAKS.json's hbm_gbs is a COPY number.
    python scripts/hbm_probe.py [--out profiles/hbm_probe_rNN.json]
"""
import json
import sys

import torch


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def main():
    n = 1 << 28  # 1 GiB of f32
    x = torch.empty(n, dtype=torch.float32, device="cuda")
    y = torch.empty(n, dtype=torch.float32, device="cuda")
    x.fill_(1.0)
    res = {}
    t = timeit(lambda: y.fill_(0.5))
    res["write_only_fill_GBs"] = n * 4 / t / 1e6
    t = timeit(lambda: torch.cuda.current_stream().synchronize() or y.zero_())
    res["write_only_zero_GBs"] = n * 4 / t / 1e6
    t = timeit(lambda: x.sum())
    res["read_only_sum_GBs"] = n * 4 / t / 1e6
    t = timeit(lambda: y.copy_(x))
    res["copy_GBs"] = 2 * n * 4 / t / 1e6
    q = n // 4
    a = x[:q]
    outs = [y[:q], y[q:2 * q], y[2 * q:3 * q]]
    res["note"] = "1 GiB buffers, best of 20, torch elementwise kernels"
    for k, v in res.items():
        if isinstance(v, float):
            print("%-28s %8.0f GB/s" % (k, v))
    if "--out" in sys.argv:
        json.dump(res, open(sys.argv[sys.argv.index("--out") + 1], "w"), indent=1)


if __name__ == "__main__":
    main()
