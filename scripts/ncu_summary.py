#!/usr/bin/env python
"""Turn an .ncu-rep (ncu --set full) into a small tracked summary under profiles/.

    python scripts/ncu_summary.py gpurun_out/prof_x.ncu-rep profiles/prof_x_r01 [--alg-bytes N]

Writes <out>.json (selected raw metrics per profiled launch) and <out>.txt (the same,
readable, plus the hottest source lines when the report carries -lineinfo source).
Runs on the CPU container: `ncu -i` only reads the report.
"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum",
    "lts__t_sector_hit_rate.pct",
    "l1tex__t_bytes.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.max",
    "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__cycles_active.avg",
    "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic",
    "launch__shared_mem_per_block_static",
    "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps",
    "launch__waves_per_multiprocessor",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    # drop ==PROF== style lines
    lines = [ln for ln in out.splitlines() if not ln.startswith("==")]
    return list(csv.reader(io.StringIO("\n".join(lines))))


def to_num(v):
    try:
        return float(v.replace(",", ""))
    except Exception:
        return v


def main():
    rep, out = sys.argv[1], sys.argv[2]
    alg = None
    if "--alg-bytes" in sys.argv:
        alg = float(sys.argv[sys.argv.index("--alg-bytes") + 1])
    rows = ncu_csv(rep, "raw")
    hdr, units = rows[0], rows[1]
    launches = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        e = {"kernel": d.get("Kernel Name"), "grid": d.get("Grid Size"), "block": d.get("Block Size"), "metrics": {}}
        for k in KEYS:
            if k in d and d[k] != "":
                e["metrics"][k] = {"value": to_num(d[k]), "unit": u[k]}
        m = e["metrics"]

        def scaled(key):
            if key not in m:
                return None
            v, un = m[key]["value"], m[key]["unit"].lower()
            mult = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1,
                    "usecond": 1e-6, "nsecond": 1e-9, "msecond": 1e-3, "second": 1}.get(un, 1)
            return v * mult
        t = scaled("gpu__time_duration.sum")
        rd, wr = scaled("dram__bytes_read.sum"), scaled("dram__bytes_write.sum")
        if rd is not None and wr is not None:
            e["dram_bytes_per_launch"] = rd + wr
            if t:
                e["dram_GBs_under_ncu"] = (rd + wr) / t / 1e9
        if alg and t:
            e["algorithmic_bytes"] = alg
            e["algorithmic_GBs_under_ncu"] = alg / t / 1e9
        launches.append(e)
    json.dump({"report": rep.split("/")[-1], "launches": launches}, open(out + ".json", "w"), indent=1)

    with open(out + ".txt", "w") as f:
        f.write("ncu --set full --clock-control none summary of %s\n" % rep.split("/")[-1])
        f.write("(times under ncu are cold-cache, serialised replays: use them for shares and traffic, not for throughput)\n\n")
        for e in launches:
            f.write("%s  grid=%s block=%s\n" % (e["kernel"], e["grid"], e["block"]))
            for k, v in e["metrics"].items():
                f.write("  %-84s %s %s\n" % (k, v["value"], v["unit"]))
            for k in ("dram_bytes_per_launch", "dram_GBs_under_ncu", "algorithmic_bytes", "algorithmic_GBs_under_ncu"):
                if k in e:
                    f.write("  %-84s %.6g\n" % (k, e[k]))
            f.write("\n")
        # hottest source lines
        try:
            src = ncu_csv(rep, "source")
            hi = next((i for i, r in enumerate(src[:4]) if "Source" in r), None)
            if hi is not None:
                h = src[hi]
                src = src[hi:]
                col = None
                for cand in ("Warp Stall Sampling (All Samples)", "Warp Stall Sampling (All Cycles)", "# Samples", "Sampling Data (All)"):
                    if cand in h:
                        col = h.index(cand)
                        break
                scol = h.index("Source") if "Source" in h else None
                if col is not None and scol is not None:
                    body = []
                    for r in src[1:]:
                        if len(r) <= max(col, scol):
                            continue
                        v = to_num(r[col])
                        if isinstance(v, float) and v > 0:
                            body.append((v, r[scol].strip()))
                    tot = sum(v for v, _ in body) or 1
                    body.sort(reverse=True)
                    f.write("hottest lines by warp-stall samples (%s):\n" % h[col])
                    for v, s in body[:25]:
                        f.write("  %6.2f%%  %s\n" % (100 * v / tot, s[:150]))
        except Exception as ex:  # source page is optional
            f.write("(no source page: %s)\n" % ex)
    print("wrote", out + ".json", out + ".txt")


if __name__ == "__main__":
    main()
