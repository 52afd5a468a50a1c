#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small text files for profiles/.
  python tools/ncu_summary.py launches <launches.csv>           per-kernel count / total us / share of the step
  python tools/ncu_summary.py full <file.ncu-rep> [kernel-substring]   DRAM bytes, duration, throughput per launch
"""
import collections
import csv
import subprocess
import sys


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg, tot = collections.defaultdict(lambda: [0, 0.0]), 0.0
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(row["Metric Unit"], 1.0)
        k = row["Kernel Name"].split("(")[0].replace("void ", "") + " grid" + row["Grid Size"].replace(" ", "")
        agg[k][0] += 1
        agg[k][1] += v
        tot += v
    print(f"# {path}: gpu__time_duration.sum per launch (cold cache, serialised: compare shares)")
    print(f"{'kernel':70s} {'count':>6} {'total us':>11} {'avg us':>9} {'share':>6}")
    for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:40]:
        print(f"{k[:70]:70s} {c:6d} {t:11.1f} {t / c:9.2f} {100 * t / tot:5.1f}%")
    print(f"{'total':70s} {sum(c for c, _ in agg.values()):6d} {tot:11.1f}")


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"]


def full(path, sub=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units, data = rows[0], rows[1], rows[2:]
    kn = h.index("Kernel Name")
    print(f"# {path}: ncu --set full, per launch")
    for r in data:
        if sub and sub not in r[kn]:
            continue
        print("kernel:", r[kn][:110])
        for w in WANT:
            if w in h:
                i = h.index(w)
                print(f"  {w:75s} {r[i]:>16} {units[i]}")
        try:
            rd = float(r[h.index("dram__bytes_read.sum")].replace(",", ""))
            wr = float(r[h.index("dram__bytes_write.sum")].replace(",", ""))
            mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            tr = rd * mul[units[h.index("dram__bytes_read.sum")]] + wr * mul[units[h.index("dram__bytes_write.sum")]]
            print(f"  {'traffic = dram read + write':75s} {tr:16.0f} byte")
        except Exception:
            pass


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
