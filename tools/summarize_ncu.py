#!/usr/bin/env python
"""Summarise ncu outputs into small text files under profiles/ (run in the build container).
  launch list : python tools/summarize_ncu.py launches gpurun_out/launches.csv > profiles/...
  full capture: python tools/summarize_ncu.py full gpurun_out/prof.ncu-rep > profiles/..."""
import collections
import csv
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hi]
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        v = float(r[mv].replace(",", ""))
        v = v / 1e3 if r[mu] == "ns" else (v * 1e3 if r[mu] == "ms" else v)
        name = re.sub(r"^void ", "", re.sub(r"\(.*", "", r[kn]))
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none : {sum(cnt.values())} launches, "
          f"{T / 1e3:.2f} ms total (cold-cache, serialised: compare SHARES)")
    print(f"{'us':>12} {'share':>7} {'n':>5}  kernel")
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        print(f"{v:12.1f} {100 * v / T:6.1f}% {cnt[k]:5d}  {k[:110]}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("kernel:", r[hdr.index("Kernel Name")][:150])
        for k in KEYS:
            if k in hdr:
                print(f"  {k:90s} {r[hdr.index(k)]:>14s} {units[hdr.index(k)]}")
        print()


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
