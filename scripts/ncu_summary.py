#!/usr/bin/env python
"""Summarise ncu output into the small text files kept under profiles/.

    python scripts/ncu_summary.py launches <launches.csv>          # per-kernel device time + share
    python scripts/ncu_summary.py full <report.ncu-rep>            # key metrics of each profiled launch
    python scripts/ncu_summary.py stalls <report.ncu-rep> <kernel regex> [min share]

Runs in the build container (ncu can read reports without a GPU).
"""
import collections
import csv
import subprocess
import sys

FULL_METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
]


def launches(path):
    """Per-kernel device time (+ DRAM bytes when the list was taken with dram__bytes_read/write.sum too)."""
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ki, vi, mi, ii = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name"), hdr.index("ID")
    per_launch = collections.OrderedDict()
    for r in rows[1:]:
        name = r[ki].split("(")[0].replace("void ", "")
        per_launch.setdefault((r[ii], name), {})[r[mi]] = float(r[vi].replace(",", ""))
    tot = collections.OrderedDict()
    for (_, name), m in per_launch.items():
        t = tot.setdefault(name, [0, 0.0, 0.0, 0.0])
        t[0] += 1
        t[1] += m.get("gpu__time_duration.sum", 0.0)
        t[2] += m.get("dram__bytes_read.sum", 0.0)
        t[3] += m.get("dram__bytes_write.sum", 0.0)
    total = sum(v[1] for v in tot.values())
    print(f"{'kernel':60s} {'launches':>8s} {'avg us':>10s} {'share':>7s} {'DRAM rd MB':>11s} {'DRAM wr MB':>11s}")
    for name, (n, ns, rd, wr) in tot.items():
        print(f"{name[:60]:60s} {n:8d} {ns / n / 1e3:10.2f} {100 * ns / total:6.1f}% {rd / n / 1e6:11.1f} {wr / n / 1e6:11.1f}")


def full(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("----", r[hdr.index("Kernel Name")])
        for m in FULL_METRICS:
            if m in hdr:
                print(f"  {m:66s} {r[hdr.index(m)]} {units[hdr.index(m)]}")


def stalls(rep, regex, thr=0.02):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{regex}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        print(f"no kernel matching {regex!r} in {rep}")
        return
    hdr = rows[1]
    data = [r for r in rows[2:] if len(r) == len(hdr) and r[0] != "Address"]

    def num(x):
        try:
            return int(x)
        except ValueError:
            return 0

    si, src, ie = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
    names = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(num(r[si]) for r in data)
    print(f"{rows[0][1] if len(rows[0]) > 1 else regex}")
    print(f"samples {tot}, SASS instructions {len(data)}, executed warp instructions {sum(num(r[ie]) for r in data)}")
    for k, v in sorted(((s, sum(num(r[hdr.index(s)]) for r in data)) for s in names), key=lambda x: -x[1])[:8]:
        print(f"  {k:28s} {100 * v / max(tot, 1):5.1f}%")
    print("hottest instructions (share of samples):")
    for n, r in enumerate(data):
        if num(r[si]) > tot * thr:
            print(f"  #{n:5d} {100 * num(r[si]) / tot:5.1f}%  {r[src].strip()[:90]}")


if __name__ == "__main__":
    cmd = sys.argv[1]
    if cmd == "launches":
        launches(sys.argv[2])
    elif cmd == "full":
        full(sys.argv[2])
    else:
        stalls(sys.argv[2], sys.argv[3], float(sys.argv[4]) if len(sys.argv) > 4 else 0.02)
