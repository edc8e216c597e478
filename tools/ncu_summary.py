#!/usr/bin/env python
"""Turn ncu artefacts from gpurun_out/ into the small text summaries committed under profiles/.

    python tools/ncu_summary.py launches gpurun_out/x_launches.csv  > profiles/x_launches.txt
    python tools/ncu_summary.py full     gpurun_out/x.ncu-rep       > profiles/x_full.txt
"""
import collections
import csv
import re
import subprocess
import sys

KEEP = re.compile(
    r"^(gpu__time_duration\.sum|dram__bytes_(read|write)\.sum|dram__throughput\.avg\.pct_of_peak_sustained_elapsed|"
    r"lts__t_bytes\.sum|lts__throughput\.avg\.pct_of_peak_sustained_elapsed|"
    r"launch__(registers_per_thread|grid_size|block_size|waves_per_multiprocessor|occupancy_limit_\w+)|"
    r"sm__warps_active\.avg\.pct_of_peak_sustained_active|sm__cycles_(active\.avg|elapsed\.max)|"
    r"smsp__inst_executed\.sum|smsp__issue_active\.avg\.pct_of_peak_sustained_active|"
    r"sm__pipe_fp64_cycles_active\.avg\.pct_of_peak_sustained_(active|elapsed)|"
    r"sm__inst_executed_pipe_(alu|fma|fp64|lsu)\.avg\.pct_of_peak_sustained_active|"
    r"sm__throughput\.avg\.pct_of_peak_sustained_elapsed|"
    r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum|"
    r"smsp__average_warps_issue_stalled_\w+_per_issue_active\.ratio)$")


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
        a = agg.setdefault(r[ki][:70], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"# {path}: per-kernel device time (ncu gpu__time_duration.sum, --clock-control none; cold-cache, serialised)")
    print(f"{'kernel':70s} {'launches':>8s} {'total ms':>10s} {'avg us':>10s} {'share':>6s}")
    for k, a in agg.items():
        print(f"{k:70s} {a[0]:8d} {a[1] / 1e3:10.3f} {a[1] / a[0]:10.1f} {a[1] / tot:6.3f}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# {path}: selected metrics of `ncu --set full --clock-control none` (one block per profiled launch)")
    for r in rows[2:]:
        print(f"## {r[hdr.index('Kernel Name')]}  grid={r[hdr.index('Grid Size')]} block={r[hdr.index('Block Size')]}")
        for h, u, v in zip(hdr, units, r):
            if KEEP.match(h):
                print(f"{h:85s} {v:>18s} {u}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
