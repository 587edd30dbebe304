#!/usr/bin/env python
"""profiles/summarize.py — turns ncu outputs brought back in gpurun_out/ into the small text
summaries committed under profiles/.

  python profiles/summarize.py launches gpurun_out/launches_X.csv          # per-kernel launch table
  python profiles/summarize.py launches gpurun_out/launches_X.csv 'k_tile_mover<0>' 2   # only after the 2nd mover launch
  python profiles/summarize.py full gpurun_out/prof_X.ncu-rep [regex]      # key counters of an ncu --set full capture
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "launch__shared_mem_per_block_static", "smsp__inst_executed.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_atom.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum",
    "smsp__inst_executed_op_shared_atom.sum", "smsp__inst_executed_op_global_red.sum",
    "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum",
    "sm__sass_inst_executed_op_shared_atom.sum", "sm__sass_inst_executed_op_global_red.sum",
    "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_wait_per_warp_active.pct", "smsp__warp_issue_stalled_not_selected_per_warp_active.pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
]


def launches(path, after_kernel=None, after_count=0):
    """Per-kernel totals; with after_kernel/after_count only the launches AFTER the after_count-th launch of
    that kernel are counted (used to cut the list down to bench.py's timed region)."""
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ki, vi, mi, ui = (hdr.index(k) for k in ("Kernel Name", "Metric Value", "Metric Name", "Metric Unit"))
    agg = collections.OrderedDict()
    seen = 0
    for r in rows[1:]:
        if r[mi] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r[ki]).strip()
        if after_kernel is not None and seen < after_count:
            if after_kernel in name:
                seen += 1
            continue
        v = float(r[vi].replace(",", ""))
        scale = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}[r[ui]]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v * scale
    tot = sum(a[1] for a in agg.values())
    print("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k} | {n} | {t:.1f} | {t / n:.1f} | {100 * t / tot:.2f}% |")


def full(path, regex=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        if regex and not re.search(regex, d["Kernel Name"]):
            continue
        print(f"### {d['Kernel Name'][:110]}  (launch id {d.get('ID', '?')})")
        for k in KEYS:
            if k in d and d[k] != "":
                print(f"  {k:75s} {d[k]:>18s} {units[hdr.index(k)]}")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        if len(sys.argv) > 4:
            launches(sys.argv[2], sys.argv[3], int(sys.argv[4]))
        else:
            launches(sys.argv[2])
    else:
        full(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
