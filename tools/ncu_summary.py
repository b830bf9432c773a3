#!/usr/bin/env python
"""Summarise ncu outputs into the small text files kept under profiles/ (the .ncu-rep itself stays in gpurun_out/).

  python tools/ncu_summary.py launches gpurun_out/launches_r1.csv          # per-kernel count / mean time / share of the step
  python tools/ncu_summary.py full gpurun_out/prof_r1.ncu-rep               # DRAM bytes, durations, occupancy per captured launch
"""
import csv
import subprocess
import sys
from collections import defaultdict


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    i_name, i_val, i_unit = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = defaultdict(list)
    for r in rows[1:]:
        v = float(r[i_val].replace(",", ""))
        v = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[i_unit], 1e-6) * v
        agg[r[i_name]].append(v)
    tot = sum(sum(v) for v in agg.values())
    print(f"# {path}: gpu__time_duration.sum per launch (ncu --clock-control none; serialised, cold-cache: compare SHARES)")
    print(f"{'kernel':100s} {'n':>4s} {'mean ms':>10s} {'share':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k[:100]:100s} {len(v):4d} {sum(v) / len(v):10.4f} {sum(v) / tot * 100:6.1f}%")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
            "launch__occupancy_limit_registers", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "lts__t_sector_hit_rate.pct"]
    idx = [hdr.index(w) for w in want if w in hdr]
    print(f"# {path}: ncu --set full --clock-control none")
    for r in rows[2:]:
        print("-" * 100)
        for i in idx:
            print(f"  {hdr[i]:70s} {r[i]:>24s} {units[i]}")
        try:
            rd = float(r[hdr.index('dram__bytes_read.sum')].replace(",", ""))
            wr = float(r[hdr.index('dram__bytes_write.sum')].replace(",", ""))
            ur, uw = units[hdr.index('dram__bytes_read.sum')], units[hdr.index('dram__bytes_write.sum')]
            sc = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            t = float(r[hdr.index('gpu__time_duration.sum')].replace(",", ""))
            ut = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}[units[hdr.index('gpu__time_duration.sum')]]
            traffic = rd * sc[ur] + wr * sc[uw]
            print(f"  {'=> DRAM traffic (read+write)':70s} {traffic / 1e9:24.4f} GB   => {traffic / (t * ut) / 1e9:.1f} GB/s under ncu")
        except Exception:
            pass


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
