#!/usr/bin/env python
"""Turns ncu exports into the small text summaries committed under profiles/.

  ncu -i X.ncu-rep --page raw --csv > raw.csv ; python tools/summarize_ncu.py raw raw.csv > profiles/...
  python tools/summarize_ncu.py launches gpurun_out/launches.csv > profiles/...
"""
import collections
import csv
import sys


def launches(path):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[start]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[start + 1:]:
        if len(r) > vi:
            try:
                agg.setdefault(r[ki].split("(")[0].replace("void ", ""), []).append(float(r[vi].replace(",", "")))
            except ValueError:
                pass
    tot = sum(sum(v) for v in agg.values())
    print("%-58s %6s %12s %10s %7s" % ("kernel", "n", "total_us", "avg_us", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print("%-58s %6d %12.1f %10.1f %6.1f%%" % (k[:58], len(v), sum(v) / 1e3, sum(v) / len(v) / 1e3, 100 * sum(v) / tot))


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
            "sm__cycles_elapsed.max"]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("-" * 100)
        for w in want:
            if w in idx:
                print("%-66s %s %s" % (w, r[idx[w]], units[idx[w]]))
        for h in hdr:
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                try:
                    v = float(r[idx[h]])
                except ValueError:
                    continue
                if v > 0.05:
                    print("  %-64s %.3f" % (h.replace("smsp__average_warps_issue_stalled_", "stall ").replace("_per_issue_active.ratio", ""), v))


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2])
