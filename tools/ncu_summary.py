#!/usr/bin/env python
"""Summarise ncu output for profiles/: `ncu_summary.py raw <rep> [kernel-regex]` (metrics of a --set full capture)
or `ncu_summary.py launches <csv>` (per-kernel totals and shares of a gpu__time_duration launch list)."""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict

RAW_METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__cycles_active.avg",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
]


def raw(rep, pattern=None):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        if pattern and not re.search(pattern, name):
            continue
        print("kernel:", name)
        for m in hdr:
            if m in RAW_METRICS or "pipe_tensor" in m:
                i = hdr.index(m)
                print(f"  {m} = {r[i]} {units[i]}")
        print()


def launches(path):
    text = open(path).read()
    start = text.find('"ID"')
    rows = list(csv.DictReader(io.StringIO(text[start:])))
    tot = defaultdict(lambda: [0, 0.0])
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v_us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3 if unit in ("ms", "msecond") else v)
        k = re.sub(r"\(.*", "", r["Kernel Name"])[:110]
        tot[k][0] += 1
        tot[k][1] += v_us
    total = sum(v[1] for v in tot.values())
    print(f"{'kernel':110s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}")
    for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:110s} {n:8d} {t:12.1f} {t / n:10.1f} {t / total:7.2%}")
    print(f"total {total:.1f} us over {sum(v[0] for v in tot.values())} launches")


if __name__ == "__main__":
    if sys.argv[1] == "raw":
        raw(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
    else:
        launches(sys.argv[2])
