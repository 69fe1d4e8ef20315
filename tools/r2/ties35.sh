#!/bin/bash
# ncu --set full of the TIES kernels at the final state (prefetch distance 1 x SMs), bench.py's adapter set
mkdir -p gpurun_out
{
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ties_merge_kernel|ties_count_kernel|ties_sample_kernel|ties_fix_kernel" -s 15 -c 5 -o gpurun_out/r2_ties35 -f python bench.py --workload ties --no-e2e > /dev/null 2>&1
ncu -i gpurun_out/r2_ties35.ncu-rep --page raw --csv > gpurun_out/r2_ties35_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows = list(csv.reader(open('gpurun_out/r2_ties35_raw.csv')))
hdr = rows[0]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'lts__t_sector_hit_rate.pct', 'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio', 'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct']
idx = [hdr.index(w) if w in hdr else None for w in want]
for r in rows[2:]:
    print('---')
    for w, i in zip(want, idx):
        if i is not None: print(f"  {w} = {r[i][:90]} {rows[1][i]}")
PY
} > gpurun_out/r2_ties35.log 2>&1
rm -f gpurun_out/r2_ties35.ncu-rep
tail -c 6000 gpurun_out/r2_ties35.log
