#!/bin/bash
# TIES: ncu --set full of the merge / count / sample kernels after the rewrite, and a sweep of the merge pass' L2 prefetch distance
mkdir -p gpurun_out
{
echo "=== prefetch distance sweep (bench_ties mean 320M)"
for pf in 0 148 296 592 1184 2368; do echo "pf=$pf"; MC_TIES_PREFETCH=$pf timeout 300 python tools/bench_ties.py --func mean --elements 320e6 2>&1 | cut -c1-150; done
echo "=== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ties_merge_kernel|ties_count_kernel|ties_sample_kernel|ties_fix_kernel" -c 5 -o gpurun_out/r2_ties26 -f python tools/bench_ties.py --func mean --elements 320e6 --iters 1 > /dev/null 2>&1
ncu -i gpurun_out/r2_ties26.ncu-rep --page raw --csv > gpurun_out/r2_ties26_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows = list(csv.reader(open('gpurun_out/r2_ties26_raw.csv')))
hdr = rows[0]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct',
        'smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct', 'smsp__warp_issue_stalled_not_selected_per_warp_active.pct']
idx = [hdr.index(w) if w in hdr else None for w in want]
for r in rows[2:]:
    print('---')
    for w, i in zip(want, idx):
        if i is not None: print(f"  {w} = {r[i][:90]} {rows[1][i]}")
PY
} > gpurun_out/r2_ties26.log 2>&1
rm -f gpurun_out/r2_ties26.ncu-rep
tail -c 7000 gpurun_out/r2_ties26.log
