#!/bin/bash
# full GPU suite + default bench (merge + ties + prefill) + TIES per-kernel launch list and ncu --set full of its two streaming kernels
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -15 > gpurun_out/pytest_z2.log
timeout 900 python bench.py > gpurun_out/bench_z2.json 2> gpurun_out/bench_z2.err
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:ties -c 60 --csv --log-file gpurun_out/launches_ties_z2.csv python tools/bench_ties.py --iters 2 --func sum > gpurun_out/ncu_ties_z2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:ties_count_kernel|ties_merge_kernel' -s 4 -c 2 -o gpurun_out/prof_ties_z2 -f python tools/bench_ties.py --iters 1 --func sum > gpurun_out/ncu_full_ties_z2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:linear|rmsnorm|rope|silu|flash|fmha|cudnn|splice|route|gather' -s 1000 -c 600 --csv --log-file gpurun_out/launches_prefill_z2.csv \
    python bench.py --workload prefill --prefill-steps 1 --warmup 1 > gpurun_out/ncu_launches_z2.log 2>&1
