#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_e.log
timeout 900 python bench.py > gpurun_out/bench_e.json 2> gpurun_out/bench_e.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_e.json 2> gpurun_out/bench_ref_e.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:linear_kernel|splice|rmsnorm|rope|silu|flash|fmha|route_tile|merge_kernel' -c 3000 --csv --log-file gpurun_out/launches_prefill_e.csv \
    python bench.py --workload prefill --prefill-steps 1 --warmup 1 > gpurun_out/ncu_launches_e.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:linear_kernel -s 8 -c 8 -o gpurun_out/prof_linear_e -f \
    python tools/profile_linear.py 2 > gpurun_out/ncu_full_e.log 2>&1
ls -la gpurun_out | tail -20
