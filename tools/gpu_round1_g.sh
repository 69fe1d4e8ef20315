#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_linear_gpu.py tests/test_prefill_gpu.py -q -x 2>&1 | tail -15 > gpurun_out/pytest_g.log
timeout 600 python bench.py --workload prefill --prefill-steps 5 > gpurun_out/bench_prefill_g.json 2> gpurun_out/bench_prefill_g.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:linear_kernel|rmsnorm|rope|silu|flash|fmha' -s 1000 -c 600 --csv --log-file gpurun_out/launches_prefill_g.csv \
    python bench.py --workload prefill --prefill-steps 1 --warmup 1 > gpurun_out/ncu_launches_g.log 2>&1
