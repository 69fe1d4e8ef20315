#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --workload prefill --prefill-steps 5 > gpurun_out/bench_prefill_l0.json 2> gpurun_out/bench_prefill_l0.err
MC_LINEAR_UP_TUNING=3 timeout 600 python bench.py --workload prefill --prefill-steps 5 > gpurun_out/bench_prefill_l3.json 2> gpurun_out/bench_prefill_l3.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:linear|rmsnorm|rope|silu|flash|fmha' -s 1000 -c 500 --csv --log-file gpurun_out/launches_prefill_l.csv \
    python bench.py --workload prefill --prefill-steps 1 --warmup 1 > gpurun_out/ncu_launches_l.log 2>&1
