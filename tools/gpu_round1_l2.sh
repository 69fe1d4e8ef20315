#!/bin/bash
# ncu launch list of the prefill step with the auto kernel choice (512x256 pair kernel for the base + LoRA-up launches)
set -x
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:linear|rmsnorm|rope|silu|flash|fmha|cudnn|splice|route|gather' -s 1000 -c 600 --csv --log-file gpurun_out/launches_prefill_l2.csv \
    python bench.py --workload prefill --prefill-steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches_l2.log 2>&1
timeout 400 python bench.py --workload merge > gpurun_out/bench_l2_merge.json 2> gpurun_out/bench_l2_merge.err
