#!/bin/bash
# TIES merge bring-up: parity tests, timing, per-kernel ncu durations + DRAM traffic; fused prefill launch list
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ties_gpu.py -q -x 2>&1 | tail -15 > gpurun_out/pytest_ties_w.log
for args in "--func mean" "--func sum" "--func max --kind neg" "--func sum --kind zeros" "--func mean --dtype f32 --elements 80e6" "--func sum --src 4 --elements 320e6"; do
  timeout 120 python tools/bench_ties.py $args >> gpurun_out/bench_ties_w.log 2>&1
done
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:ties -c 40 --csv --log-file gpurun_out/launches_ties_w.csv python tools/bench_ties.py --iters 2 > gpurun_out/ncu_ties_w.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:linear|rmsnorm|rope|silu|flash|fmha|cudnn|splice|route' -s 1000 -c 600 --csv --log-file gpurun_out/launches_prefill_w.csv \
    python bench.py --workload prefill --prefill-steps 1 --warmup 1 > gpurun_out/ncu_launches_w.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/pytest_w.log
