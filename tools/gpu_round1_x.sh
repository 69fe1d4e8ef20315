#!/bin/bash
# TIES v2 (restructured merge math, sparse fix-up, sampled bracket) + modality-major prefill rows
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ties_gpu.py tests/test_linear_gpu.py tests/test_prefill_gpu.py -q -x 2>&1 | tail -15 > gpurun_out/pytest_x.log
for args in "--func mean" "--func sum" "--func max --kind neg" "--func sum --kind zeros" "--func mean --dtype f32 --elements 80e6" "--func sum --src 4 --elements 320e6" "--func sum --dtype f16"; do
  timeout 120 python tools/bench_ties.py $args >> gpurun_out/bench_ties_x.log 2>&1
done
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:ties -c 60 --csv --log-file gpurun_out/launches_ties_x.csv python tools/bench_ties.py --iters 2 --func sum > gpurun_out/ncu_ties_x.log 2>&1
timeout 600 python bench.py --workload prefill --prefill-steps 5 > gpurun_out/bench_prefill_x.json 2> gpurun_out/bench_prefill_x.err
MC_MODALITY_MAJOR=0 timeout 600 python bench.py --workload prefill --prefill-steps 5 > gpurun_out/bench_prefill_x_seqorder.json 2> gpurun_out/bench_prefill_x_seqorder.err
MC_LINEAR_UP_TUNING=3 timeout 600 python bench.py --workload prefill --prefill-steps 5 > gpurun_out/bench_prefill_x_pair.json 2> gpurun_out/bench_prefill_x_pair.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:linear|rmsnorm|rope|silu|flash|fmha|cudnn|splice|route|gather' -s 1000 -c 600 --csv --log-file gpurun_out/launches_prefill_x.csv \
    python bench.py --workload prefill --prefill-steps 1 --warmup 1 > gpurun_out/ncu_launches_x.log 2>&1
