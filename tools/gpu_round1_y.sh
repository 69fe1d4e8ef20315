#!/bin/bash
# TIES v3 (bracket/count tuning, -0 fix) + linear3 (256x256 CTA-pair tiles, overlapped epilogue) bring-up
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_linear_gpu.py -q -x --timeout 120 2>&1 | tail -15 > gpurun_out/pytest_linear_y.log
timeout 900 python -m pytest tests/test_ties_gpu.py tests/test_prefill_gpu.py -q -x --timeout 300 2>&1 | tail -15 > gpurun_out/pytest_y.log
for args in "--func mean" "--func sum" "--func max --kind neg" "--func sum --kind zeros" "--func sum --src 4 --elements 320e6" "--func sum --dtype f16"; do
  timeout 120 python tools/bench_ties.py $args >> gpurun_out/bench_ties_y.log 2>&1
done
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:ties -c 60 --csv --log-file gpurun_out/launches_ties_y.csv python tools/bench_ties.py --iters 2 --func sum > gpurun_out/ncu_ties_y.log 2>&1
timeout 300 python tools/profile_two_cta.py > gpurun_out/two_cta_y.log 2>&1
MC_LINEAR_UP_TUNING=4 timeout 600 python bench.py --workload prefill --prefill-steps 5 > gpurun_out/bench_prefill_y_pair256.json 2> gpurun_out/bench_prefill_y_pair256.err
timeout 600 python bench.py --workload prefill --prefill-steps 5 > gpurun_out/bench_prefill_y.json 2> gpurun_out/bench_prefill_y.err
