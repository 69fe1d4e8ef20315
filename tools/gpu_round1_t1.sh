#!/bin/bash
# TIES merge pass on the FMA pipe (ties_one_fast), metadata-pipelined counting pass, per-slot plans in mc_merge_host
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ties_gpu.py tests/test_merge_gpu.py -q -x --timeout 300 2>&1 | tail -15 > gpurun_out/pytest_t1_ties.log
for args in "--func mean" "--func sum" "--func max --kind neg" "--func sum --kind zeros" "--func sum --src 4 --elements 320e6" "--func sum --dtype f16" "--func mean --dtype f16"; do
  timeout 120 python tools/bench_ties.py $args >> gpurun_out/bench_ties_t1.log 2>&1
done
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:ties -c 60 --csv --log-file gpurun_out/launches_ties_t1_sum.csv python tools/bench_ties.py --iters 2 --func sum > gpurun_out/ncu_ties_t1.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:ties -c 60 --csv --log-file gpurun_out/launches_ties_t1_mean.csv python tools/bench_ties.py --iters 2 --func mean >> gpurun_out/ncu_ties_t1.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"ties_merge_kernel|ties_count_kernel" -c 4 -o gpurun_out/ties_t1_full -f python tools/bench_ties.py --iters 1 --func mean >> gpurun_out/ncu_ties_t1.log 2>&1
timeout 600 python bench.py --workload merge > gpurun_out/bench_t1_merge.json 2> gpurun_out/bench_t1_merge.err
timeout 300 python bench.py --workload ties > gpurun_out/bench_t1_ties.json 2> gpurun_out/bench_t1_ties.err
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -8 > gpurun_out/pytest_t1_all.log
