#!/bin/bash
# TIES: counting pass with 1-4 bin SIMD counters, packed trim in the merge pass, prefetch distance sweep
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ties_gpu.py -q -x --timeout 300 2>&1 | tail -15 > gpurun_out/pytest_t5_ties.log
for pf in 592; do
  for args in "--func mean" "--func sum"; do
    echo "MC_TIES_PREFETCH=$pf $args" >> gpurun_out/bench_ties_t5.log
    MC_TIES_PREFETCH=$pf timeout 120 python tools/bench_ties.py --iters 40 $args >> gpurun_out/bench_ties_t5.log 2>&1
  done
done
for args in "--func max --kind neg" "--func sum --kind zeros" "--func sum --src 4 --elements 320e6" "--func sum --dtype f16" "--func mean --dtype f16" "--func mean --src 2" "--func mean --src 8 --elements 80e6" "--func mean --elements 320e6"; do
  timeout 120 python tools/bench_ties.py $args >> gpurun_out/bench_ties_t5.log 2>&1
done
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:ties -c 60 --csv --log-file gpurun_out/launches_ties_t5_sum.csv python tools/bench_ties.py --iters 2 --func sum > gpurun_out/ncu_ties_t5.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:ties -c 60 --csv --log-file gpurun_out/launches_ties_t5_mean.csv python tools/bench_ties.py --iters 2 --func mean >> gpurun_out/ncu_ties_t5.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"ties_merge_kernel|ties_count_kernel" -c 3 -o gpurun_out/ties_t5_full -f python tools/bench_ties.py --iters 1 --func mean >> gpurun_out/ncu_ties_t5.log 2>&1
timeout 300 python bench.py --workload ties > gpurun_out/bench_t5_ties.json 2> gpurun_out/bench_t5_ties.err
