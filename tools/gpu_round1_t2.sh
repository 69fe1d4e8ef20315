#!/bin/bash
# TIES: narrow-bracket SIMD counters in the counting pass, L2 bulk prefetch (one CTA generation ahead) in the merge pass
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ties_gpu.py -q -x --timeout 300 2>&1 | tail -15 > gpurun_out/pytest_t2_ties.log
for pf in 0 -1 592 1776; do
  for args in "--func mean" "--func sum"; do
    echo "MC_TIES_PREFETCH=$pf $args" >> gpurun_out/bench_ties_t2.log
    if [ "$pf" = "-1" ]; then timeout 120 python tools/bench_ties.py $args >> gpurun_out/bench_ties_t2.log 2>&1
    else MC_TIES_PREFETCH=$pf timeout 120 python tools/bench_ties.py $args >> gpurun_out/bench_ties_t2.log 2>&1; fi
  done
done
for args in "--func max --kind neg" "--func sum --kind zeros" "--func sum --src 4 --elements 320e6" "--func sum --dtype f16" "--func mean --dtype f16" "--func mean --src 2" "--func mean --src 8 --elements 80e6"; do
  timeout 120 python tools/bench_ties.py $args >> gpurun_out/bench_ties_t2.log 2>&1
done
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:ties -c 60 --csv --log-file gpurun_out/launches_ties_t2_sum.csv python tools/bench_ties.py --iters 2 --func sum > gpurun_out/ncu_ties_t2.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:ties -c 60 --csv --log-file gpurun_out/launches_ties_t2_mean.csv python tools/bench_ties.py --iters 2 --func mean >> gpurun_out/ncu_ties_t2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"ties_merge_kernel|ties_count_kernel" -c 3 -o gpurun_out/ties_t2_full -f python tools/bench_ties.py --iters 1 --func mean >> gpurun_out/ncu_ties_t2.log 2>&1
timeout 300 python bench.py --workload ties > gpurun_out/bench_t2_ties.json 2> gpurun_out/bench_t2_ties.err
timeout 600 python bench.py --workload merge > gpurun_out/bench_t2_merge.json 2> gpurun_out/bench_t2_merge.err
