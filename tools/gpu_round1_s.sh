#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/pytest_s.log
timeout 600 python bench.py --workload prefill --prefill-steps 5 > gpurun_out/bench_prefill_s.json 2> gpurun_out/bench_prefill_s.err
for c in c3 c4 c5; do timeout 120 python tools/bench_splice.py --config $c >> gpurun_out/splice_bench_s.log 2>&1; done
timeout 900 ncu --set full --clock-control none -k regex:merge_kernel -s 3 -c 1 -o gpurun_out/prof_merge_n1_s -f python bench.py --workload merge --steps 3 --warmup 3 --no-e2e > gpurun_out/ncu_merge_n1_s.log 2>&1
