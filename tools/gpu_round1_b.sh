#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_splice_gpu.py -x -q 2>&1 | tail -30 > gpurun_out/pytest_splice_b.log
for c in c3 c4 c5; do timeout 120 python tools/bench_splice.py --config $c >> gpurun_out/splice_bench_b.log 2>&1; done
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:splice -c 60 --csv --log-file gpurun_out/launches_splice_b.csv python tools/bench_splice.py --config c3 --iters 5 > gpurun_out/ncu_splice_b.log 2>&1
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_splice_gpu.py -x -q -k "fixtures or error or hacky or modal_id" 2>&1 | tail -15 > gpurun_out/sanitizer_splice_b.log
