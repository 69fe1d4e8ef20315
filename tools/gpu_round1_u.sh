#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/pytest_u.log
timeout 300 python tools/diag_prefill.py > gpurun_out/diag_u.log 2>&1
timeout 600 python bench.py --workload prefill --prefill-steps 5 > gpurun_out/bench_prefill_u.json 2> gpurun_out/bench_prefill_u.err
for c in c3 c5; do timeout 120 python tools/bench_splice.py --config $c >> gpurun_out/splice_bench_u.log 2>&1; done
