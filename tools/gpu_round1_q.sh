#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/pytest_q.log
timeout 900 python bench.py > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err
timeout 600 python bench.py --workload prefill --prefill-config c4 --prefill-steps 5 > gpurun_out/bench_c4_q.json 2> gpurun_out/bench_c4_q.err
timeout 600 python bench.py --workload prefill --prefill-config c5 --prefill-steps 3 > gpurun_out/bench_c5_q.json 2> gpurun_out/bench_c5_q.err
