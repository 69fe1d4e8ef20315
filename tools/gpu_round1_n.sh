#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python tools/debug_linear.py --two-cta > gpurun_out/debug_linear_n.log 2>&1; echo "rc=$?" >> gpurun_out/debug_linear_n.log
timeout 600 python -m pytest tests/test_linear_gpu.py -q -x 2>&1 | tail -5 > gpurun_out/pytest_n.log
MC_LINEAR_UP_TUNING=3 timeout 600 python -m pytest tests/test_prefill_gpu.py -q -x 2>&1 | tail -5 >> gpurun_out/pytest_n.log
MC_LINEAR_UP_TUNING=3 timeout 600 python bench.py --workload prefill --prefill-steps 5 > gpurun_out/bench_prefill_n3.json 2> gpurun_out/bench_prefill_n3.err
