#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_attention_gpu.py -q -x --timeout 120 2>&1 | tail -5 > gpurun_out/pytest_att3.log
timeout 300 python tools/bench_attention.py > gpurun_out/bench_att3.log 2>&1
MC_ATTENTION_NATIVE=1 timeout 600 python -m pytest tests/test_prefill_gpu.py -q -x --timeout 300 2>&1 | tail -5 > gpurun_out/pytest_prefill_native_att.log
MC_ATTENTION_NATIVE=1 timeout 600 python bench.py --workload prefill --prefill-steps 5 > gpurun_out/bench_prefill_native_att.json 2> gpurun_out/bench_prefill_native_att.err
