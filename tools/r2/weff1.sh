#!/bin/bash
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_prefill_gpu.py tests/test_linear_gpu.py -x -q 2>&1 | tail -25
echo "=== prefill c3 branch form"
timeout 600 python bench.py --workload prefill --prefill-config c3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-400
echo "=== prefill c3 materialised"
MC_MATERIALIZE=1 timeout 600 python bench.py --workload prefill --prefill-config c3 --no-cpu-baseline 2>&1 | tail -3 | cut -c1-1500
} > gpurun_out/r2_weff1.log 2>&1
tail -c 5000 gpurun_out/r2_weff1.log
