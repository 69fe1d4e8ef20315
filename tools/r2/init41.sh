#!/bin/bash
# compute-sanitizer initcheck on the TIES device plan (new last-CTA steps, arrival counters, sampling stride)
mkdir -p gpurun_out
{
timeout 100 compute-sanitizer --tool initcheck --print-limit 5 python -m pytest tests/test_ties_gpu.py -x -q --timeout 90 -k "(device_plan_bit_exact and bfloat16 and 3-gauss) or sampled_select or bracket_miss" 2>&1 | grep -v "Host Frame\|^=========$" | tail -8
} > gpurun_out/r2_init41.log 2>&1
cat gpurun_out/r2_init41.log
