#!/bin/bash
# full GPU test suite + prefill A/B (native attention vs library) + the default bench line
mkdir -p gpurun_out
{
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "=== prefill c3 native attention"
timeout 600 python bench.py --workload prefill --prefill-config c3 --no-cpu-baseline 2>&1 | tail -1
echo "=== prefill c3 library attention"
MC_ATTENTION_NATIVE=0 timeout 600 python bench.py --workload prefill --prefill-config c3 --no-cpu-baseline 2>&1 | tail -1
echo "=== default bench"
/usr/bin/time -v timeout 1500 python bench.py 2> gpurun_out/r2_full1_bench.err | tail -1
grep -E "Elapsed|Maximum resident" gpurun_out/r2_full1_bench.err
tail -5 gpurun_out/r2_full1_bench.err
} > gpurun_out/r2_full1.log 2>&1
tail -c 6000 gpurun_out/r2_full1.log
