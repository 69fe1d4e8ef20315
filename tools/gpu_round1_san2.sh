#!/bin/bash
# compute-sanitizer (memcheck, racecheck, initcheck) over the rewritten TIES kernels and the per-slot host merge
set -x
mkdir -p gpurun_out
SAN="compute-sanitizer --error-exitcode 9 --launch-timeout 120"
K="device_plan_bit_exact and gauss-20-3 or fix_pass or unaligned or bracket_miss or error_behaviour or counting_pass and dtype0 and (8-50 or 256-20 or 3000-20) or quotient and dtype0"
timeout 1500 $SAN --tool memcheck python -m pytest tests/test_ties_gpu.py -q -x -k "$K" > gpurun_out/san2_memcheck_ties.log 2>&1; echo "rc=$?" >> gpurun_out/san2_memcheck_ties.log
timeout 900 $SAN --tool memcheck python -m pytest tests/test_merge_gpu.py -q -x > gpurun_out/san2_memcheck_merge.log 2>&1; echo "rc=$?" >> gpurun_out/san2_memcheck_merge.log
timeout 900 $SAN --tool racecheck python -m pytest tests/test_ties_gpu.py -q -x -k "device_plan_bit_exact and gauss-20-3 and dtype0 or counting_pass and dtype0 and (256-20 or 3000-20)" > gpurun_out/san2_racecheck_ties.log 2>&1; echo "rc=$?" >> gpurun_out/san2_racecheck_ties.log
timeout 900 $SAN --tool initcheck python -m pytest tests/test_merge_gpu.py -q -x -k "host_streaming" > gpurun_out/san2_initcheck_merge_host.log 2>&1; echo "rc=$?" >> gpurun_out/san2_initcheck_merge_host.log
