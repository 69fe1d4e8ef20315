#!/bin/bash
# metrics parity + compute-sanitizer (memcheck, racecheck) over small-shape runs of the HBM-bound kernel families
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ties_gpu.py -q -x --timeout 600 2>&1 | tail -8 > gpurun_out/pytest_ties_san.log
SAN="compute-sanitizer --error-exitcode 9 --launch-timeout 120"
K="device_plan_bit_exact and gauss-20-3 or fix_pass or unaligned or bracket_miss or error_behaviour"
timeout 1500 $SAN --tool memcheck python -m pytest tests/test_ties_gpu.py -q -x -k "$K" > gpurun_out/san_memcheck_ties.log 2>&1; echo "rc=$?" >> gpurun_out/san_memcheck_ties.log
timeout 900 $SAN --tool memcheck python -m pytest tests/test_merge_gpu.py tests/test_splice_gpu.py -q -x > gpurun_out/san_memcheck_merge_splice.log 2>&1; echo "rc=$?" >> gpurun_out/san_memcheck_merge_splice.log
timeout 900 $SAN --tool racecheck python -m pytest tests/test_ties_gpu.py -q -x -k "device_plan_bit_exact and gauss-20-3 and dtype0 or bracket_miss" > gpurun_out/san_racecheck_ties.log 2>&1; echo "rc=$?" >> gpurun_out/san_racecheck_ties.log
timeout 900 $SAN --tool memcheck python -m pytest tests/test_linear_gpu.py -q -x -k "row_map or rowwise_glue or plain_linear and 129" > gpurun_out/san_memcheck_linear.log 2>&1; echo "rc=$?" >> gpurun_out/san_memcheck_linear.log
