#!/bin/bash
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_decode_gpu.py -x -q --timeout 180 2>&1 | tail -4
echo "=== sweep: one vs two CTAs per SM"
timeout 600 python tools/decode_dev2.py 2>&1 | tail -18
} > gpurun_out/r2_two14.log 2>&1
tail -c 5000 gpurun_out/r2_two14.log
