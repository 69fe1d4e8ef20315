#!/bin/bash
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_decode_gpu.py -x -q --timeout 180 2>&1 | tail -6
} > gpurun_out/r2_dec17.log 2>&1
cat gpurun_out/r2_dec17.log
