#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/decode_fixed_cost.py > gpurun_out/r2_fix19.log 2>&1
cat gpurun_out/r2_fix19.log | tail -20
