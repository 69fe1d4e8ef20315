#!/bin/bash
mkdir -p gpurun_out
{
for i in 1 2; do
timeout 900 python -m pytest tests/test_decode_gpu.py -x -q --timeout 300 -k "graph_equals_eager or decode_matches or pdl" 2>&1 | tail -60
done
timeout 900 python -m pytest tests/test_cabi.py tests/test_decode_gpu.py -m gpu -x -q --timeout 300 2>&1 | tail -40
} > gpurun_out/r2_dbg29.log 2>&1
tail -c 9000 gpurun_out/r2_dbg29.log
