#!/bin/bash
mkdir -p gpurun_out
{
echo "=== initcheck: kernel-level decode tests"
timeout 900 compute-sanitizer --tool initcheck --print-limit 4 python -m pytest tests/test_decode_gpu.py -x -q --timeout 600 \
  -k "skinny_multi or argmax or (rope_append and 37) or (rope_append and 129) or (skinny_dual and 96) or (skinny_linear and 200 and bf16)" 2>&1 | grep -v "Host Frame\|^=========$" | tail -40
echo "=== initcheck: decode loop (model level)"
timeout 900 compute-sanitizer --tool initcheck --print-limit 4 python -m pytest tests/test_decode_gpu.py -x -q --timeout 600 -k "fused_rope and FUSED_ROPE and False" 2>&1 | grep -v "Host Frame\|^=========$" | tail -40
} > gpurun_out/r2_init24.log 2>&1
cat gpurun_out/r2_init24.log | cut -c1-250
