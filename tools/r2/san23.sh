#!/bin/bash
# compute-sanitizer over the decode-step kernels (small cases): memcheck, racecheck (shared-memory hazards), initcheck
mkdir -p gpurun_out
{
for tool in memcheck racecheck initcheck; do
echo "=== $tool"
timeout 900 compute-sanitizer --tool $tool python -m pytest tests/test_decode_gpu.py -x -q --timeout 600 \
  -k "skinny_multi or argmax or (rope_append and 37) or (rope_append and 129) or (skinny_dual and 96) or (skinny_linear and 200 and bf16) or fused_rope" 2>&1 | tail -6
done
} > gpurun_out/r2_san23.log 2>&1
cat gpurun_out/r2_san23.log
