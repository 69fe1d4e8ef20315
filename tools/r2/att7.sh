#!/bin/bash
mkdir -p gpurun_out
{
for t in 0x20 0x10 0x30; do
  extra=""; [ "$t" = "0x20" ] && extra="--cudnn"
  timeout 300 python tools/att_dev.py --tuning $t $extra || echo "variant $t exit code $?"
done
timeout 600 python -m pytest tests/test_attention_gpu.py tests/test_prefill_gpu.py -x -q 2>&1 | tail -5
} > gpurun_out/r2_att7.log 2>&1
grep -v "ok$" gpurun_out/r2_att7.log | tail -60
