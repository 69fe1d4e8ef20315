#!/bin/bash
# round 2, call 1: attention v2 bring-up (parity + timing per variant, each in its own process under timeout)
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv
for t in 0x1 0x12 0x22 0x32 0x42 0x132; do
  extra=""; [ "$t" = "0x1" ] && extra="--cudnn"
  timeout 200 python tools/att_dev.py --tuning $t $extra || echo "variant $t exit code $?"
done
timeout 600 python -m pytest tests/test_attention_gpu.py tests/test_prefill_gpu.py -x -q 2>&1 | tail -5
} > gpurun_out/r2_att1.log 2>&1
tail -80 gpurun_out/r2_att1.log
