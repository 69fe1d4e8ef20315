#!/bin/bash
# attention v3 (64-key steps, double-buffered scores) bring-up: parity + timing per variant
mkdir -p gpurun_out
{
for t in 0x13 0x23 0x33 0x12; do
  timeout 200 python tools/att_dev.py --tuning $t || echo "variant $t exit code $?"
done
} > gpurun_out/r2_att4.log 2>&1
grep -v "ok$" gpurun_out/r2_att4.log | tail -60
