#!/bin/bash
# attention v2 with the elect-one issue path: parity + timing per variant
mkdir -p gpurun_out
{
for t in 0x12 0x22 0x32 0x42; do
  timeout 200 python tools/att_dev.py --tuning $t || echo "variant $t exit code $?"
done
} > gpurun_out/r2_att3.log 2>&1
grep -v "parity tuning" gpurun_out/r2_att3.log | tail -40
