#!/bin/bash
mkdir -p gpurun_out
{
for t in 0x10 0x20 0x30; do
  timeout 300 python tools/att_dev.py --tuning $t  || echo "variant $t exit code $?"
done
} > gpurun_out/r2_att8.log 2>&1
cat gpurun_out/r2_att8.log | tail -40
