#!/bin/bash
mkdir -p gpurun_out
export ATT_SHAPES=8x3046
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention3 -s 5 -c 1 -f -o gpurun_out/r2_att5 \
  python tools/att_dev.py --tuning 0x13 --no-parity > gpurun_out/r2_att5_ncu.log 2>&1
tail -3 gpurun_out/r2_att5_ncu.log
