#!/bin/bash
# ncu --set full of the v2 attention kernel at B=8, S=3046 (one launch), plus the launch list of a short run
mkdir -p gpurun_out
export ATT_SHAPES=8x3046
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention2 -s 5 -c 1 -f -o gpurun_out/r2_att2 \
  python tools/att_dev.py --tuning ${ATT_TUNING:-0x12} --no-parity > gpurun_out/r2_att2_ncu.log 2>&1
tail -5 gpurun_out/r2_att2_ncu.log
