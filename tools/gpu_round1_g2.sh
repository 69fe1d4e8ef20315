#!/bin/bash
# in-step sweep of the up-projection kernel variant / rasterisation group (C3 prefill, 5 steps each, same box, t0 repeated last)
set -x
mkdir -p gpurun_out
for t in 0 3 259 1027 2051 4 0 3; do
  MC_LINEAR_UP_TUNING=$t timeout 300 python bench.py --workload prefill --prefill-steps 5 --no-cpu-baseline >> gpurun_out/bench_g2_sweep.json 2>> gpurun_out/bench_g2_sweep.err
  echo "tuning=$t" >> gpurun_out/bench_g2_sweep.json
done
