#!/bin/bash
# single-CTA vs 512x256 pair kernel inside the C4 / C5 prefill (after the rasterisation fix), same box, alternating
set -x
mkdir -p gpurun_out
for cfg in c4 c5; do
  for t in 0 3 0 3; do
    MC_LINEAR_UP_TUNING=$t timeout 300 python bench.py --workload prefill --prefill-config $cfg --prefill-steps 3 --no-cpu-baseline >> gpurun_out/bench_g4_sweep.json 2>> gpurun_out/bench_g4_sweep.err
    echo "cfg=$cfg tuning=$t" >> gpurun_out/bench_g4_sweep.json
  done
done
