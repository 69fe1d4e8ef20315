#!/bin/bash
# auto = pair kernel except up_proj(+SiLU*mul) and o_proj on the single-CTA kernel, vs the pure pair kernel and the pure single-CTA kernel
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_prefill_gpu.py -q -x --timeout 300 2>&1 | tail -3 > gpurun_out/pytest_g5.log
for t in auto 3 auto 3 0; do
  MC_LINEAR_UP_TUNING=$t timeout 300 python bench.py --workload prefill --prefill-steps 5 --no-cpu-baseline >> gpurun_out/bench_g5_sweep.json 2>> gpurun_out/bench_g5_sweep.err
  echo "tuning=$t" >> gpurun_out/bench_g5_sweep.json
done
