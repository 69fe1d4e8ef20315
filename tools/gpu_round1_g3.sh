#!/bin/bash
# 512x256 pair kernel with 16 epilogue warps (column halves): parity, then in-step against the single-CTA default
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_linear_gpu.py tests/test_prefill_gpu.py -q -x --timeout 300 2>&1 | tail -5 > gpurun_out/pytest_g3.log
for t in 3 0 3 0; do
  MC_LINEAR_UP_TUNING=$t timeout 300 python bench.py --workload prefill --prefill-steps 5 --no-cpu-baseline >> gpurun_out/bench_g3_sweep.json 2>> gpurun_out/bench_g3_sweep.err
  echo "tuning=$t" >> gpurun_out/bench_g3_sweep.json
done
MC_TIME=30 timeout 300 python tools/profile_variants.py "31360,4096,4096;31360,11008,4096;31360,4096,11008" "0,3,0,3" 3 > gpurun_out/variants_g3.log 2>&1
