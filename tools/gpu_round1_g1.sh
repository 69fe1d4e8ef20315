#!/bin/bash
# pair-kernel rasterisation: group size in rows instead of tiles; standalone sweep + in-step comparison
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_linear_gpu.py -q -x --timeout 120 2>&1 | tail -4 > gpurun_out/pytest_g1_linear.log
SH="31360,4096,4096;31360,11008,4096;31360,4096,11008"
MC_TIME=30 timeout 600 python tools/profile_variants.py "$SH" "0,3:1,3:2,3:4,3:8,3:32,4:2,4:4,4:8,4:16,4:32,3,4" 3 > gpurun_out/variants_g1.log 2>&1
for t in 0 3 4; do
  MC_LINEAR_UP_TUNING=$t timeout 600 python bench.py --workload prefill --prefill-steps 5 > gpurun_out/bench_g1_prefill_t$t.json 2> gpurun_out/bench_g1_prefill_t$t.err
done
