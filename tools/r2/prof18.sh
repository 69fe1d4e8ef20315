#!/bin/bash
# DRAM bytes of every kernel of ONE C3 prefill step (NVTX range), branch form and materialised form
mkdir -p gpurun_out
{
for mat in 0 1; do
MC_BENCH_NVTX=1 MC_MATERIALIZE=$mat timeout 900 ncu --nvtx --nvtx-include "mc_prefill_step/" --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
   --csv --log-file gpurun_out/r2_traffic_prefill_mat$mat.csv python bench.py --workload prefill --prefill-steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_prof18_$mat.log 2>&1
tail -1 gpurun_out/r2_prof18_$mat.log | cut -c1-120
done
} > gpurun_out/r2_prof18.log 2>&1
cat gpurun_out/r2_prof18.log
