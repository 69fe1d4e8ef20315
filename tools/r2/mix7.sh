#!/bin/bash
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_decode_gpu.py -x -q --timeout 180 2>&1 | tail -4
echo "=== knob sweep"
timeout 600 python tools/decode_dev2.py 2>&1 | tail -26
echo "=== decode bench"
for mat in 0 1; do
MC_MATERIALIZE=$mat timeout 900 python bench.py --workload decode --no-cpu-baseline 2>gpurun_out/r2_mix7_bench.err | tail -1 > gpurun_out/r2_mix7_bench_$mat.json
python -c "
import json; d=json.loads(open('gpurun_out/r2_mix7_bench_$mat.json').read()); r=d['roofline']
print('decode mat=$mat', d['value'], 'tok/s', d['ms_per_step'], 'ms frac', r['frac'], 'linears', r['kernel_ms_per_step'], 'ms', r['kernel_achieved_GBps_on_weight_bytes'], 'GB/s e2e', d['e2e']['value'], 'ok', d['verification']['decode_vs_prefill_check']['ok'])"
tail -2 gpurun_out/r2_mix7_bench.err
done
echo "=== ncu launch lists (NVTX ranges): prefill step and decode step, materialised form then branch form"
for mat in 0 1; do
MC_BENCH_NVTX=1 MC_MATERIALIZE=$mat timeout 900 ncu --nvtx --nvtx-include "mc_prefill_step/" --nvtx-include "mc_decode_step/" --metrics gpu__time_duration.sum --clock-control none \
   --csv --log-file gpurun_out/r2_launches_step_mat$mat.csv python bench.py --workload decode --prefill-steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_mix7_ncu_$mat.log 2>&1
tail -1 gpurun_out/r2_mix7_ncu_$mat.log | cut -c1-200
done
} > gpurun_out/r2_mix7.log 2>&1
tail -c 6000 gpurun_out/r2_mix7.log
