#!/bin/bash
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_decode_gpu.py -x -q --timeout 180 2>&1 | tail -4
echo "=== fixed cost"
timeout 600 python tools/decode_fixed_cost.py 2>&1 | tail -12
echo "=== decode bench"
for mat in 0 1; do
MC_MATERIALIZE=$mat timeout 900 python bench.py --workload decode --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']
print('decode mat=$mat', d['value'], 'tok/s', d['ms_per_step'], 'ms frac', r['frac'], 'linears', r['kernel_ms_per_step'], 'ok', d['verification']['decode_vs_prefill_check']['ok'])"
done
} > gpurun_out/r2_fence20.log 2>&1
cat gpurun_out/r2_fence20.log
