#!/bin/bash
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_decode_gpu.py -x -q --timeout 180 2>&1 | tail -4
echo "=== decode bench: RMSNorm fused into the consuming skinny launch vs separate launches"
for fused in 1 0; do for mat in 0 1; do
MC_DECODE_FUSED_NORM=$fused MC_MATERIALIZE=$mat timeout 900 python bench.py --workload decode --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']
print('decode fused_norm=$fused mat=$mat', d['value'], 'tok/s', d['ms_per_step'], 'ms frac', r['frac'], 'launches/step', d['launches_per_step'], 'ok', d['verification']['decode_vs_prefill_check']['ok'])"
done; done
} > gpurun_out/r2_norm22.log 2>&1
cat gpurun_out/r2_norm22.log
