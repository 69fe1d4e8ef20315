#!/bin/bash
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_decode_gpu.py -x -q --timeout 180 2>&1 | tail -4
echo "=== knob sweep (new defaults)"
timeout 600 python tools/decode_dev2.py 2>&1 | tail -24
echo "=== decode bench c3 (branch form, then materialised)"
for mat in 0 1; do
MC_MATERIALIZE=$mat timeout 900 python bench.py --workload decode --no-cpu-baseline 2>gpurun_out/r2_mix5_bench.err | tail -1 > gpurun_out/r2_mix5_bench_$mat.json
python -c "
import json; d=json.loads(open('gpurun_out/r2_mix5_bench_$mat.json').read()); p=d.pop('prefill',None); r=d['roofline']
print('decode mat=$mat', d['value'], 'tok/s', d['ms_per_step'], 'ms frac', r['frac'], 'linears', r['kernel_ms_per_step'], 'ms', r['kernel_achieved_GBps_on_weight_bytes'], 'GB/s e2e', d['e2e']['value'], d['verification']['decode_vs_prefill_check'])
print('prefill', p and p['value'], p and p['ms_per_step'])"
tail -3 gpurun_out/r2_mix5_bench.err
done
echo "=== staged epilogue A/B"
bash tools/r2/epi1.sh
} > gpurun_out/r2_mix5.log 2>&1
tail -c 9000 gpurun_out/r2_mix5.log
