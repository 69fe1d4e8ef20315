#!/bin/bash
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_decode_gpu.py tests/test_linear_gpu.py -x -q --timeout 180 2>&1 | tail -5
echo "=== sweep"
timeout 600 python tools/decode_dev2.py 2>&1 | tail -18
echo "=== decode bench: stream-K vs K-slice major"
for tun in 0 32768; do for mat in 0 1; do
MC_DECODE_SKINNY_TUNING=$tun MC_MATERIALIZE=$mat timeout 900 python bench.py --workload decode --no-cpu-baseline 2>gpurun_out/r2_ks12_bench.err | tail -1 > gpurun_out/r2_ks12_bench_${tun}_$mat.json
python -c "
import json; d=json.loads(open('gpurun_out/r2_ks12_bench_${tun}_$mat.json').read()); r=d['roofline']
print('decode tuning=$tun mat=$mat', d['value'], 'tok/s', d['ms_per_step'], 'ms frac', r['frac'], 'linears', r['kernel_ms_per_step'], 'ms', r['kernel_achieved_GBps_on_weight_bytes'], 'GB/s e2e', d['e2e']['value'], 'ok', d['verification']['decode_vs_prefill_check']['ok'])"
tail -2 gpurun_out/r2_ks12_bench.err
done; done
} > gpurun_out/r2_ks12.log 2>&1
tail -c 7000 gpurun_out/r2_ks12.log
