#!/bin/bash
# PDL validation + round-2 ncu evidence (launch lists and --set full captures)
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_decode_gpu.py -x -q --timeout 180 2>&1 | tail -4
echo "=== decode bench: PDL on / off, branch / materialised"
for pdl in 1 0; do for mat in 0 1; do
MC_DECODE_PDL=$pdl MC_MATERIALIZE=$mat timeout 900 python bench.py --workload decode --no-cpu-baseline 2>gpurun_out/r2_prof6_bench.err | tail -1 > gpurun_out/r2_prof6_bench_${pdl}_$mat.json
python -c "
import json; d=json.loads(open('gpurun_out/r2_prof6_bench_${pdl}_$mat.json').read()); r=d['roofline']
print('decode pdl=$pdl mat=$mat', d['value'], 'tok/s', d['ms_per_step'], 'ms frac', r['frac'], 'linears', r['kernel_ms_per_step'], 'ms e2e', d['e2e']['value'], 'ok', d['verification']['decode_vs_prefill_check']['ok'])"
tail -2 gpurun_out/r2_prof6_bench.err
done; done
echo "=== ncu full: decode kernels"
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:streamk|decode_attention' -s 6 -c 9 -o gpurun_out/r2_decode_full -f python tools/profile_decode.py > gpurun_out/r2_prof6_ncu1.log 2>&1
tail -2 gpurun_out/r2_prof6_ncu1.log
echo "=== ncu launch list: prefill step c3"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 700 --csv --log-file gpurun_out/r2_launches_prefill.csv \
    python bench.py --workload prefill --prefill-steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_prof6_ncu2.log 2>&1
tail -2 gpurun_out/r2_prof6_ncu2.log
echo "=== ncu full: pair kernel launches in-step"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:linear2_kernel -s 232 -c 5 -o gpurun_out/r2_prefill_linear2_full -f \
    python bench.py --workload prefill --prefill-steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_prof6_ncu3.log 2>&1
tail -2 gpurun_out/r2_prof6_ncu3.log
} > gpurun_out/r2_prof6.log 2>&1
tail -c 5000 gpurun_out/r2_prof6.log
