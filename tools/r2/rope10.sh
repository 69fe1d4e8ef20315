#!/bin/bash
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_linear_gpu.py tests/test_prefill_gpu.py tests/test_attention_gpu.py -x -q --timeout 300 2>&1 | tail -4
echo "=== packed-half RoPE epilogue inside the C3 / C4 prefill (previous build on the earlier box: staged hybrid 78.8 k)"
run() { local label=$1; shift
  env "$@" timeout 600 python bench.py --workload prefill --no-cpu-baseline 2>/dev/null | tail -1 | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$label', d['value'], 'tok/s', d['ms_per_step'], 'ms  linears', r['kernel_ms_per_step'], 'ms frac', r['frac'], 'clk', d['clocks']['sm_mhz'])"; }
run "c3 hybrid      " MC_X=1
run "c3 pair-all    " MC_LINEAR_UP_TUNING=3
run "c3 hybrid      " MC_X=1
run "c3 materialised" MC_MATERIALIZE=1
echo "=== ncu: q/k/v launch in-step"
timeout 900 ncu --set full --clock-control none -k regex:linear2_kernel -s 232 -c 3 -o gpurun_out/r2_prefill_linear2_rope -f \
    python bench.py --workload prefill --prefill-steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_rope10_ncu.log 2>&1
tail -1 gpurun_out/r2_rope10_ncu.log | cut -c1-100
bash tools/r2/prof9.sh
} > gpurun_out/r2_rope10.log 2>&1
tail -c 8000 gpurun_out/r2_rope10.log
