#!/bin/bash
# packed-half RoPE (bit-exact?), native row permutation, traffic fields: full suite, A/B, default bench
mkdir -p gpurun_out
{
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -6
echo "=== smoke"
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "=== C3 prefill: hybrid / pair-all / materialised"
run() { local label=$1; shift
  env "$@" timeout 600 python bench.py --workload prefill --no-cpu-baseline 2>/dev/null | tail -1 | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$label', d['value'], 'tok/s', d['ms_per_step'], 'ms  linears', r['kernel_ms_per_step'], 'ms frac', r['frac'], 'clk', d['clocks']['sm_mhz'], 'launches', d['gpu_launches'])"; }
run "c3 hybrid      " MC_X=1
run "c3 pair-all    " MC_LINEAR_UP_TUNING=3
run "c3 hybrid      " MC_X=1
echo "=== default bench"
SECONDS=0
timeout 1800 python bench.py > gpurun_out/r2_bench11.json 2> gpurun_out/r2_bench11.err
echo "rc=$? wall=${SECONDS}s"
tail -3 gpurun_out/r2_bench11.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench11.json').read().strip().splitlines()[-1])
def show(k,v):
    print(k, '|', v.get('metric'), v.get('value'), v.get('unit'), 'ms', v.get('ms_per_step'), 'frac', v.get('roofline',{}).get('frac'), 'traffic', v.get('roofline',{}).get('traffic'), 'e2e', v.get('e2e',{}).get('value') if v.get('e2e') else None)
show('primary', d)
for k in d:
    if isinstance(d[k], dict) and 'metric' in d[k]: show(k, d[k])
PY
echo "=== ncu launch list of one prefill step (NVTX)"
MC_BENCH_NVTX=1 timeout 900 ncu --nvtx --nvtx-include "mc_prefill_step/" --metrics gpu__time_duration.sum --clock-control none \
   --csv --log-file gpurun_out/r2_launches_prefill_step.csv python bench.py --workload prefill --prefill-steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_val11_ncu.log 2>&1
python tools/ncu_summary.py launches gpurun_out/r2_launches_prefill_step.csv | cut -c1-150 | head -30
} > gpurun_out/r2_val11.log 2>&1
tail -c 9000 gpurun_out/r2_val11.log
