#!/bin/bash
# full GPU suite, smoke, staged-RoPE A/B, default bench + reference arm
mkdir -p gpurun_out
{
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -6
echo "=== smoke"
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "=== A/B staged (incl. RoPE) vs row-per-thread epilogues, C3"
run() { local label=$1; shift
  env "$@" timeout 600 python bench.py --workload prefill --prefill-config c3 --no-cpu-baseline 2>/dev/null | tail -1 | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$label', d['value'], 'tok/s', d['ms_per_step'], 'ms  linears', r['kernel_ms_per_step'], 'ms frac', r['frac'], 'clk', d['clocks']['sm_mhz'])"; }
for rep in 1 2; do
run "staged hybrid   " MC_X=1
run "rowwise hybrid  " MC_LINEAR_EPI_ROWWISE=1
run "staged pair-all " MC_LINEAR_UP_TUNING=3
done
echo "=== default bench"
SECONDS=0
timeout 1800 python bench.py > gpurun_out/r2_bench8.json 2> gpurun_out/r2_bench8.err
echo "rc=$? wall=${SECONDS}s"
tail -3 gpurun_out/r2_bench8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench8.json').read().strip().splitlines()[-1])
def show(k,v):
    print(k, '|', v.get('metric'), v.get('value'), v.get('unit'), 'ms', v.get('ms_per_step'), 'frac', v.get('roofline',{}).get('frac'), 'e2e', v.get('e2e',{}).get('value') if v.get('e2e') else None)
show('primary', d)
for k in d:
    if isinstance(d[k], dict) and 'metric' in d[k]: show(k, d[k])
PY
echo "=== reference arm"
SECONDS=0
timeout 900 python bench.py --impl reference > gpurun_out/r2_bench8_ref.json 2> gpurun_out/r2_bench8_ref.err
echo "rc=$? wall=${SECONDS}s"; cut -c1-300 gpurun_out/r2_bench8_ref.json
} > gpurun_out/r2_full8.log 2>&1
tail -c 6000 gpurun_out/r2_full8.log
