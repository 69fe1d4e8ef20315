#!/bin/bash
# end-of-round validation of HEAD: full GPU suite, smoke, default bench, reference arm
mkdir -p gpurun_out
{
echo "=== default bench"
SECONDS=0
timeout 1800 python bench.py > gpurun_out/r2_bench40.json 2> gpurun_out/r2_bench40.err
echo "rc=$? wall=${SECONDS}s"
tail -3 gpurun_out/r2_bench40.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench40.json').read().strip().splitlines()[-1])
def show(k,v):
    print(k, '|', v.get('metric'), v.get('value'), v.get('unit'), 'ms', v.get('ms_per_step'), 'frac', v.get('roofline',{}).get('frac'), 'traffic', v.get('roofline',{}).get('traffic'), 'e2e', v.get('e2e',{}).get('value') if v.get('e2e') else None)
show('primary', d)
for k in d:
    if isinstance(d[k], dict) and 'metric' in d[k]: show(k, d[k])
print('cpu_baseline', d.get('cpu_baseline'))
print('clocks', d.get('clocks'))
PY
echo "=== reference arm"
SECONDS=0
timeout 900 python bench.py --impl reference > gpurun_out/r2_bench40_ref.json 2> gpurun_out/r2_bench40_ref.err
echo "rc=$? wall=${SECONDS}s"; cut -c1-200 gpurun_out/r2_bench40_ref.json
} > gpurun_out/r2_final40.log 2>&1
tail -c 7000 gpurun_out/r2_final40.log
