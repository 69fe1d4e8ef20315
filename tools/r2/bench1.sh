#!/bin/bash
mkdir -p gpurun_out
SECONDS=0
timeout 1500 python bench.py > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err
echo "rc=$? wall=${SECONDS}s"
tail -3 gpurun_out/r2_bench1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench1.json').read().strip().splitlines()[-1])
def show(k,v):
    print(k, v.get('metric'), v.get('value'), v.get('unit'), 'ms', v.get('ms_per_step'), 'frac', v.get('roofline',{}).get('frac'), 'e2e', v.get('e2e',{}).get('value') if v.get('e2e') else None)
show('primary', d)
for k in ('merge_n4','merge_c2b','ties','prefill','prefill_c4','prefill_c5'):
    if k in d: show(k, d[k])
print('cpu', d.get('cpu_baseline'))
print('e2e', d.get('e2e'))
for k in ('prefill','prefill_c4','prefill_c5'):
    if k in d: print(k, 'spot', d[k]['verification']['oracle_spot_check'])
PY
