#!/bin/bash
# bench.py --workload decode with the nested decode_dense line (both decode forms after one branch-form prefill)
mkdir -p gpurun_out
{
SECONDS=0
timeout 600 python bench.py --workload decode --no-cpu-baseline > gpurun_out/r2_dense39.json 2> gpurun_out/r2_dense39.err
echo "rc=$? wall=${SECONDS}s"; tail -3 gpurun_out/r2_dense39.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_dense39.json').read().strip().splitlines()[-1])
for k, v in (('decode', d), ('decode_dense', d.get('decode_dense'))):
    if v: print(k, v['value'], v['unit'], v['ms_per_step'], 'frac', v['roofline']['frac'], 'e2e', v['e2e']['value'], '|', v['config']['linear_form'][:90], '| launches', v.get('launches_per_step'))
print('prefill', d['prefill']['value'])
PY
} > gpurun_out/r2_dense39.log 2>&1
cat gpurun_out/r2_dense39.log
