#!/bin/bash
# the driver's scaling launch at N = 8 (default bench: merge + nested merges + prefill C3/C4/C5 + decode), one box
mkdir -p gpurun_out
SECONDS=0
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 50 --warmup 5 \
   > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
echo "rc=$? wall=${SECONDS}s"
tail -3 gpurun_out/r2_bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n2.json').read().strip().splitlines()[-1])
def show(k,v):
    e=v.get('e2e') or {}
    print(k, '|', v.get('metric'), v.get('value'), v.get('unit'), 'ms', v.get('ms_per_step'), 'frac', v.get('roofline',{}).get('frac'), 'e2e', e.get('value'), 'bound', e.get('pcie_bound_GBps'), 'fracbound', e.get('frac_of_pcie_bound'))
show('primary', d)
for k in d:
    if isinstance(d[k], dict) and 'metric' in d[k]: show(k, d[k])
print('probe', (d.get('e2e') or {}).get('pcie_probe'))
print('verification', {k: d[k]['verification'].get('probe_request_identical_across_ranks') for k in d if isinstance(d[k], dict) and 'verification' in d[k]})
PY
