#!/bin/bash
mkdir -p gpurun_out
{
echo "=== merge e2e with adaptive warm-up (trace)"
timeout 900 python bench.py --workload merge --merge-all --e2e-trace 2> gpurun_out/r2_e2e16.err | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read())
for k in ('', 'merge_n4', 'merge_c2b'):
    v = d[k] if k else d
    print(k or 'c2', v['value'], 'GB/s  e2e', v['e2e']['value'], 'warm', v['e2e'].get('warmup_passes'), 'bound', v['e2e'].get('pcie_bound_GBps'), 'frac', v['e2e'].get('frac_of_pcie_bound'))"
grep "e2e" gpurun_out/r2_e2e16.err | head -40
} > gpurun_out/r2_e2e16.log 2>&1
cat gpurun_out/r2_e2e16.log
