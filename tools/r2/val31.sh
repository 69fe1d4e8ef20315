#!/bin/bash
# full GPU suite (new: linear RoPE scaling, host-buffer TIES outputs) and the TIES bench line with its new host-buffer e2e leg
mkdir -p gpurun_out
{
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -8
echo "=== bench.py --workload ties"
timeout 600 python bench.py --workload ties 2>gpurun_out/r2_val31.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['metric'], d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['traffic'])
print('e2e', json.dumps(d.get('e2e'))[:700])
print('clocks', d.get('clocks'))"
tail -3 gpurun_out/r2_val31.err
} > gpurun_out/r2_val31.log 2>&1
tail -c 4000 gpurun_out/r2_val31.log
