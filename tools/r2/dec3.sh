#!/bin/bash
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_decode_gpu.py -x -q --timeout 180 -k "skinny or decode_steps" 2>&1 | tail -4
echo "=== knob sweep"
timeout 900 python tools/decode_dev2.py 2>&1 | tail -45
echo "=== decode bench c3"
timeout 900 python bench.py --workload decode --no-cpu-baseline 2>gpurun_out/r2_dec4_bench.err | tail -1 > gpurun_out/r2_dec4_bench.json
python -c "import json; d=json.loads(open('gpurun_out/r2_dec4_bench.json').read()); p=d.pop('prefill',None); print(json.dumps(d, indent=1)); print('prefill', p and p['value'], p and p['ms_per_step'])"
tail -5 gpurun_out/r2_dec4_bench.err
} > gpurun_out/r2_dec4.log 2>&1
tail -c 9000 gpurun_out/r2_dec4.log
