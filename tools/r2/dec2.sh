#!/bin/bash
# decode kernels (TMA stream-K, [B, heads, capacity, D] cache): parity tests, staged-epilogue tests, timings, decode bench
mkdir -p gpurun_out
{
timeout 1200 python -m pytest tests/test_decode_gpu.py -x -q --timeout 180 2>&1 | tail -30
echo "=== staged epilogue + prefill suite"
timeout 900 python -m pytest tests/test_linear_gpu.py tests/test_prefill_gpu.py -x -q --timeout 300 2>&1 | tail -8
echo "=== timings"
timeout 900 python tools/decode_dev.py --m 1,32,64 --tunings 0,16,32 2>&1 | tail -40
echo "=== decode bench c3"
timeout 900 python bench.py --workload decode --no-cpu-baseline 2>gpurun_out/r2_dec2_bench.err | tail -1 > gpurun_out/r2_dec2_bench.json
python -c "import json; d=json.loads(open('gpurun_out/r2_dec2_bench.json').read()); p=d.pop('prefill',None); print(json.dumps(d, indent=1)); print('prefill', p and p['value'], p and p['ms_per_step'])"
tail -5 gpurun_out/r2_dec2_bench.err
} > gpurun_out/r2_dec2.log 2>&1
tail -c 9000 gpurun_out/r2_dec2.log
