#!/bin/bash
# decode kernels: parity tests, then kernel timings at 7B shapes
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_decode_gpu.py -x -q 2>&1 | tail -30
echo "=== staged epilogue"
timeout 600 python -m pytest tests/test_linear_gpu.py -x -q -k "staged or silu or pair" 2>&1 | tail -5
echo "=== compute-sanitizer (skinny + attention, small)"
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_decode_gpu.py -x -q -k "skinny_multi or argmax or (rope_append and 37)" 2>&1 | tail -8
echo "=== timings"
timeout 900 python tools/decode_dev.py --m 1,32,64 --tunings 0,16,32 2>&1 | tail -40
echo "=== decode bench c3"
timeout 900 python bench.py --workload decode --no-cpu-baseline 2>gpurun_out/r2_dec1_bench.err | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); d.pop('prefill',None); print(json.dumps(d, indent=1))"
tail -5 gpurun_out/r2_dec1_bench.err
} > gpurun_out/r2_dec1.log 2>&1
tail -c 7000 gpurun_out/r2_dec1.log
