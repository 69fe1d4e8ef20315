#!/bin/bash
# RoPE epilogue of the q/k/v launch with 32-byte cos / sin loads and output stores (new.so) against 16-byte ones (old.so), inside the C3 prefill,
# alternating on one box; then parity of the new library (linear + prefill + decode tests)
mkdir -p gpurun_out
L=modelcompose_b200/_lib
run() { # label, lib
  cp $L/ab/$2.so $L/libmodelcompose_b200.so
  timeout 600 python bench.py --workload prefill --prefill-config c3 --no-cpu-baseline 2>/dev/null | tail -1 | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$1', d['value'], 'tok/s', d['ms_per_step'], 'ms  linears', r['kernel_ms_per_step'], 'ms frac', r['frac'], 'clk', d['clocks']['sm_mhz'])"
}
{
for rep in 1 2 3; do
run "16-byte (old)" old
run "32-byte (new)" new
done
cp $L/ab/new.so $L/libmodelcompose_b200.so
timeout 900 python -m pytest tests/test_linear_gpu.py tests/test_prefill_gpu.py tests/test_decode_gpu.py -q --timeout 300 2>&1 | tail -4
} > gpurun_out/r2_rope32.log 2>&1
cat gpurun_out/r2_rope32.log
