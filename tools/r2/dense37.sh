#!/bin/bash
# decode_dense: branch-form prefill + dense text-group weights for the decode steps; tests, then bench.py --workload decode with and without it
mkdir -p gpurun_out
run() { # label, env...
  local label=$1; shift
  env "$@" timeout 600 python bench.py --workload decode --no-cpu-baseline 2>gpurun_out/r2_dense37.err | tail -1 | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$label', d['value'], 'tok/s', d['ms_per_step'], 'ms frac', r['frac'], 'launches/step', d.get('launches_per_step'), 'e2e', d['e2e']['value'], 'form:', d['config'].get('linear_form'), 'parity', json.dumps(d.get('parity'))[:200])"
  tail -2 gpurun_out/r2_dense37.err
}
{
timeout 900 python -m pytest tests/test_decode_gpu.py tests/test_prefill_gpu.py -q --timeout 300 2>&1 | tail -3
run "branch decode      " MC_X=1
run "decode_dense       " MC_DECODE_DENSE=1
run "branch decode      " MC_X=1
run "decode_dense       " MC_DECODE_DENSE=1
} > gpurun_out/r2_dense37.log 2>&1
cat gpurun_out/r2_dense37.log
