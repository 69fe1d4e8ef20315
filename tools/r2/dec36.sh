#!/bin/bash
# branch-form decode step: the four low-rank "down" launches per layer on the register kernel (MC_DECODE_DOWN_TUNING=16) against the stream-K kernel, alternating
mkdir -p gpurun_out
run() { # label, env...
  local label=$1; shift
  env "$@" timeout 600 python bench.py --workload decode --no-cpu-baseline 2>/dev/null | tail -1 | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$label', d['value'], 'tok/s', d['ms_per_step'], 'ms frac', r['frac'], 'launches/step', d.get('launches_per_step'), 'clk', (d.get('clocks') or {}).get('sm_mhz'))"
}
{
for rep in 1 2 3; do
run "down: stream-K " MC_X=1
run "down: register " MC_DECODE_DOWN_TUNING=16
done
MC_DECODE_DOWN_TUNING=16 timeout 600 python -m pytest tests/test_decode_gpu.py -q --timeout 300 -k "decode_steps_vs_prefill or graph_equals or full_width or generate" 2>&1 | tail -2
} > gpurun_out/r2_dec36.log 2>&1
cat gpurun_out/r2_dec36.log
