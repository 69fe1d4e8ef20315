#!/bin/bash
# staged (coalesced) epilogue vs row-per-thread epilogue inside the C3 prefill, alternating on one box; hybrid vs pair-everywhere
mkdir -p gpurun_out
run() { # label, env...
  local label=$1; shift
  env "$@" timeout 600 python bench.py --workload prefill --prefill-config c3 --no-cpu-baseline 2>/dev/null | tail -1 | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$label', d['value'], 'tok/s', d['ms_per_step'], 'ms  linears', r['kernel_ms_per_step'], 'ms frac', r['frac'], 'clk', d['clocks']['sm_mhz'])"
}
{
for rep in 1 2; do
run "staged hybrid      " MC_X=1
run "rowwise hybrid     " MC_LINEAR_EPI_ROWWISE=1
run "staged pair-all    " MC_LINEAR_UP_TUNING=3
run "rowwise pair-all   " MC_LINEAR_UP_TUNING=3 MC_LINEAR_EPI_ROWWISE=1
run "staged single-all  " MC_LINEAR_UP_TUNING=0
done
run "staged pair-all materialised " MC_LINEAR_UP_TUNING=3 MC_MATERIALIZE=1
run "staged hybrid materialised   " MC_MATERIALIZE=1
} > gpurun_out/r2_epi1.log 2>&1
cat gpurun_out/r2_epi1.log
