#!/bin/bash
# tile rasterisation of the base launches inside the C3 prefill: M tiles per sweep over N (library default: pair kernel 2 for q/k/v, 8 for gate / down;
# single-CTA kernel 8 / 32), alternating on one box
mkdir -p gpurun_out
run() { # label, env...
  local label=$1; shift
  env "$@" timeout 600 python bench.py --workload prefill --prefill-config c3 --no-cpu-baseline 2>/dev/null | tail -1 | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$label', d['value'], 'tok/s', d['ms_per_step'], 'ms  linears', r['kernel_ms_per_step'], 'ms frac', r['frac'], 'clk', d['clocks']['sm_mhz'])"
}
{
for rep in 1 2; do
run "default          " MC_X=1
run "pair 1           " MC_LINEAR_GROUP_M_PAIR=1
run "pair 2           " MC_LINEAR_GROUP_M_PAIR=2
run "pair 4           " MC_LINEAR_GROUP_M_PAIR=4
run "pair 6           " MC_LINEAR_GROUP_M_PAIR=6
run "pair 16          " MC_LINEAR_GROUP_M_PAIR=16
run "single 8         " MC_LINEAR_GROUP_M_SINGLE=8
run "single 16        " MC_LINEAR_GROUP_M_SINGLE=16
run "single 64        " MC_LINEAR_GROUP_M_SINGLE=64
done
} > gpurun_out/r2_raster33.log 2>&1
cat gpurun_out/r2_raster33.log
