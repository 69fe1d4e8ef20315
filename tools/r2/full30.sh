#!/bin/bash
# full GPU suite of HEAD (the previous call stopped at a graph-capture flake, fixed since), then the dense re-merge geometry sweep
mkdir -p gpurun_out
{
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -6
echo "=== dense re-merge (max, negative majority, 3 x 160 M): CTAs per SM, prefetch distance in grids"
for cfg in 4,1 4,0 8,1 8,0 16,0 16,1 2,1 4,2; do echo "remerge=$cfg"; MC_TIES_REMERGE=$cfg timeout 300 python tools/bench_ties.py --func max --kind neg 2>&1 | cut -c90-160; done
} > gpurun_out/r2_full30.log 2>&1
tail -c 4000 gpurun_out/r2_full30.log
