#!/bin/bash
# TIES: finer sweep of the merge pass' prefetch distance on both data sets, parity, the dense re-merge case
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_ties_gpu.py -x -q --timeout 600 2>&1 | tail -2
echo "=== prefetch sweep, bench_ties mean 320M / bench.py ties"
for pf in 74 111 148 185 222 296; do
  echo "pf=$pf"
  MC_TIES_PREFETCH=$pf timeout 300 python tools/bench_ties.py --func mean --elements 320e6 2>&1 | cut -c90-150
  MC_TIES_PREFETCH=$pf timeout 600 python bench.py --workload ties 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('  bench.py', d['value'], d['ms_per_step'], d['roofline']['frac'])"
done
echo "=== defaults"
for f in mean sum; do timeout 300 python tools/bench_ties.py --func $f 2>&1 | cut -c1-150; done
timeout 300 python tools/bench_ties.py --func max --kind neg 2>&1 | cut -c1-150
timeout 300 python tools/bench_ties.py --func mean --dtype f16 2>&1 | cut -c1-150
timeout 300 python tools/bench_ties.py --func mean --src 8 --elements 80e6 2>&1 | cut -c1-150
timeout 300 python tools/bench_ties.py --func sum --src 4 --elements 320e6 2>&1 | cut -c1-150
} > gpurun_out/r2_ties27.log 2>&1
tail -c 5000 gpurun_out/r2_ties27.log
