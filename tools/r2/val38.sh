#!/bin/bash
# last validation of HEAD: full GPU suite + smoke (the default bench of this state differs from tools/r2/final34.sh's only by the decode_dense option, off by default)
mkdir -p gpurun_out
{
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -4
echo "=== smoke"
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "=== quick bench (merge + ties lines only)"
timeout 600 python bench.py --workload ties 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['metric'], d['value'], d['roofline']['frac'], d['e2e']['value'])"
} > gpurun_out/r2_val38.log 2>&1
cat gpurun_out/r2_val38.log
