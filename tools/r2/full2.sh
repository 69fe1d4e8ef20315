#!/bin/bash
# full GPU test suite + smoke + the default bench line (state of HEAD at session start)
mkdir -p gpurun_out
{
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "=== smoke"
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
echo "=== default bench"
SECONDS=0
timeout 1500 python bench.py > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err
echo "rc=$? wall=${SECONDS}s"
tail -5 gpurun_out/r2_bench2.err
echo "=== reference arm"
SECONDS=0
timeout 900 python bench.py --impl reference > gpurun_out/r2_bench2_ref.json 2> gpurun_out/r2_bench2_ref.err
echo "rc=$? wall=${SECONDS}s"
} > gpurun_out/r2_full2.log 2>&1
tail -c 3000 gpurun_out/r2_full2.log
