#!/bin/bash
# end-of-round validation of HEAD: whole GPU suite, smoke(), default bench line, C4 / C5 prefill lines, reference arm
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -8 > gpurun_out/pytest_final4.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_final4.log 2>&1
timeout 900 python bench.py > gpurun_out/bench_final4.json 2> gpurun_out/bench_final4.err
