#!/bin/bash
# end-of-session validation of HEAD: whole GPU suite, smoke(), default bench line, C4 / C5 prefill lines, attention kernel timing
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -8 > gpurun_out/pytest_final.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_final.log 2>&1
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
timeout 600 python bench.py --workload prefill --prefill-config c4 --prefill-steps 5 > gpurun_out/bench_final_c4.json 2> gpurun_out/bench_final_c4.err
timeout 600 python bench.py --workload prefill --prefill-config c5 --prefill-steps 3 > gpurun_out/bench_final_c5.json 2> gpurun_out/bench_final_c5.err
timeout 300 python tools/bench_attention.py > gpurun_out/bench_att_final.log 2>&1
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err
