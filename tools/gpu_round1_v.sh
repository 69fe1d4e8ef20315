#!/bin/bash
# re-entry sanity: GPU parity tests, default bench line, launch list of the fused prefill step
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/pytest_v.log
timeout 900 python bench.py > gpurun_out/bench_v.json 2> gpurun_out/bench_v.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_v.csv python bench.py --workload prefill --prefill-steps 1 --warmup 1 --no-e2e > gpurun_out/ncu_launch_v.log 2>&1
