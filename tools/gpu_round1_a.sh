#!/bin/bash
# First GPU pass of the round: GPU parity tests, smoke, bench, ncu launch list + full capture of the merge kernel.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks_a.csv &
SMI=$!
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu_a.log
python __graft_entry__.py smoke > gpurun_out/smoke_a.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_a.json 2> gpurun_out/bench_ref_a.err
kill $SMI
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_merge_a.csv \
    python bench.py --steps 3 --warmup 3 --no-e2e > gpurun_out/ncu_launches_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:merge_kernel -s 3 -c 2 -o gpurun_out/prof_merge_a -f \
    python bench.py --steps 3 --warmup 3 --no-e2e --emulate-world 8 > gpurun_out/ncu_full_a.log 2>&1
ls -la gpurun_out
