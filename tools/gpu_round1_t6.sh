#!/bin/bash
# TIES: max on the fast path, new bracket-width / on-device quotient tests; bench.py stdout = one JSON line
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_ties_gpu.py -q -x --timeout 400 2>&1 | tail -15 > gpurun_out/pytest_t6_ties.log
for args in "--func max --kind neg" "--func max" "--func mean" "--func sum"; do
  timeout 120 python tools/bench_ties.py $args >> gpurun_out/bench_ties_t6.log 2>&1
done
timeout 300 python bench.py --workload ties > gpurun_out/bench_t6_ties.json 2> gpurun_out/bench_t6_ties.err
timeout 300 python -m torch.distributed.run --standalone --local-addr 127.0.0.1 --nproc-per-node 1 bench.py --gpus 1 --workload merge --steps 5 --warmup 3 > gpurun_out/bench_t6_merge.json 2> gpurun_out/bench_t6_merge.err
