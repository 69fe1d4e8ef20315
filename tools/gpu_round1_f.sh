#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_f.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2_f.json 2> gpurun_out/bench_n2_f.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2_f.json 2> gpurun_out/bench_ref_n2_f.err
tail -5 gpurun_out/bench_n2_f.err
