#!/bin/bash
# 4-GPU validation of the default bench line (merge strong scaling + nested C3 prefill weak scaling), launched as the driver does
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/n4_gpus.txt 2>&1
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 20 --warmup 3 --prefill-steps 3 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err
tail -5 gpurun_out/bench_n4.err
