#!/bin/bash
# 2-GPU validation of the default bench line (merge strong scaling + nested C3 prefill weak scaling), launched as the driver does
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/n2b_gpus.txt 2>&1
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 3 --prefill-steps 3 > gpurun_out/bench_n2b.json 2> gpurun_out/bench_n2b.err
tail -5 gpurun_out/bench_n2b.err
