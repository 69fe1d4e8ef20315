#!/bin/bash
# 2-GPU runs exactly as the driver launches them: default line (merge sharded by tensor + C3 prefill sharded by request batch,
# NCCL all-gather of the probe request's logits for verification), the reference arm, and the C4 prefill
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
timeout 600 $TR bench.py --gpus 2 --steps 2 --warmup 1 --impl reference > gpurun_out/bench_n2_ref.json 2> gpurun_out/bench_n2_ref.err
timeout 900 $TR bench.py --gpus 2 --workload prefill --prefill-config c4 --prefill-steps 5 > gpurun_out/bench_n2_c4.json 2> gpurun_out/bench_n2_c4.err
