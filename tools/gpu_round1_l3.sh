#!/bin/bash
# ncu --set full of one decoder layer's base + LoRA-up launches on the pair kernel, inside the C3 prefill (second step)
set -x
mkdir -p gpurun_out
timeout 700 ncu --set full --clock-control none --import-source on -k regex:linear2_kernel -s 232 -c 6 -o gpurun_out/prefill_linear2_full -f \
    python bench.py --workload prefill --prefill-steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_l3.log 2>&1
