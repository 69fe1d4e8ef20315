#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_linear_gpu.py tests/test_prefill_gpu.py -q -x 2>&1 | tail -15 > gpurun_out/pytest_h.log
timeout 180 python tools/debug_linear.py --two-cta > gpurun_out/debug_linear_h.log 2>&1; echo "rc=$?" >> gpurun_out/debug_linear_h.log
nvidia-smi --query-gpu=name,memory.used --format=csv >> gpurun_out/debug_linear_h.log 2>&1
