#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_prefill_gpu.py tests/test_linear_gpu.py -q -x -s 2>&1 | grep -E "max-abs|passed|failed|Error|error|assert" | tail -40 > gpurun_out/pytest_r.log
