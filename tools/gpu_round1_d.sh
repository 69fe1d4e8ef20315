#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_linear_gpu.py tests/test_prefill_gpu.py -q -s 2>&1 | grep -E "max-abs|passed|failed|Error|error" | tail -60 > gpurun_out/pytest_d.log
