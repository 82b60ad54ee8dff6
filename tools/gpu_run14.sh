#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider --timeout=600 -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|Error|error|assert" gpurun_out/pytest_gpu.log | tail -15
timeout -s KILL 900 python tools/time_fit.py config3 2 --r2 2>&1 | tail -8
timeout -s KILL 900 python tools/time_fit.py config4 1 2>&1 | tail -4
