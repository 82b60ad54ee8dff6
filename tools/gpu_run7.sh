#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider --timeout=600 -s > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|rounding noise|identical to the oracle|Error|error" gpurun_out/pytest_gpu.log | tail -30
timeout -s KILL 600 python tools/time_fit.py config2 2 2>&1 | tail -7
timeout -s KILL 900 python tools/time_fit.py config3 2 2>&1 | tail -7
