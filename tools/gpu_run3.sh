#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do python tools/debug_rv.py small 3 0.0 2>&1 | grep -E "rep|e-0[0-9]"; done
timeout -s KILL 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider --timeout=600 -s > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|rounding noise|identical to the oracle|Error|error" gpurun_out/pytest_gpu.log | tail -40
timeout -s KILL 600 python tools/gemm_probe.py > gpurun_out/gemm_probe.log 2>&1; cat gpurun_out/gemm_probe.log
