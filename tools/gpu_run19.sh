#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider --timeout=600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|Error|error|assert" gpurun_out/pytest_gpu.log | tail -15
timeout -s KILL 900 python tools/time_fit.py config3 2 2>&1 | grep -E "rep|gemm" | cut -c1-330
timeout -s KILL 900 python bench.py --steps 1 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','gpu_launches','assignment_checksum')}, d['e2e']['value'])"
