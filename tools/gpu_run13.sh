#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider --timeout=600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|Error|error|assert" gpurun_out/pytest_gpu.log | tail -15
timeout -s KILL 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_config3_n1.json 2> gpurun_out/bench_config3_n1.err
echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench_config3_n1.json')); print({k:d[k] for k in ('value','e2e','gpu_launches') if k in d}); print(d.get('assignment_checksum'))"
