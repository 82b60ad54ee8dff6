#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider --timeout=600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|Error|error|assert" gpurun_out/pytest_gpu.log | tail -15
timeout -s KILL 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_config3_n1.json 2> gpurun_out/bench_config3_n1.err
echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench_config3_n1.json')); print({k:d[k] for k in ('value','e2e','gpu_launches') if k in d}); print(d.get('assignment_checksum'), {k:round(v['ms_per_step']) for k,v in d['kernel_classes_rank0'].items()})"
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tn_kernel -s 40 -c 2 -o gpurun_out/prof_admm_gemm_tall \
    python tools/time_fit.py config3 1 > gpurun_out/ncu_gemm_tall.log 2>&1
echo "ncu exit $?"
