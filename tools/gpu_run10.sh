#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider --timeout=600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|Error|error|assert" gpurun_out/pytest_gpu.log | tail -15
timeout -s KILL 600 python tools/microbench.py --quick 2>&1 | tail -12
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:"resid_vote_kernel|nnls_kernel" -c 4 -o gpurun_out/prof_rv2 \
    python tools/time_fit.py config3 1 > gpurun_out/ncu_rv2.log 2>&1
echo "ncu exit $?"
