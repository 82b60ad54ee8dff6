#!/bin/bash
# tests + timings, then one ncu --set full capture of the vote / NNLS / update kernels
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider --timeout=600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|Error|error|assert" gpurun_out/pytest_gpu.log | tail -10
timeout -s KILL 900 python tools/time_fit.py config3 1 2>&1 | tail -4
timeout -s KILL 900 python tools/time_fit.py config4 1 2>&1 | tail -4
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:"resid_vote_kernel|nnls_kernel|admm_update_kernel" -s 12 -c 6 -o gpurun_out/prof_rv \
    python tools/time_fit.py config3 1 > gpurun_out/ncu_rv.log 2>&1
echo "ncu exit $?"
