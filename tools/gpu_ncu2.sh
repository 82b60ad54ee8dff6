#!/bin/bash
mkdir -p gpurun_out
python tools/time_fit.py config3 1 > gpurun_out/plain_c3b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"resid_vote_kernel|admm_update_kernel|nnls_kernel|admm_rhs" -s 30 -c 8 -o gpurun_out/prof_other \
    python tools/time_fit.py config3 1 > gpurun_out/ncu_c3b.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/plain_c3b.log
