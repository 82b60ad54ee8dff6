#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:"resid_vote_kernel" -c 3 -o gpurun_out/prof_rv3 \
    python tools/time_fit.py config3 1 > gpurun_out/ncu_rv3.log 2>&1
echo "ncu exit $?"
