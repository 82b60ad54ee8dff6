#!/bin/bash
# One `ncu --set full` capture of the kernels matching $1 (regex), skipping $2 launches, capturing $3;
# run only after the same command has exited 0 without the profiler.  Summarise with tools/ncu_summary.py.
#   gpurun --timeout 1200 -- 'bash tools/gpu_ncu.sh "gemm_tn_kernel" 40 2 admm_gemm'
mkdir -p gpurun_out
K=${1:-gemm_tn_kernel}; S=${2:-0}; C=${3:-2}; NAME=${4:-capture}
timeout -s KILL 600 python tools/time_fit.py config3 1 > gpurun_out/plain_$NAME.log 2>&1 || { echo "plain run failed"; exit 1; }
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:"$K" -s $S -c $C -o gpurun_out/prof_$NAME \
    python tools/time_fit.py config3 1 > gpurun_out/ncu_$NAME.log 2>&1
echo "ncu exit $?"
