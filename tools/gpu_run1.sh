#!/bin/bash
# first GPU contact: kernel parity tests + GEMM probe
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi.txt 2>&1
timeout -s KILL 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider --timeout=600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout -s KILL 600 python tools/gemm_probe.py > gpurun_out/gemm_probe.log 2>&1
echo "probe exit $?" >> gpurun_out/gemm_probe.log
tail -40 gpurun_out/pytest_gpu.log; cat gpurun_out/gemm_probe.log
