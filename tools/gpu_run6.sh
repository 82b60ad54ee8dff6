#!/bin/bash
# first bench lines + ncu evidence
mkdir -p gpurun_out
timeout -s KILL 900 python bench.py --workload config3 --steps 2 --warmup 3 > gpurun_out/bench_config3.json 2> gpurun_out/bench_config3.err
echo "bench config3 exit $?"; tail -c 3000 gpurun_out/bench_config3.json; tail -5 gpurun_out/bench_config3.err
timeout -s KILL 600 python bench.py --workload config2 --steps 3 --warmup 3 > gpurun_out/bench_config2.json 2> gpurun_out/bench_config2.err
echo "bench config2 exit $?"; tail -c 2500 gpurun_out/bench_config2.json
# ncu launch list of a short bench command (cold-cache, serialised: compare SHARES)
python bench.py --workload config2 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain_c2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 3000 --csv --log-file gpurun_out/launches_config2.csv \
    python bench.py --workload config2 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c2.log 2>&1
echo "ncu list exit $?"; wc -l gpurun_out/launches_config2.csv
# full-set capture of the top kernel at config 3
python tools/time_fit.py config3 1 > gpurun_out/plain_c3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_tn_kernel -s 40 -c 3 -o gpurun_out/prof_admm_gemm \
    python tools/time_fit.py config3 1 > gpurun_out/ncu_c3.log 2>&1
echo "ncu full exit $?"; ls -la gpurun_out/
