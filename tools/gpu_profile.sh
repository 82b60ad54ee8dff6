#!/bin/bash
# Round-end evidence: bench line (no profiler), then the ncu launch list of the same command shape.
mkdir -p gpurun_out
timeout -s KILL 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_config3_n1.json 2> gpurun_out/bench_config3_n1.err
echo "bench exit $?"
timeout -s KILL 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 14000 --csv --log-file gpurun_out/launches_config3.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
echo "ncu exit $?"
python tools/summarize_launches.py gpurun_out/launches_config3.csv "ncu --metrics gpu__time_duration.sum --clock-control none -c 14000 : python bench.py --steps 1 --warmup 1 --no-cpu-baseline  (config3, B200)" > gpurun_out/launches_config3_summary.txt
head -30 gpurun_out/launches_config3_summary.txt
