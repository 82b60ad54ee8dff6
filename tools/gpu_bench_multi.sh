#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 2 --warmup 2 > gpurun_out/bench_config3_n$N.json 2> gpurun_out/bench_config3_n$N.err
echo "bench N=$N exit $?"; tail -c 2500 gpurun_out/bench_config3_n$N.json; tail -5 gpurun_out/bench_config3_n$N.err
