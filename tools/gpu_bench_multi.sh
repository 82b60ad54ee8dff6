#!/bin/bash
# tools/gpu_bench_multi.sh N [tag] [--tests]: the bench line of config 3 on N GPUs of one box (torchrun, one rank per GPU);
# with --tests also the multi-GPU parity tests (two GPUs == one GPU bit for bit).
mkdir -p gpurun_out
N=${1:-2}; tag=${2:-r02}
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/${tag}_bench_config3_n$N.json 2> gpurun_out/${tag}_bench_config3_n$N.err
echo "bench N=$N exit $?"
python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_bench_config3_n$N.json"))
print($N, round(d["value"],4), round(d["e2e"]["value"],4), "share", round(d["kernel_share_of_step"],3), "frac", round(d["roofline"]["frac"],3),
      {k:round(v["ms_per_step"]) for k,v in d["kernel_classes_rank0"].items()}, d.get("assignment_checksum"), d["profiled_pass"]["ms_per_step"])
PY
tail -3 gpurun_out/${tag}_bench_config3_n$N.err
if [ "$3" == "--tests" ]; then
  timeout -s KILL 600 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider --timeout=300 -k "gpus or multi or two_gpu or sharded" > gpurun_out/${tag}_pytest_multigpu.log 2>&1
  echo "pytest exit $?"; tail -3 gpurun_out/${tag}_pytest_multigpu.log
fi
