#!/bin/bash
# guarded: a quick loop test first under a short timeout; stop at once if it hangs or fails
mkdir -p gpurun_out
timeout -s KILL 200 python -m pytest tests/test_gpu_loop.py -m gpu -q --no-header -p no:cacheprovider -x -k "tiny" > gpurun_out/pytest_quick.log 2>&1
rc=$?
tail -3 gpurun_out/pytest_quick.log
if [ $rc -ne 0 ]; then echo "QUICK TEST FAILED rc=$rc -- stopping"; exit 1; fi
timeout -s KILL 400 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider --timeout=200 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|Error|error|assert" gpurun_out/pytest_gpu.log | tail -15
if [ "$1" == "--tests-only" ]; then exit 0; fi
timeout -s KILL 200 python tools/time_fit.py config3 1 2>&1 | grep -E "rep|gemm" | cut -c1-330
timeout -s KILL 400 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_config3_n1.json 2> gpurun_out/bench_config3_n1.err
echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench_config3_n1.json')); print({k:d[k] for k in ('value','gpu_launches','assignment_checksum')}, d['e2e']['value'], {k:round(v['ms_per_step']) for k,v in d['kernel_classes_rank0'].items()})"
