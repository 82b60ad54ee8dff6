#!/bin/bash
# One-GPU measurement pass of a round: parity tests, bench lines of the three BASELINE workloads, the reference
# (CPU) arm, k-means++ timing and one ncu capture of the vote kernel.  Everything lands in gpurun_out/.
#   tools/gpu_measure.sh [tag] [lite]   (run through gpurun on a B200 box; lite: no config 4, k-means++ timing, ncu)
tag=${1:-r02}; lite=${2:-full}
mkdir -p gpurun_out
timeout -s KILL 200 python -m pytest tests/test_gpu_loop.py -m gpu -q --no-header -p no:cacheprovider -x -k "tiny" > gpurun_out/${tag}_pytest_quick.log 2>&1
rc=$?; tail -2 gpurun_out/${tag}_pytest_quick.log
if [ $rc -ne 0 ]; then echo "QUICK TEST FAILED rc=$rc -- stopping"; exit 1; fi
timeout -s KILL 600 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider --timeout=300 > gpurun_out/${tag}_pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -2 gpurun_out/${tag}_pytest_gpu.log
timeout -s KILL 600 python bench.py --steps 3 --warmup 3 > gpurun_out/${tag}_bench_config3_n1.json 2> gpurun_out/${tag}_bench_config3_n1.err; echo "bench c3 exit $?"
timeout -s KILL 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${tag}_bench_config3_reference.json 2> gpurun_out/${tag}_bench_config3_reference.err; echo "bench ref exit $?"
timeout -s KILL 300 python bench.py --workload config2 --steps 5 --warmup 3 > gpurun_out/${tag}_bench_config2_n1.json 2> gpurun_out/${tag}_bench_config2_n1.err; echo "bench c2 exit $?"
if [ "$lite" != "lite" ]; then
timeout -s KILL 600 python bench.py --workload config4 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_config4_n1.json 2> gpurun_out/${tag}_bench_config4_n1.err; echo "bench c4 exit $?"
timeout -s KILL 200 python tools/time_kmeanspp.py config3 > gpurun_out/${tag}_kmeanspp_config3.log 2>&1; tail -3 gpurun_out/${tag}_kmeanspp_config3.log
timeout -s KILL 200 python tools/time_cycle.py config3 > gpurun_out/${tag}_cycle_config3.log 2>&1 && \
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:resid_vote_kernel -c 1 -f -o gpurun_out/${tag}_prof_vote_ship python tools/time_cycle.py config3 > gpurun_out/${tag}_ncu_vote.log 2>&1; echo "ncu exit $?"
cat gpurun_out/${tag}_cycle_config3.log | cut -c1-300
fi
python - <<PY
import json
for w in ("config3","config2","config4"):
    try:
        d=json.load(open("gpurun_out/${tag}_bench_%s_n1.json" % w))
        print(w, round(d["value"],4), round(d["e2e"]["value"],4), "share", round(d["kernel_share_of_step"],3), "frac", round(d["roofline"]["frac"],3),
              {k:round(v["ms_per_step"]) for k,v in d["kernel_classes_rank0"].items()}, d.get("assignment_checksum"), d.get("parity_spot",{}).get("iterations_equal"), (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e: print(w, "unreadable", e)
PY
