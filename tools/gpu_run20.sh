#!/bin/bash
SCRC_BENCH_DEBUG=1 timeout -s KILL 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "bench step"
