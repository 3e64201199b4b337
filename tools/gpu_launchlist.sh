#!/bin/bash
# ncu launch list of the bench command (256^3 so that the whole step fits in the time limit).
set -x
mkdir -p gpurun_out
R=${ROUND_TAG:-r01}
BENCH_CUDA_PROFILER_API=1 timeout 1100 ncu --clock-control none --profile-from-start off --metrics gpu__time_duration.sum --csv \
    --log-file gpurun_out/launches_${R}.csv python bench.py --size 256 --steps 1 --warmup 1 --no-cpu --no-2d > gpurun_out/ncu_bench_${R}.log 2>&1
tail -2 gpurun_out/ncu_bench_${R}.log | cut -c1-300
wc -l gpurun_out/launches_${R}.csv
