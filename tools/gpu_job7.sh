#!/bin/bash
set -x
mkdir -p gpurun_out
R=${ROUND_TAG:-r01}
python -m pytest tests -m gpu -q 2>&1 | tail -12
B200_EMPANADA_PROFILE=1 python tools/profile_pipeline.py 1024 16 > gpurun_out/phases_${R}.txt 2>&1
tail -3 gpurun_out/phases_${R}.txt
python bench.py --steps 2 --warmup 3 > gpurun_out/bench_${R}.json 2> gpurun_out/bench_${R}.err
tail -c 2500 gpurun_out/bench_${R}.json; tail -5 gpurun_out/bench_${R}.err
