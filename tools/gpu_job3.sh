#!/bin/bash
set -x
mkdir -p gpurun_out
R=${ROUND_TAG:-r01}
python bench.py --steps 2 --warmup 3 > gpurun_out/bench_${R}.json 2> gpurun_out/bench_${R}.err
tail -c 4000 gpurun_out/bench_${R}.json; tail -5 gpurun_out/bench_${R}.err
bash tools/gpu_ncu.sh
