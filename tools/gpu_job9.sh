#!/bin/bash
set -x
mkdir -p gpurun_out
R=${ROUND_TAG:-r01}
python -m pytest tests -m gpu -q 2>&1 | tail -8
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 2 --warmup 3 > gpurun_out/bench_${R}.json 2> gpurun_out/bench_${R}.err
tail -c 1800 gpurun_out/bench_${R}.json; tail -5 gpurun_out/bench_${R}.err
timeout 200 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_${R}.json 2> gpurun_out/bench_ref_${R}.err
tail -c 600 gpurun_out/bench_ref_${R}.json
