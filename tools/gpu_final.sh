#!/bin/bash
# Round-end evidence run (one B200): GPU tests, smoke, conv DRAM-traffic capture of the bench's
# launch list, the bench line and the reference arm.
set -x
mkdir -p gpurun_out
R=${ROUND_TAG:-r01}
python -m pytest tests -m gpu -q 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active"
timeout 400 ncu --clock-control none --profile-from-start off --metrics $M --csv --log-file gpurun_out/fwd_metrics_${R}.csv python tools/ncu_forward.py 37 1024 > gpurun_out/ncu_fwd_${R}.log 2>&1
tail -2 gpurun_out/ncu_fwd_${R}.log | cut -c1-200
python bench.py --steps 2 --warmup 3 > gpurun_out/bench_${R}.json 2> gpurun_out/bench_${R}.err
tail -c 1200 gpurun_out/bench_${R}.json; tail -5 gpurun_out/bench_${R}.err
