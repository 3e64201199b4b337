#!/bin/bash
set -x
mkdir -p gpurun_out
R=${ROUND_TAG:-r01}
timeout 120 ./tools/test_conv_gemm perf 6 all 2>&1 | grep -E "ok|FAIL|TOTAL|l1 3x3" | tail -30
python -m pytest tests -m gpu -q 2>&1 | tail -6
python tools/profile_forward.py 16 1024 > gpurun_out/prof_fwd_${R}.txt 2>&1
head -8 gpurun_out/prof_fwd_${R}.txt; grep -n "layer1.*c2" gpurun_out/prof_fwd_${R}.txt
