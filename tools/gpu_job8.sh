#!/bin/bash
set -x
mkdir -p gpurun_out
R=${ROUND_TAG:-r01}
python -m pytest tests -m gpu -q 2>&1 | tail -6
python tools/profile_forward.py 16 1024 > gpurun_out/prof_fwd_${R}.txt 2>&1
head -6 gpurun_out/prof_fwd_${R}.txt; grep -n "dwconv\|bilinear" gpurun_out/prof_fwd_${R}.txt
B200_EMPANADA_FUSED_UPSAMPLE=0 python tools/profile_forward.py 16 1024 > gpurun_out/prof_fwd_unfused_${R}.txt 2>&1
head -6 gpurun_out/prof_fwd_unfused_${R}.txt; grep -n "dwconv\|bilinear\|\.low " gpurun_out/prof_fwd_unfused_${R}.txt
