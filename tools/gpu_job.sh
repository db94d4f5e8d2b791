#!/bin/bash
# ad-hoc GPU job for the current iteration
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
for pv in 0 3; do KF_ATTN_POLY=$pv timeout 300 python tools/gpu_attn.py $( [ $pv = 0 ] && echo "--parity --bwd" ); done
timeout 300 python bench.py --no-extras --steps 30
