#!/bin/bash
# ad-hoc GPU job for the current iteration
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_attention_gpu.py tests/test_block_gpu.py -q -x 2>&1 | tail -15
for pv in 0 2 3 4; do KF_ATTN_POLY=$pv timeout 300 python tools/gpu_attn.py $( [ $pv = 3 ] && echo "--parity --bwd" ); done
