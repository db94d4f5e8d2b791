#!/bin/bash
set -u
tag=$1
mkdir -p gpurun_out
KF_ATTN_TRACE=1 timeout 300 python tools/gpu_attn.py --bwd > gpurun_out/${tag}_trace.log 2>&1; echo "trace rc=$?"; grep -B3 -A8 "absolute" gpurun_out/${tag}_trace.log | head -44 | cut -c1-200
timeout 300 python tools/gpu_attn.py --parity --bwd 2>&1 | tail -8
timeout 600 python -m pytest tests/test_attention_gpu.py tests/test_baseline_shapes_gpu.py tests/test_block_gpu.py -m gpu -q -x 2>&1 | tail -5
