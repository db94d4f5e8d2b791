#!/bin/bash
set -u
tag=$1
mkdir -p gpurun_out
KF_ATTN_BWD=wide timeout 300 python tools/gpu_attn.py --parity --bwd 2>&1 | grep -E "max_|attn " 
KF_ATTN_BWD=wide KF_ATTN_TRACE=1 timeout 300 python tools/gpu_attn.py --bwd > gpurun_out/${tag}_trace.log 2>&1; grep -B2 -A5 "absolute" gpurun_out/${tag}_trace.log | head -24 | cut -c1-190
timeout 600 python -m pytest tests/test_attention_gpu.py tests/test_baseline_shapes_gpu.py tests/test_block_gpu.py -m gpu -q -x 2>&1 | tail -3
timeout 300 python tools/gpu_block_probe.py 2>&1 | tail -3
