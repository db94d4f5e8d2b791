#!/bin/bash
# bring-up visit for the fused attention backward / strided attention / chunked fp32 GEMM
set -u
tag=$1
mkdir -p gpurun_out
timeout 300 python tools/gpu_attn.py --parity --bwd > gpurun_out/${tag}_attn.log 2>&1; echo "attn rc=$?"; tail -12 gpurun_out/${tag}_attn.log
KF_ATTN_BWD_TWO_KERNEL=1 timeout 300 python tools/gpu_attn.py --bwd > gpurun_out/${tag}_attn_two.log 2>&1; echo "attn2 rc=$?"; tail -4 gpurun_out/${tag}_attn_two.log
timeout 1200 python -m pytest tests/test_attention_gpu.py tests/test_round2_ops_gpu.py tests/test_baseline_shapes_gpu.py tests/test_block_gpu.py -m gpu -q --durations=10 > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"
tail -40 gpurun_out/${tag}_pytest.log
