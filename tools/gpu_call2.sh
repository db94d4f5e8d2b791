#!/bin/bash
set -u
mkdir -p gpurun_out
tag=s5b
timeout 300 python tools/gpu_tune_reduce.py > gpurun_out/${tag}_tune_reduce.log 2>&1; echo "tune rc=$?"
cat gpurun_out/${tag}_tune_reduce.log
timeout 200 python -m pytest tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -3
bash tools/ncu_capture.sh $tag gemm attn attn_bwd > gpurun_out/${tag}_ncu.log 2>&1
tail -12 gpurun_out/${tag}_ncu.log
