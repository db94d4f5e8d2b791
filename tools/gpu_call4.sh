#!/bin/bash
set -u
mkdir -p gpurun_out
tag=s8k
timeout 300 python -m pytest tests/test_attention_gpu.py tests/test_block_gpu.py -m gpu -x -q 2>&1 | tail -3
for cfg in "1 2" "1 0" "1 3"; do
  set -- $cfg
  echo "SPLIT=$1 POLY=$2"
  KF_ATTN_SPLIT=$1 KF_ATTN_POLY=$2 timeout 120 python tools/gpu_attn.py --parity 2>&1 | tail -5
done | tee gpurun_out/${tag}_attn.log
