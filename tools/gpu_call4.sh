#!/bin/bash
set -u
mkdir -p gpurun_out
tag=s8f
timeout 600 python -m pytest tests/test_attention_gpu.py -m gpu -x -q 2>&1 | tail -4
for cfg in "2 0" "2 3" "2 2" "1 0" "1 3"; do
  set -- $cfg
  echo "SPLIT=$1 POLY=$2"
  KF_ATTN_SPLIT=$1 KF_ATTN_POLY=$2 timeout 120 python tools/gpu_attn.py 2>&1 | tail -1
done | tee gpurun_out/${tag}_attn_split.log
