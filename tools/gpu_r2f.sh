#!/bin/bash
set -u
tag=$1
mkdir -p gpurun_out
for cfg in "1 2" "0 2" "1 3" "1 0" "1 4"; do
  set -- $cfg
  echo "== KF_ATTN_STALE=$1 KF_ATTN_POLY=$2"
  KF_ATTN_STALE=$1 KF_ATTN_POLY=$2 timeout 300 python tools/gpu_attn.py --parity 2>&1 | grep -E "out max|attn fwd|rror|imed out" | cut -c1-150
done
timeout 600 python -m pytest tests/test_attention_gpu.py tests/test_baseline_shapes_gpu.py tests/test_block_gpu.py -m gpu -q -x 2>&1 | tail -3
