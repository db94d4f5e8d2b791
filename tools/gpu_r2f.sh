#!/bin/bash
set -u
tag=$1
mkdir -p gpurun_out
for cfg in "ps 2" "ps 0" "ps 3" "ps 4" "tc 2"; do
  set -- $cfg
  echo "== KF_ATTN_FWD=$1 KF_ATTN_POLY=$2"
  KF_ATTN_FWD=$1 KF_ATTN_POLY=$2 timeout 300 python tools/gpu_attn.py --parity 2>&1 | grep -E "out max|attn fwd|rror|imed out" | cut -c1-150
done
KF_ATTN_FWD=ps timeout 600 python -m pytest tests/test_attention_gpu.py tests/test_baseline_shapes_gpu.py -m gpu -q -x 2>&1 | tail -3
