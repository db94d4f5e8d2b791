#!/bin/bash
set -u
mkdir -p gpurun_out
tag=s8d
timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "reduce" 2>&1 | tail -4
timeout 300 python tools/gpu_tune_reduce.py --quick 2>&1 | tee gpurun_out/${tag}_tune_reduce.log
