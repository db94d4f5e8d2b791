#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py -m gpu -x -q -k "gemm_host" 2>&1 | tail -3
timeout 300 python bench.py --no-extras > gpurun_out/s8p_bench.json 2> gpurun_out/s8p_bench.err; tail -2 gpurun_out/s8p_bench.err; cat gpurun_out/s8p_bench.json
