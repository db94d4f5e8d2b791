#!/bin/bash
# final-numbers visit: full bench, then an ncu launch list of one transformer-block step
set -u
tag=$1
mkdir -p gpurun_out
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${tag}_block_launches.csv python tools/gpu_block_probe.py --once > gpurun_out/${tag}_block_probe.log 2>&1; echo "ncu rc=$?"
timeout 300 python -m pytest tests/test_block_gpu.py tests/test_round2_ops_gpu.py tests/test_parity_gpu.py -m gpu -q -x 2>&1 | tail -3
