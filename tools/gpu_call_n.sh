#!/bin/bash
# multi-GPU visit: tools/gpu_call_n.sh <N> <tag>
set -u
N=$1; tag=$2
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv | head -10
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 300 $TR tools/gpu_dist_check.py 2>&1 | grep -v "^W\|^\*\*\*" | tail -12 | tee gpurun_out/${tag}_dist_check.log
timeout 600 $TR bench.py --gpus $N --workload block --steps 5 --warmup 3 2> gpurun_out/${tag}_block.err | tee gpurun_out/${tag}_block_n${N}.json
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 2> gpurun_out/${tag}_gemm.err | tee gpurun_out/${tag}_bench_n${N}.json
tail -3 gpurun_out/${tag}_block.err gpurun_out/${tag}_gemm.err
