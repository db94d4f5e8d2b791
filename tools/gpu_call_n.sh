#!/bin/bash
# multi-GPU visit: tools/gpu_call_n.sh <N> <tag> [lean]
set -u
N=$1; tag=$2; lean=${3:-}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 300 $TR tools/gpu_dist_check.py 2>&1 | grep -v "^W\|^\*\*\*" | tail -12 | tee gpurun_out/${tag}_dist_check.log
timeout 600 $TR tools/gpu_dist_check.py 4096 4096 2>&1 | grep -v "^W\|^\*\*\*" | tail -12 | tee gpurun_out/${tag}_dist_check_full.log
KF_DP_OVERLAP=0 timeout 600 $TR bench.py --gpus $N --workload block --steps 20 --warmup 3 2> gpurun_out/${tag}_block.err | tee gpurun_out/${tag}_block_n${N}.json
KF_DP_OVERLAP=1 timeout 600 $TR bench.py --gpus $N --workload block --steps 20 --warmup 3 2> gpurun_out/${tag}_block_ov.err | tee gpurun_out/${tag}_block_overlap_n${N}.json
grep "block rank 0" gpurun_out/${tag}_block.err gpurun_out/${tag}_block_ov.err
if [ -z "$lean" ]; then
  timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 5 2> gpurun_out/${tag}_gemm.err | tee gpurun_out/${tag}_bench_n${N}.json
  grep "block rank 0" gpurun_out/${tag}_gemm.err
fi
