#!/bin/bash
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544"
timeout 200 $TR bench.py --gpus 2 --steps 20 --warmup 5 2> gpurun_out/s8q_n2.err | tee gpurun_out/s8q_bench_n2.json | cut -c1-300
grep "block rank 0" gpurun_out/s8q_n2.err | cut -c1-300
timeout 100 $TR bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | cut -c1-200
