#!/bin/bash
set -u
mkdir -p gpurun_out
tag=s8m
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_block_launches.csv \
   python tools/gpu_block_probe.py --once > gpurun_out/${tag}_block_ncu.log 2>&1
tail -2 gpurun_out/${tag}_block_ncu.log
