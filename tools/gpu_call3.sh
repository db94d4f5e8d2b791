#!/bin/bash
# session-8 visit 1: validate HEAD (parity), reduce sweep, attention exp2-share sweep, bench
set -u
mkdir -p gpurun_out
tag=s8a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -6 gpurun_out/${tag}_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${tag}_smoke.log
tail -2 gpurun_out/${tag}_smoke.log
timeout 300 python tools/gpu_tune_reduce.py > gpurun_out/${tag}_tune_reduce.log 2>&1; echo "tune rc=$?"
cat gpurun_out/${tag}_tune_reduce.log
for poly in 0 2 3 4; do
  KF_ATTN_POLY=$poly timeout 120 python tools/gpu_attn.py 2>&1 | tail -2
done | tee gpurun_out/${tag}_attn_poly.log
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
cat gpurun_out/${tag}_bench.json
