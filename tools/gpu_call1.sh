#!/bin/bash
# call 1 of the session: new CTA-pair GEMM first (bounded), then the whole GPU suite, both bench arms, launch list
set -u
mkdir -p gpurun_out
tag=s5a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.log 2>&1
timeout 300 python -m pytest tests/test_gemm_gpu.py -m gpu -x -q > gpurun_out/${tag}_pytest_gemm.log 2>&1; echo "pytest gemm rc=$?"
tail -15 gpurun_out/${tag}_pytest_gemm.log
timeout 300 python tools/gpu_gemm2.py > gpurun_out/${tag}_gemm2.log 2>&1; echo "gemm2 rc=$?"
cat gpurun_out/${tag}_gemm2.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/${tag}_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"
tail -2 gpurun_out/${tag}_smoke.log
timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
cat gpurun_out/${tag}_bench.json; tail -5 gpurun_out/${tag}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; echo "ref rc=$?"
cat gpurun_out/${tag}_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
   python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/${tag}_launches_bench.log 2>&1
