#!/bin/bash
# Round-2 GPU-box visit.  Usage: tools/gpu_r2.sh <tag> [golden] [tests] [bench] [ref]
set -u
tag=$1; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.log 2>&1
for what in "$@"; do
  case $what in
    golden)
      timeout 300 python oracle/make_golden_from_ref.py r2 > gpurun_out/${tag}_golden.log 2>&1; echo "golden rc=$?"; tail -3 gpurun_out/${tag}_golden.log ;;
    tests)
      timeout 2400 python -m pytest tests -m gpu -q -x --durations=15 > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
      tail -40 gpurun_out/${tag}_pytest.log ;;
    newtests)
      timeout 2400 python -m pytest tests/test_round2_ops_gpu.py tests/test_baseline_shapes_gpu.py tests/test_ref_suite_gpu.py -m gpu -q --durations=15 > gpurun_out/${tag}_pytest_new.log 2>&1
      echo "pytest rc=$?" >> gpurun_out/${tag}_pytest_new.log; tail -60 gpurun_out/${tag}_pytest_new.log ;;
    smoke)
      timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${tag}_smoke.log
      tail -2 gpurun_out/${tag}_smoke.log ;;
    bench)
      timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
      cat gpurun_out/${tag}_bench.json; tail -25 gpurun_out/${tag}_bench.err ;;
    ref)
      timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; echo "ref rc=$?"
      cat gpurun_out/${tag}_bench_ref.json; tail -5 gpurun_out/${tag}_bench_ref.err ;;
    launches)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
         python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/${tag}_launches_bench.log 2>&1 ;;
    *) echo "unknown step $what" ;;
  esac
done
