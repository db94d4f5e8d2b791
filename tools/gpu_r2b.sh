#!/bin/bash
# bring-up visit for attention variants: parity + C3 timing per KF_ATTN_BWD mode, then the attention / block tests
set -u
tag=$1; shift
modes=${*:-wide two}
mkdir -p gpurun_out
for m in $modes; do
  KF_ATTN_BWD=$m timeout 300 python tools/gpu_attn.py --parity --bwd > gpurun_out/${tag}_attn_$m.log 2>&1; echo "attn[$m] rc=$?"; tail -8 gpurun_out/${tag}_attn_$m.log
done
timeout 1200 python -m pytest tests/test_attention_gpu.py tests/test_baseline_shapes_gpu.py tests/test_block_gpu.py tests/test_gemm_gpu.py -m gpu -q --durations=5 > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/${tag}_pytest.log
