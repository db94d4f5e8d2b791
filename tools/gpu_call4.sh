#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_block_gpu.py -m gpu -x -q 2>&1 | tail -2
timeout 300 python tools/gpu_block_ab.py 2>&1 | tail -3 | tee gpurun_out/s8o_block_ab.log
timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/s8o_ln.log
import sys; sys.path.insert(0, '.')
import numpy as np, kfunca_b200 as kf
from kfunca_b200.runtime import Event
rng = np.random.default_rng(0); N = 4096
A = [kf.from_numpy(rng.uniform(-1, 1, (N, N)).astype(np.float32), 0) for _ in range(4)]
B = [kf.from_numpy(rng.uniform(-1, 1, (N, N)).astype(np.float32), 0) for _ in range(4)]
gain = kf.from_numpy(rng.uniform(0.5, 1.5, (1, N)).astype(np.float32), 0)
for a in A: a.set_requires_grad(True)
ys = [kf.layer_norm(A[i], gain, 1e-5) for i in range(4)]
def t(fn, iters=30):
    for i in range(5): fn(i % 4)
    e0, e1 = Event(), Event(); e0.record()
    for i in range(iters): fn(i % 4)
    e1.record(); e1.synchronize(); return e0.elapsed_ms(e1) / iters * 1e3
def bwd(i):
    A[i].zero_grad(); ys[i].backward(B[i])
print("ln fwd fp32 4096^2: %.1f us" % t(lambda i: kf.layer_norm(A[i], gain, 1e-5)))
print("ln bwd fp32 4096^2: %.1f us (dx + dgain + fold + grad accumulate copies)" % t(bwd))
PY
