"""The memory-bound C1 / norm-family lines of bench.py on their own (same protocol: four rotating 64 MiB input sets, 5 warm-ups, 30
back-to-back calls between two events), for quick A/B runs of a reduce / norm kernel change.  Usage: python tools/gpu_mem_ops.py [name ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import kfunca_b200 as kf
from kfunca_b200.runtime import Event

HBM = 6534.8
want = set(sys.argv[1:])
rng = np.random.default_rng(1234)
N, NSETS = 4096, 4
A = [kf.from_numpy(rng.uniform(-10, 10, (N, N)).astype(np.float32), 0) for _ in range(NSETS)]
B = [kf.from_numpy(rng.uniform(-10, 10, (N, N)).astype(np.float32), 0) for _ in range(NSETS)]
flat = [a.view(-1) for a in A]
gain = kf.from_numpy(rng.uniform(0.5, 1.5, (1, N)).astype(np.float32), 0)
nb = N * N * 4


def t(fn, iters=30, warm=5):
    for i in range(warm):
        fn(i % NSETS)
    e0, e1 = Event(), Event()
    e0.record()
    for i in range(iters):
        fn(i % NSETS)
    e1.record()
    e1.synchronize()
    return e0.elapsed_ms(e1) / iters


cases = [
    ("add", lambda i: A[i] + B[i], 3 * nb),
    ("sum_dim0", lambda i: A[i].sum(0), nb + N * 4),
    ("sum_dim1", lambda i: A[i].sum(1), nb + N * 4),
    ("mean_dim0", lambda i: A[i].mean(0), nb + N * 4),
    ("sum_all", lambda i: flat[i].sum(0), nb + 4),
    ("permute", lambda i: A[i].permute(1, 0).contiguous(), 2 * nb),
    ("mean_var_dim1", lambda i: A[i].mean_var(1, False), nb + 2 * N * 4),
    ("norm_stat_dim0", lambda i: A[i].norm_stat(0), nb + 2 * N * 4),
    ("mean_var_dim0", lambda i: A[i].mean_var(0, False), nb + 2 * N * 4),
    ("layer_norm_fwd", lambda i: kf.layer_norm(A[i], gain, 1e-5), 2 * nb),
    ("rms_norm_fwd", lambda i: kf.rms_norm(A[i], gain, 1e-5), 2 * nb),
]
for rep in range(2):
    for name, fn, byts in cases:
        if want and name not in want:
            continue
        ms = t(fn)
        print(f"{name:16s} {ms * 1e3:7.2f} us  {byts / ms / 1e6:7.1f} GB/s  {byts / ms / 1e6 / HBM:.3f} of {HBM}", flush=True)
for a in A:
    a.set_requires_grad(True)
ys = [kf.layer_norm(A[i], gain, 1e-5) for i in range(NSETS)]


def ln_bwd(i):
    A[i].zero_grad()
    ys[i].backward(B[i])


if not want or "layer_norm_bwd" in want:
    for rep in range(2):
        ms = t(ln_bwd)
        print(f"{'layer_norm_bwd':16s} {ms * 1e3:7.2f} us  {3 * nb / ms / 1e6:7.1f} GB/s  {3 * nb / ms / 1e6 / HBM:.3f} of {HBM} (through autograd)", flush=True)
