"""Profiling workload: one pass over each hot-path family at its BASELINE size (for `ncu`; never a bench number).

  python tools/prof_ops.py [family ...]      families: ew reduce permute topk gemm gemm_f32 norm attn attn_bwd attn_f32 attn_pers
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import kfunca_b200 as kf

fams = set(sys.argv[1:]) or {"ew", "reduce", "permute", "topk", "gemm", "attn"}
rng = np.random.default_rng(1234)
REPS = int(os.environ.get("KF_PROF_REPS", "2"))

if fams & {"ew", "reduce", "permute"}:
    N = 4096
    a = kf.from_numpy(rng.uniform(-10, 10, (N, N)).astype(np.float32), 0)
    b = kf.from_numpy(rng.uniform(-10, 10, (N, N)).astype(np.float32), 0)
    for _ in range(REPS):
        if "ew" in fams:
            a + b
            a * b
        if "reduce" in fams:
            a.sum(0)
            a.sum(1)
            a.mean(0)
            a.mean(1)
            a.view(-1).sum(0)
        if "permute" in fams:
            a.permute(1, 0).contiguous()
    del a, b
if "topk" in fams:
    x = kf.from_numpy(rng.uniform(-1e5, 1e5, (8192, 32768)).astype(np.float32), 0)
    for _ in range(REPS):
        x.topk(64, 1, True)
    del x
if "gemm" in fams:
    n = 8192
    A = kf.empty([n, n], kf.bfloat16, 0)
    B = kf.empty([n, n], kf.bfloat16, 0)
    A.fill_(0.01)
    B.fill_(0.02)
    for _ in range(REPS):
        kf.gemm(A, B, 1.0, 0.0)
    del A, B
if "gemm_f32" in fams:
    n = 8192
    A = kf.empty([n, n], kf.float, 0)
    B = kf.empty([n, n], kf.float, 0)
    A.random_uniform_(1, -1.0, 1.0)
    B.random_uniform_(2, -1.0, 1.0)
    for _ in range(REPS):
        kf.gemm(A, B, 1.0, 0.0)
    del A, B
if "norm" in fams:
    N = 4096
    a = kf.from_numpy(rng.uniform(-10, 10, (N, N)).astype(np.float32), 0)
    g = kf.from_numpy(rng.uniform(0.5, 1.5, (1, N)).astype(np.float32), 0)
    dy = kf.from_numpy(rng.uniform(-1, 1, (N, N)).astype(np.float32), 0)
    a.set_requires_grad(True)
    for _ in range(REPS):
        a.mean_var(1, False)
        a.norm_stat(0)
        y = kf.layer_norm(a, g, 1e-5)
        a.zero_grad()
        y.backward(dy)
    del a, g, dy, y
if fams & {"attn", "attn_bwd"}:
    Bq, H, S, D = (8, 32, 4096, 128) if "KF_PROF_SMALL" not in os.environ else (1, 4, 4096, 128)
    q = kf.empty([Bq, H, S, D], kf.bfloat16, 0)
    k = kf.empty([Bq, H, S, D], kf.bfloat16, 0)
    v = kf.empty([Bq, H, S, D], kf.bfloat16, 0)
    q.fill_(0.05)
    k.fill_(0.03)
    v.fill_(0.5)
    do = kf.empty([Bq, H, S, D], kf.bfloat16, 0)
    do.fill_(0.1)
    for _ in range(REPS):
        o, lse = kf.causal_attention_fwd(q, k, v)
        if "attn_bwd" in fams:
            kf.causal_attention_bwd(do, q, k, v, o, lse)
if "attn_f32" in fams:  # C3 in fp32: three plane splits + the tcgen05 split-precision kernel
    Bq, H, S, D = 8, 32, 4096, 128
    q, k, v = (kf.empty([Bq, H, S, D], kf.float, 0) for _ in range(3))
    for i, t in enumerate((q, k, v)):
        t.random_uniform_(10 + i, -1.0, 1.0)
    for _ in range(REPS):
        kf.causal_attention_fwd(q, k, v)
    del q, k, v
if "attn_pers" in fams:  # the persistent forward kernel at the length it is the default for (S = 1024, same B H S as C3)
    Bq, H, S, D = 32, 32, 1024, 128
    q, k, v = (kf.empty([Bq, H, S, D], kf.bfloat16, 0) for _ in range(3))
    for i, t in enumerate((q, k, v)):
        t.random_uniform_(10 + i, -1.0, 1.0)
    for _ in range(REPS):
        kf.causal_attention_fwd(q, k, v)
    del q, k, v
kf.synchronize()
print("prof_ops done")
