"""sweep split/cluster settings of the column / full reductions (env hooks KF_RED_S / KF_RED_C), one subprocess per setting"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, os
sys.path.insert(0, %r)
import numpy as np
import kfunca_b200 as kf
from kfunca_b200.runtime import Event
rng = np.random.default_rng(1)
N = 4096
A = [kf.from_numpy(rng.uniform(-10, 10, (N, N)).astype(np.float32), 0) for _ in range(4)]
flat = [a.view(-1) for a in A]
def t(fn, iters=40, warm=5):
    for i in range(warm): fn(i %% 4)
    e0, e1 = Event(), Event(); e0.record()
    for i in range(iters): fn(i %% 4)
    e1.record(); e1.synchronize()
    return e0.elapsed_ms(e1) / iters * 1e3
ref = A[0].numpy().astype(np.float64)
ok0 = np.allclose(A[0].sum(0).numpy(), ref.sum(0, keepdims=True), rtol=1e-4, atol=1e-2)
okA = np.allclose(flat[0].sum(0).numpy(), ref.sum(), rtol=1e-4, atol=1e-1)
print("S=%%s C=%%s L=%%s  sum0 %%.2f us  sumall %%.2f us  sum1 %%.2f us ok=%%s,%%s" %% (os.environ.get("KF_RED_S"), os.environ.get("KF_RED_C"), os.environ.get("KF_RED_LPR"), t(lambda i: A[i].sum(0)), t(lambda i: flat[i].sum(0)), t(lambda i: A[i].sum(1)), ok0, okA))
''' % ROOT
for S, C, L in [(None, None, None), (8, 8, 8), (8, 8, 16), (16, 8, 8), (16, 1, 32), (8, 8, 32), (4, 4, 8)]:
    env = dict(os.environ)
    if S is not None:
        env["KF_RED_S"], env["KF_RED_C"], env["KF_RED_LPR"] = str(S), str(C), str(L)
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=120)
    print((r.stdout.strip() or r.stderr.strip()[-300:]))
