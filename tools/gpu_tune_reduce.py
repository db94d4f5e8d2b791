"""sweep split / cluster / lanes-per-row / loads-in-flight / PDL settings of the reductions (env hooks read per call)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import kfunca_b200 as kf
from kfunca_b200.runtime import Event

rng = np.random.default_rng(1)
N = 4096
A = [kf.from_numpy(rng.uniform(-10, 10, (N, N)).astype(np.float32), 0) for _ in range(4)]
flat = [a.view(-1) for a in A]
ref = A[0].numpy().astype(np.float64)


def t(fn, iters=50, warm=5):
    for i in range(warm):
        fn(i % 4)
    e0, e1 = Event(), Event()
    e0.record()
    for i in range(iters):
        fn(i % 4)
    e1.record()
    e1.synchronize()
    return e0.elapsed_ms(e1) / iters * 1e3


def setenv(d):
    for k in ("KF_RED_S", "KF_RED_C", "KF_RED_LPR", "KF_RED_U", "KF_RED_W", "KF_PDL", "KF_RED_STREAM", "KF_RED_STREAM_CTAS", "KF_RED_PUSH"):
        os.environ.pop(k, None)
    for k, v in d.items():
        os.environ[k] = str(v)


HOOKS = ("KF_RED_S", "KF_RED_C", "KF_RED_LPR", "KF_RED_U", "KF_RED_W")
for pdl in ((1,) if "--quick" in sys.argv else (1, 0)):
    print("== PDL", pdl)
    setenv({"KF_PDL": pdl})
    print("defaults: sum0 %.2f  sum1 %.2f  sumall %.2f us" % (t(lambda i: A[i].sum(0)), t(lambda i: A[i].sum(1)), t(lambda i: flat[i].sum(0))), flush=True)
    for push in (1, 0):
        for (S, C, L) in [(None, None, None), (8, 8, 16), (8, 8, 32), (4, 4, 8), (4, 4, 16), (8, 8, 8), (16, 16, 16)]:
            d = {"KF_PDL": pdl, "KF_RED_PUSH": push}
            if S:
                d.update({"KF_RED_S": S, "KF_RED_C": C, "KF_RED_LPR": L})
            setenv(d)
            try:
                ok = np.allclose(A[0].sum(0).numpy(), ref.sum(0, keepdims=True), rtol=1e-4, atol=1e-2)
                print("  cols push=%d S=%s C=%s L=%s: sum0 %.2f us  mean0 %.2f us ok=%s" % (push, S, C, L, t(lambda i: A[i].sum(0)), t(lambda i: A[i].mean(0)), ok), flush=True)
            except Exception as e:
                print("  cols push=%d S=%s C=%s L=%s: FAILED %s" % (push, S, C, L, str(e)[:100]), flush=True)
    for push in (1, 0):
        for (S, C) in [(None, None), (512, 8), (512, 4), (512, 2), (512, 1), (592, 8), (592, 4), (592, 1), (296, 1), (296, 4), (1024, 8), (1024, 16)]:
            d = {"KF_PDL": pdl, "KF_RED_PUSH": push}
            if S:
                d.update({"KF_RED_S": S, "KF_RED_C": C})
            setenv(d)
            try:
                ok = np.allclose(flat[0].sum(0).numpy(), ref.sum(), rtol=1e-4, atol=1e-1)
                print("  all push=%d S=%s C=%s: sumall %.2f us ok=%s" % (push, S, C, t(lambda i: flat[i].sum(0)), ok), flush=True)
            except Exception as e:
                print("  all push=%d S=%s C=%s: FAILED %s" % (push, S, C, str(e)[:100]), flush=True)
    if "--quick" in sys.argv:
        continue
    for st in (0, 1, 2):
        setenv({"KF_PDL": pdl, "KF_RED_STREAM": 1 if st else 0, "KF_RED_STREAM_CTAS": max(st, 1)})
        ok = np.allclose(A[0].sum(0).numpy(), ref.sum(0, keepdims=True), rtol=1e-4, atol=1e-2)
        print("  cols stream=%d ctas/sm=%d: sum0 %.2f us  mean0 %.2f us ok=%s" % (1 if st else 0, max(st, 1), t(lambda i: A[i].sum(0)), t(lambda i: A[i].mean(0)), ok), flush=True)
    setenv({"KF_PDL": pdl, "KF_RED_STREAM": 0})
    for W in (2, 4, 8):
        setenv({"KF_PDL": pdl, "KF_RED_W": W})
        ok = np.allclose(A[0].sum(1).numpy(), ref.sum(1, keepdims=True), rtol=1e-4, atol=1e-2)
        print("  rows W=%d: sum1 %.2f us ok=%s" % (W, t(lambda i: A[i].sum(1)), ok), flush=True)
    for (S, C, L) in [(8, 8, 16), (8, 8, 8), (8, 8, 32), (16, 8, 16), (16, 16, 16), (16, 16, 8), (32, 8, 32), (8, 1, 16)]:
        for U in (8, 16):
            setenv({"KF_PDL": pdl, "KF_RED_STREAM": 0, "KF_RED_S": S, "KF_RED_C": C, "KF_RED_LPR": L, "KF_RED_U": U})
            try:
                ok = np.allclose(A[0].sum(0).numpy(), ref.sum(0, keepdims=True), rtol=1e-4, atol=1e-2)
                print("  cols S=%d C=%d L=%d U=%d: sum0 %.2f us ok=%s" % (S, C, L, U, t(lambda i: A[i].sum(0)), ok), flush=True)
            except Exception as e:
                print("  cols S=%d C=%d L=%d U=%d: FAILED %s" % (S, C, L, U, str(e)[:100]), flush=True)
    for (S, C) in [(512, 8), (256, 8), (1024, 8), (512, 16), (296, 8), (592, 8), (148, 4)]:
        setenv({"KF_PDL": pdl, "KF_RED_S": S, "KF_RED_C": C})
        try:
            ok = np.allclose(flat[0].sum(0).numpy(), ref.sum(), rtol=1e-4, atol=1e-1)
            print("  all S=%d C=%d: sumall %.2f us ok=%s" % (S, C, t(lambda i: flat[i].sum(0)), ok), flush=True)
        except Exception as e:
            print("  all S=%d C=%d: FAILED %s" % (S, C, str(e)[:100]), flush=True)
