"""GEMM bring-up diagnostics: prints error statistics per layout/shape instead of asserting, then timings."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import kfunca_b200 as kf
from kfunca_b200.runtime import Event
from oracle import oracle as O

rng = np.random.default_rng(5)
def g(a): return kf.from_numpy(a, 0)
def b16(x): return x.astype(np.float32).astype(O.bfloat16)

for (ta, tb) in [(False, True), (False, False), (True, False), (True, True)]:
    for (m, k, n) in [(128, 64, 128), (128, 64, 256), (128, 128, 256), (256, 256, 512), (200, 136, 72)]:
        a = b16(rng.uniform(-1, 1, (k, m) if ta else (m, k)))
        b = b16(rng.uniform(-1, 1, (n, k) if tb else (k, n)))
        try:
            out = kf.matmul(g(a), ta, g(b), tb, 1.0).float().numpy().astype(np.float64)
        except RuntimeError as e:
            print("FAIL", ta, tb, m, k, n, str(e)[:200]); continue
        af = a.astype(np.float64).T if ta else a.astype(np.float64)
        bf = b.astype(np.float64).T if tb else b.astype(np.float64)
        ex = af @ bf
        err = np.abs(out - ex)
        print(f"ta={int(ta)} tb={int(tb)} {m}x{k}x{n}: max_err={err.max():.4g} mean_err={err.mean():.4g} ref_absmean={np.abs(ex).mean():.4g} frac_bad={(err > 0.05*np.abs(ex)+0.05).mean():.4f}")
        if err.max() > 0.5:
            bad = np.argwhere(err > 0.5)
            print("   first bad idx", bad[:5].tolist(), "rows bad:", np.unique(bad[:,0])[:10].tolist(), "cols bad:", np.unique(bad[:,1])[:10].tolist())

def timeit(name, fn, flops, iters=20, warm=3):
    for _ in range(warm): fn()
    e0, e1 = Event(), Event(); e0.record()
    for _ in range(iters): fn()
    e1.record(); e1.synchronize()
    ms = e0.elapsed_ms(e1) / iters
    print(f"{name:34s} {ms:8.3f} ms  {flops/ms/1e9:9.1f} TFLOP/s ({flops/ms/1e9/1634.2*100:5.1f}% of measured burst 1634.2)")

for n in (2048, 4096, 8192):
    A = g(b16(rng.uniform(-1, 1, (n, n)))); B = g(b16(rng.uniform(-1, 1, (n, n))))
    timeit(f"gemm bf16 {n}^3 (B [K,N])", lambda: kf.gemm(A, B, 1.0, 0.0), 2 * n**3)
    timeit(f"matmul bf16 {n}^3 (B [N,K])", lambda: kf.matmul(A, False, B, True, 1.0), 2 * n**3)
A = g(rng.uniform(-1, 1, (4096, 4096)).astype(np.float32)); B = g(rng.uniform(-1, 1, (4096, 4096)).astype(np.float32))
timeit("gemm fp32 4096^3 (SIMT)", lambda: kf.gemm(A, B, 1.0, 0.0), 2 * 4096**3, iters=5, warm=1)
