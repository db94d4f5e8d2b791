"""GEMM variant timing on the GPU box: single-CTA vs CTA-pair kernel at the BASELINE shape and the block's shapes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import kfunca_b200 as kf
from kfunca_b200.runtime import Event

rng = np.random.default_rng(0)


def rnd(shape):
    t = kf.from_numpy(rng.uniform(-1, 1, shape).astype(np.float32), 0)
    return t.bfloat16()


def bench(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = Event(), Event()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_ms(e1) / iters


shapes = [(8192, 8192, 8192, False, False), (4096, 4096, 12288, False, False), (4096, 16384, 4096, False, False),
          (32768, 4096, 4096, False, False), (4096, 32768, 4096, True, False), (32768, 4096, 16384, False, True)]
for (m, k, n, ta, tb) in shapes:
    a = rnd((k, m) if ta else (m, k))
    b = rnd((n, k) if tb else (k, n))
    res = {}
    for cg in ("1", "2"):
        os.environ["KF_GEMM_CTA_GROUP"] = cg
        ms = bench(lambda: kf.matmul(a, ta, b, tb, 1.0))
        res[cg] = ms
    os.environ.pop("KF_GEMM_CTA_GROUP")
    fl = 2.0 * m * n * k
    print(f"gemm M={m} K={k} N={n} ta={ta} tb={tb}: 1cta {res['1']:.4f} ms {fl/res['1']/1e9:.0f} TF | pair {res['2']:.4f} ms {fl/res['2']/1e9:.0f} TF", flush=True)
    del a, b
