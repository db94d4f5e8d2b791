"""top-k timing at C4 shape (reduced row count) + exactness vs the oracle on a few rows."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import kfunca_b200 as kf
from kfunca_b200.runtime import Event
from oracle import oracle as O

rng = np.random.default_rng(7)
rows, cols, k = 8192, 32768, 64
x = rng.uniform(-1e5, 1e5, (rows, cols)).astype(np.float32)
X = kf.from_numpy(x, 0)
for largest in (True, False):
    v, i = X.topk(k, 1, largest)
    vn, inn = v.numpy(), i.numpy()
    ev, ei = O.topk(x[:64], k, 1, largest)
    print("largest", largest, "values ok", np.array_equal(vn[:64], ev), "indices ok", np.array_equal(inn[:64], ei))
    ev, ei = O.topk(x[-32:], k, 1, largest)
    print("   tail rows ok", np.array_equal(vn[-32:], ev) and np.array_equal(inn[-32:], ei))
def timeit(name, fn, bytes_alg, iters=5, warm=2):
    for _ in range(warm): fn()
    e0, e1 = Event(), Event(); e0.record()
    for _ in range(iters): fn()
    e1.record(); e1.synchronize()
    ms = e0.elapsed_ms(e1) / iters
    print(f"{name:34s} {ms*1e3:9.1f} us  {bytes_alg/ms/1e6:8.1f} GB/s ({bytes_alg/ms/1e6/6553.6*100:5.1f}% of measured 6553.6)")
timeit("topk64 8192x32768 fp32", lambda: X.topk(k, 1, True), rows * cols * 4 + rows * k * 12)
timeit("topk8 8192x32768 fp32", lambda: X.topk(8, 1, True), rows * cols * 4 + rows * 8 * 12)
timeit("topk256 8192x32768 fp32", lambda: X.topk(256, 1, True), rows * cols * 4 + rows * 256 * 12)
del X
x2 = rng.uniform(-1e5, 1e5, (32768, 8192)).astype(np.float32)
X2 = kf.from_numpy(x2, 0)
timeit("topk64 32768x8192 fp32", lambda: X2.topk(k, 1, True), x2.size * 4 + x2.shape[0] * k * 12)
