"""First-contact GPU script: device info + C1 timings through the public API with CUDA events."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import kfunca_b200 as kf
from kfunca_b200.runtime import Event, launch_count

kf.device_info()
rng = np.random.default_rng(1234)
N = 4096
NSETS = 4  # rotate inputs: 4 x (2 x 64 MiB) > 126 MB L2
A = [kf.from_numpy(rng.uniform(-10, 10, (N, N)).astype(np.float32), 0) for _ in range(NSETS)]
B = [kf.from_numpy(rng.uniform(-10, 10, (N, N)).astype(np.float32), 0) for _ in range(NSETS)]

def timeit(name, fn, bytes_alg, iters=40, warm=5):
    for i in range(warm):
        fn(i % NSETS)
    e0, e1 = Event(), Event()
    e0.record()
    for i in range(iters):
        fn(i % NSETS)
    e1.record(); e1.synchronize()
    ms = e0.elapsed_ms(e1) / iters
    print(f"{name:28s} {ms*1e3:9.2f} us  {bytes_alg/ms/1e6:9.1f} GB/s  ({bytes_alg/ms/1e6/6555.8*100:5.1f}% of measured 6555.8)")

nb = N * N * 4
timeit("add fp32 4096^2", lambda i: A[i] + B[i], 3 * nb)
timeit("mul fp32 4096^2", lambda i: A[i] * B[i], 3 * nb)
timeit("iadd fp32 4096^2", lambda i: A[i].__iadd__(B[i]), 3 * nb)
timeit("add scalar", lambda i: A[i] + 2.0, 2 * nb)
timeit("sum dim0", lambda i: A[i].sum(0), nb + N * 4)
timeit("sum dim1", lambda i: A[i].sum(1), nb + N * 4)
timeit("mean dim0", lambda i: A[i].mean(0), nb + N * 4)
timeit("mean dim1", lambda i: A[i].mean(1), nb + N * 4)
flat = [a.view(-1) for a in A]
timeit("sum all", lambda i: flat[i].sum(0), nb + 4)
timeit("permute(1,0).contiguous", lambda i: A[i].permute(1, 0).contiguous(), 2 * nb)
timeit("clone", lambda i: A[i].clone(), 2 * nb)
H = [a.bfloat16() for a in A]; H2 = [b.bfloat16() for b in B]
timeit("add bf16 4096^2", lambda i: H[i] + H2[i], 3 * nb // 2)
timeit("add bf16+fp32 mixed", lambda i: H[i] + B[i], nb // 2 + 2 * nb)
timeit("sum dim1 bf16", lambda i: H[i].sum(1), nb // 2)
big = kf.from_numpy(rng.uniform(-10, 10, (2, 1024, 1024, 64)).astype(np.float32), 0)
bb = kf.from_numpy(rng.uniform(-10, 10, (2, 1024, 1, 64)).astype(np.float32), 0)
timeit("bcast add [2,1024,1024,64]", lambda i: big + bb, 2 * big.numel() * 4)
print("launches", launch_count())
kf.memstat()
