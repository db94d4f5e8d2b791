"""fp32 causal attention forward: the tcgen05 split-precision kernel (default) against the FFMA kernel (KF_ATTN_F32=simt) and the
float64 oracle — parity on ragged / rectangular / D = 64 / long shapes, U(-1, 1) and the reference's U(-10, 10) inputs, then the C3
timing (B=8 H=32 S=4096 D=128, fp32)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import kfunca_b200 as kf
from kfunca_b200.runtime import Event
from oracle import oracle as O

rng = np.random.default_rng(7)
g = lambda a: kf.from_numpy(a, 0)


def run(mode, q, k, v):
    if mode is None:
        os.environ.pop("KF_ATTN_F32", None)
    else:
        os.environ["KF_ATTN_F32"] = mode
    out, lse = kf.causal_attention_fwd(q, k, v)
    return out.numpy(), lse.numpy()


bad = 0
for (b, h, sq, skv, d, lo, hi) in [(1, 2, 256, 256, 128, -1, 1), (1, 1, 200, 333, 128, -1, 1), (2, 2, 640, 640, 64, -1, 1), (3, 5, 130, 130, 128, -1, 1),
                                   (2, 3, 1, 1, 128, -1, 1), (1, 3, 513, 257, 64, -1, 1), (1, 2, 100, 700, 128, -1, 1), (1, 2, 2048, 2048, 128, -1, 1),
                                   (2, 4, 32, 256, 128, -10, 10), (3, 5, 64, 32, 64, -10, 10), (1, 1, 4096, 4096, 128, 0, 1)]:
    q, k, v = (rng.uniform(lo, hi, s).astype(np.float32) for s in ((b, h, sq, d), (b, h, skv, d), (b, h, skv, d)))
    gq, gk, gv = g(q), g(k), g(v)
    o_t, l_t = run(None, gq, gk, gv)
    o_s, l_s = run("simt", gq, gk, gv)
    eo = O.causal_attention(q, k, v)
    tol = 1e-5 if hi == 1 else 1e-3
    err_t = np.abs(o_t - eo) / (tol * np.abs(eo) + tol)
    err_s = np.abs(o_s - eo) / (tol * np.abs(eo) + tol)
    lse_d = np.abs(l_t - l_s).max()
    ok = err_t.max() <= 1.0 and lse_d < 1e-3
    bad += 0 if ok else 1
    print(f"{b}x{h}x{sq}x{skv}x{d} U({lo},{hi}): tensor-core err/tol max {err_t.max():.3f}  FFMA err/tol max {err_s.max():.3f}  "
          f"max|out - oracle| {np.abs(o_t - eo).max():.3g} (FFMA {np.abs(o_s - eo).max():.3g})  lse diff {lse_d:.2g}  {'ok' if ok else 'FAIL'}", flush=True)
print("PARITY", "FAIL" if bad else "OK", flush=True)

B, H, S, D = 8, 32, 4096, 128
mk = lambda seed: kf.empty([B, H, S, D], kf.float, 0)
Q, K, V = mk(1), mk(2), mk(3)
for i, t in enumerate((Q, K, V)):
    t.random_uniform_(10 + i, -1.0, 1.0)
fl = 4 * B * H * S * S * D / 2
for mode, iters in ((None, 10), ("simt", 2)):
    if mode is None:
        os.environ.pop("KF_ATTN_F32", None)
    else:
        os.environ["KF_ATTN_F32"] = mode
    kf.causal_attention_fwd(Q, K, V)
    ev = [Event() for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        kf.causal_attention_fwd(Q, K, V)
        ev[i + 1].record()
    ev[-1].synchronize()
    ms = sorted(ev[i].elapsed_ms(ev[i + 1]) for i in range(iters))
    print(f"C3 fp32 forward, {mode or 'tcgen05 (3 bf16 planes, 6 products)'}: median {ms[len(ms) // 2]:.3f} ms  min {ms[0]:.3f} ms  {fl / ms[len(ms) // 2] / 1e9:.1f} TFLOP/s", flush=True)
