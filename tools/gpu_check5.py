"""Attention backward bring-up: per-gradient errors vs the oracle + C3 timings (fwd, bwd, fwd+bwd)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import kfunca_b200 as kf
from kfunca_b200.runtime import Event
from oracle import oracle as O

rng = np.random.default_rng(11)
def g(a): return kf.from_numpy(a, 0)
def b16(x): return x.astype(np.float32).astype(O.bfloat16)
for (b, h, sq, skv, d) in [(1, 1, 128, 128, 128), (1, 1, 128, 128, 64), (1, 2, 256, 256, 128), (1, 1, 200, 333, 128), (1, 2, 1024, 1024, 128)]:
    q, k, v = (b16(rng.uniform(-1, 1, s)) for s in ((b, h, sq, d), (b, h, skv, d), (b, h, skv, d)))
    do = b16(rng.uniform(-1, 1, (b, h, sq, d)))
    try:
        out, lse = kf.causal_attention_fwd(g(q), g(k), g(v))
        dq, dk, dv = kf.causal_attention_bwd(g(do), g(q), g(k), g(v), out, lse)
        got = [t.float().numpy().astype(np.float64) for t in (dq, dk, dv)]
    except RuntimeError as e:
        print("FAIL", (b, h, sq, skv, d), str(e)[:300]); continue
    exp = O.causal_attention_bwd(q, k, v, do)
    msg = []
    for name, a_, e_ in zip(("dq", "dk", "dv"), got, exp):
        err = np.abs(a_ - e_); sc = np.abs(e_).max()
        msg.append(f"{name}: max_rel={err.max()/sc:.3g} nan={np.isnan(a_).sum()}")
        if err.max() / sc > 0.05:
            bad = np.argwhere(err > 0.05 * sc)
            msg.append(f"[bad rows {np.unique(bad[:,2])[:8].tolist()} cols {np.unique(bad[:,3])[:8].tolist()} n={len(bad)}]")
    print(f"bwd {b}x{h}x{sq}x{skv}x{d}: " + "  ".join(msg))

def timeit(name, fn, flops, iters=5, warm=2):
    for _ in range(warm): fn()
    e0, e1 = Event(), Event(); e0.record()
    for _ in range(iters): fn()
    e1.record(); e1.synchronize()
    ms = e0.elapsed_ms(e1) / iters
    print(f"{name:40s} {ms:8.3f} ms  {flops/ms/1e9:9.1f} TFLOP/s ({flops/ms/1e9/1654.3*100:5.1f}% of measured burst 1654.3)")
    return ms
for (B, H, S, D) in [(8, 32, 4096, 128)]:
    q = g(b16(rng.uniform(-1, 1, (1, 1, S, D)))); 
    Q = kf.empty([B, H, S, D], kf.bfloat16, 0); Q.fill_(0.01)
    K = kf.empty([B, H, S, D], kf.bfloat16, 0); K.fill_(0.02)
    V = kf.empty([B, H, S, D], kf.bfloat16, 0); V.fill_(0.5)
    dO = kf.empty([B, H, S, D], kf.bfloat16, 0); dO.fill_(0.1)
    fwd = 4 * B * H * S * S * D / 2
    out, lse = kf.causal_attention_fwd(Q, K, V)
    timeit(f"attn fwd bf16 B{B} H{H} S{S} D{D}", lambda: kf.causal_attention_fwd(Q, K, V), fwd)
    timeit(f"attn bwd bf16 B{B} H{H} S{S} D{D}", lambda: kf.causal_attention_bwd(dO, Q, K, V, out, lse), 2.5 * fwd)
    def both():
        o, l = kf.causal_attention_fwd(Q, K, V)
        kf.causal_attention_bwd(dO, Q, K, V, o, l)
    timeit(f"attn fwd+bwd bf16 B{B} H{H} S{S} D{D}", both, 3.5 * fwd)
