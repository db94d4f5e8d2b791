"""Attention bring-up diagnostics + timings."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import kfunca_b200 as kf
from kfunca_b200.runtime import Event
from oracle import oracle as O

rng = np.random.default_rng(5)
def g(a): return kf.from_numpy(a, 0)
def b16(x): return x.astype(np.float32).astype(O.bfloat16)

for (b, h, sq, skv, d) in [(1, 1, 128, 128, 128), (1, 1, 128, 128, 64), (1, 2, 256, 256, 128), (1, 2, 384, 384, 64), (1, 1, 200, 333, 128), (1, 2, 1024, 1024, 128)]:
    q, k, v = (b16(rng.uniform(-1, 1, s)) for s in ((b, h, sq, d), (b, h, skv, d), (b, h, skv, d)))
    try:
        out, lse = kf.causal_attention_fwd(g(q), g(k), g(v))
        out = out.float().numpy().astype(np.float64); lse = lse.numpy()
    except RuntimeError as e:
        print("FAIL", (b, h, sq, skv, d), str(e)[:300]); continue
    ex, lex = O.causal_attention(q, k, v, return_lse=True)
    err = np.abs(out - ex)
    print(f"attn {b}x{h}x{sq}x{skv}x{d}: max_err={err.max():.4g} mean_err={err.mean():.4g} ref_absmean={np.abs(ex).mean():.4g} lse_err={np.abs(lse-lex).max():.4g} nan={np.isnan(out).sum()}")
    if err.max() > 0.05:
        bad = np.argwhere(err > 0.05)
        print("   bad rows(q):", np.unique(bad[:, 2])[:16].tolist(), " bad cols(d):", np.unique(bad[:, 3])[:16].tolist(), "count", len(bad))

def timeit(name, fn, flops, iters=10, warm=2):
    for _ in range(warm): fn()
    e0, e1 = Event(), Event(); e0.record()
    for _ in range(iters): fn()
    e1.record(); e1.synchronize()
    ms = e0.elapsed_ms(e1) / iters
    print(f"{name:40s} {ms:8.3f} ms  {flops/ms/1e9:9.1f} TFLOP/s ({flops/ms/1e9/1634.2*100:5.1f}% of measured burst 1634.2)")

for (B, H, S, D) in [(2, 32, 4096, 128), (8, 32, 4096, 128), (8, 32, 4096, 64)]:
    q = kf.empty([B, H, S, D], kf.bfloat16, 0); q.fill_(0.01)
    k = kf.empty([B, H, S, D], kf.bfloat16, 0); k.fill_(0.02)
    v = kf.empty([B, H, S, D], kf.bfloat16, 0); v.fill_(0.5)
    timeit(f"attn fwd bf16 B{B} H{H} S{S} D{D}", lambda: kf.causal_attention(q, k, v), 4 * B * H * S * S * D / 2)
