"""Attention bring-up: parity vs the oracle on a few shapes + C3 timings (fwd, bwd, fwd+bwd).  KF_ATTN_POLY selects the
forward's FMA-pipe exp2 share (read once per process)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import kfunca_b200 as kf
from kfunca_b200.runtime import Event
from oracle import oracle as O

rng = np.random.default_rng(11)
g = lambda a: kf.from_numpy(a, 0)
b16 = lambda x: x.astype(np.float32).astype(O.bfloat16)
print("KF_ATTN_POLY =", os.environ.get("KF_ATTN_POLY", "default"))
if "--parity" in sys.argv:
    for (b, h, sq, skv, d) in [(1, 2, 256, 256, 128), (1, 1, 200, 333, 128), (1, 2, 1024, 1024, 128), (2, 2, 640, 640, 64)]:
        q, k, v = (b16(rng.uniform(-1, 1, s)) for s in ((b, h, sq, d), (b, h, skv, d), (b, h, skv, d)))
        do = b16(rng.uniform(-1, 1, (b, h, sq, d)))
        out, lse = kf.causal_attention_fwd(g(q), g(k), g(v))
        eo = O.causal_attention(q, k, v)
        of = out.float().numpy()
        dq, dk, dv = kf.causal_attention_bwd(g(do), g(q), g(k), g(v), out, lse)
        exp = O.causal_attention_bwd(q, k, v, do)
        msg = [f"out max_abs_err={np.abs(of - eo).max():.3g}"]
        for name, a_, e_ in zip(("dq", "dk", "dv"), (dq, dk, dv), exp):
            a_ = a_.float().numpy().astype(np.float64)
            msg.append(f"{name} max_rel={np.abs(a_ - e_).max() / np.abs(e_).max():.3g}")
        print(f"{b}x{h}x{sq}x{skv}x{d}: " + "  ".join(msg))

def timeit(name, fn, flops, iters=8, warm=3):
    for _ in range(warm): fn()
    e0, e1 = Event(), Event(); e0.record()
    for _ in range(iters): fn()
    e1.record(); e1.synchronize()
    ms = e0.elapsed_ms(e1) / iters
    print(f"{name:40s} {ms:8.3f} ms  {flops/ms/1e9:9.1f} TFLOP/s ({flops/ms/1e9/1694.9*100:5.1f}% of measured burst 1694.9)")
    return ms

B, H, S, D = 8, 32, 4096, 128
mk = lambda: g(b16(rng.uniform(-1, 1, (1, H, S, D))))  # one batch of random data, repeated over B via cat (keeps host time small)
Q, K, V, dO = (kf.cat([t] * B, 0) for t in (mk(), mk(), mk(), mk()))
fwd = 4 * B * H * S * S * D / 2
out, lse = kf.causal_attention_fwd(Q, K, V)
timeit("attn fwd bf16 C3 (random data)", lambda: kf.causal_attention_fwd(Q, K, V), fwd)
if "--bwd" in sys.argv:
    timeit("attn bwd bf16 C3 (random data)", lambda: kf.causal_attention_bwd(dO, Q, K, V, out, lse), 2.5 * fwd)
    def both():
        o, l = kf.causal_attention_fwd(Q, K, V)
        kf.causal_attention_bwd(dO, Q, K, V, o, l)
    timeit("attn fwd+bwd bf16 C3 (random data)", both, 3.5 * fwd)
