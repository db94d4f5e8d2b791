"""Attention backward (two-kernel scheme) at C3 under CTA orders KF_ATTN_HG = 1 (head-major) ... (weight-major inside groups of hg heads),
alternating call by call."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kfunca_b200 as kf
from kfunca_b200.runtime import Event
B, H, S, D = 8, 32, 4096, 128
q, k, v, do = (kf.empty([B, H, S, D], kf.bfloat16, 0) for _ in range(4))
for i, t in enumerate((q, k, v, do)): t.random_uniform_(10 + i, -1.0, 1.0)
o, lse = kf.causal_attention_fwd(q, k, v)
os.environ["KF_ATTN_BWD"] = "two"
hgs = ["1", "2", "4", "8", "16"]
times = {h: [] for h in hgs}
for rep in range(12):
    for h in hgs:
        os.environ["KF_ATTN_HG"] = h
        e0, e1 = Event(), Event(); e0.record()
        kf.causal_attention_bwd(do, q, k, v, o, lse)
        e1.record(); e1.synchronize()
        if rep >= 2: times[h].append(e0.elapsed_ms(e1))
for h, t in times.items():
    t.sort(); print(f"bwd two hg={h:3s} min {t[0]:.3f} med {t[len(t)//2]:.3f} ms", flush=True)
