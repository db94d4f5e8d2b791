"""Attention backward A/B at C3 (bf16 B=8 H=32 S=4096 D=128) in one process: the two deterministic two-kernel schemes (KF_ATTN_BWD=two |
wide) under the CTA orders KF_ATTN_HG = 1 (head-major), auto (head groups that fit the L2, weight-major inside), all (one global
weight-major list).  The variants alternate call by call (the GPU is power-capped); gradients are compared bit for bit across
orders (the order only decides which CTA computes a block)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import kfunca_b200 as kf
from kfunca_b200.runtime import Event
from oracle import oracle as O

rng = np.random.default_rng(11)
g = lambda a: kf.from_numpy(a, 0)
b16 = lambda x: x.astype(np.float32).astype(O.bfloat16)
B, H, S, D = 8, 32, 4096, 128
if len(sys.argv) > 1:
    S = int(sys.argv[1])
    B = max(1, 8 * 4096 // S)
mk = lambda: g(b16(rng.uniform(-1, 1, (1, H, S, D))))
Q, K, V, dO = (kf.cat([t] * B, 0) if B > 1 else t for t in (mk(), mk(), mk(), mk()))
fl = 2.5 * 4 * B * H * S * S * D / 2
out, lse = kf.causal_attention_fwd(Q, K, V)

variants = [(m, hg) for m in ("two", "wide") for hg in ("1", "auto", "all")]
times = {v: [] for v in variants}
ref = {}
for rep in range(14):
    for v in variants:
        os.environ["KF_ATTN_BWD"] = v[0]
        if v[1] == "auto":
            os.environ.pop("KF_ATTN_HG", None)
        else:
            os.environ["KF_ATTN_HG"] = "1" if v[1] == "1" else "1000000"
        e0, e1 = Event(), Event()
        e0.record()
        dq, dk, dv = kf.causal_attention_bwd(dO, Q, K, V, out, lse)
        e1.record()
        e1.synchronize()
        if rep >= 2:
            times[v].append(e0.elapsed_ms(e1))
        if rep == 0:
            sig = tuple(float(t.float().sum(3).sum(2).sum(1).sum(0).numpy().ravel()[0]) for t in (dq, dk, dv))
            ref.setdefault(v[0], sig)
            if sig != ref[v[0]]:
                print(f"MISMATCH {v}: {sig} vs {ref[v[0]]}", flush=True)
        del dq, dk, dv
for v, t in times.items():
    t.sort()
    med = t[len(t) // 2]
    print(f"S={S} {v[0]:5s} hg={v[1]:5s} min {t[0]:.3f} med {med:.3f} ms  {fl / t[0] / 1e9:7.1f} / {fl / med / 1e9:7.1f} TFLOP/s", flush=True)
