import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
import kfunca_b200 as kf
from kfunca_b200.runtime import Event
B, H, S, D = 8, 32, 4096, 128
if len(sys.argv) > 1:
    S = int(sys.argv[1]); B = 8 * 4096 // S
q, k, v = (kf.empty([B, H, S, D], kf.bfloat16, 0) for _ in range(3))
for i, t in enumerate((q, k, v)): t.random_uniform_(10 + i, -1.0, 1.0)
fl = 4 * B * H * S * S * D / 2
os.environ["KF_ATTN_FWD"] = "cta"
hgs = ["1", "4", "8", "16", "32", "64", "1000000"]
times = {h: [] for h in hgs}
for rep in range(14):
    for h in hgs:
        os.environ["KF_ATTN_HG"] = h
        e0, e1 = Event(), Event(); e0.record()
        kf.causal_attention_fwd(q, k, v)
        e1.record(); e1.synchronize()
        if rep >= 2: times[h].append(e0.elapsed_ms(e1))
for h, t in times.items():
    t.sort(); print(f"S={S} cta hg={h:8s} min {t[0]:.3f} med {t[len(t)//2]:.3f} ms", flush=True)
