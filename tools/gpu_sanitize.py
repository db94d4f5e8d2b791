"""Small-shape pass over the kernels written or changed in the last session, meant to run under compute-sanitizer
(`compute-sanitizer --tool memcheck|racecheck|synccheck python tools/gpu_sanitize.py`): persistent / per-CTA / 64-key-block forward,
fp32 tensor-core attention, ring layer-norm backward, column moments, PDL-launched elementwise / transpose / norm kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import kfunca_b200 as kf
from oracle import oracle as O

rng = np.random.default_rng(1)
g = lambda a: kf.from_numpy(a, 0)
b16 = lambda x: x.astype(np.float32).astype(O.bfloat16)
skip = set(os.environ.get("KF_SAN_SKIP", "").split(","))  # synccheck flags (and aborts on) the k64 kernel's per-block pv_done commits that nobody waits for
for mode in [m for m in ("pers", "cta", "k64") if m not in skip]:
    os.environ["KF_ATTN_FWD"] = mode
    for (b, h, sq, skv, d) in [(1, 3, 300, 300, 128), (2, 2, 130, 257, 64)]:
        q, k, v = (b16(rng.uniform(-1, 1, s)) for s in ((b, h, sq, d), (b, h, skv, d), (b, h, skv, d)))
        o, l = kf.causal_attention_fwd(g(q), g(k), g(v))
        err = np.abs(o.float().numpy() - O.causal_attention(q, k, v)).max()
        print(mode, (b, h, sq, skv, d), "max err", err, flush=True)
os.environ.pop("KF_ATTN_FWD")
for (b, h, sq, skv, d) in [(1, 2, 300, 300, 128), (1, 3, 200, 333, 64)]:
    q, k, v = (rng.uniform(-1, 1, s).astype(np.float32) for s in ((b, h, sq, d), (b, h, skv, d), (b, h, skv, d)))
    o, l = kf.causal_attention_fwd(g(q), g(k), g(v))
    print("fp32 tensor-core", (b, h, sq, skv, d), "max err", np.abs(o.numpy() - O.causal_attention(q, k, v)).max(), flush=True)
for dt, rows, E in ((np.float32, 700, 4096), (np.float32, 333, 1000), ("bf16", 900, 4096), ("bf16", 50, 512)):
    x = rng.uniform(-3, 3, (rows, E)).astype(np.float32)
    gn = rng.uniform(0.5, 1.5, (1, E)).astype(np.float32)
    dy = rng.uniform(-1, 1, (rows, E)).astype(np.float32)
    cv = (lambda a: a) if dt is np.float32 else b16
    gx, gg, gd = g(cv(x)), g(cv(gn)), g(cv(dy))
    gx.set_requires_grad(True)
    gg.set_requires_grad(True)
    y = kf.layer_norm(gx, gg, 1e-5)
    y.backward(gd)
    print("layer norm fwd+bwd", dt if isinstance(dt, str) else "fp32", (rows, E), float(np.abs(gx.grad().float().numpy()).max()), flush=True)
a = rng.uniform(-10, 10, (1000, 520)).astype(np.float32)
ga = g(a)
m, r = ga.norm_stat(0)
print("norm_stat", np.abs(m.numpy() - a.mean(0, keepdims=True)).max(), flush=True)
mv = ga.mean_var(1, False)
(ga + ga).sum(0).numpy(); (ga * ga).permute(1, 0).contiguous().sum(1).numpy()
big = g(rng.uniform(-1, 1, (1 << 22,)).astype(np.float32))
print("full sum (two-step)", float(big.sum(0).numpy()[0]), flush=True)
kf.synchronize()
print("sanitize workload done")
