"""Forward attention A/B in one process: persistent kernel (KF_ATTN_FWD=pers; the default for KV lengths >= 8192) against the one-CTA-per-pair kernel (KF_ATTN_FWD=cta).
Parity: both against the oracle and bit-for-bit against each other on ragged / rectangular / D = 64 / fp16 shapes; timings at
S = 1024 ... 8192 with B H S constant; `--trace` prints CTA 0's pipeline stamps of the persistent kernel at C3."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import kfunca_b200 as kf
from kfunca_b200.runtime import Event
from oracle import oracle as O

rng = np.random.default_rng(5)
g = lambda a: kf.from_numpy(a, 0)
b16 = lambda x: x.astype(np.float32).astype(O.bfloat16)
f16 = lambda x: x.astype(np.float16)


def run(mode, q, k, v):
    if mode is None:
        os.environ.pop("KF_ATTN_FWD", None)
    else:
        os.environ["KF_ATTN_FWD"] = mode
    out, lse = kf.causal_attention_fwd(q, k, v)
    return out.float().numpy(), lse.numpy()


TRACE_ONLY = "--trace-only" in sys.argv  # run with KF_ATTN_TRACE=1 in the environment (read once per process)
bad = 0
for (b, h, sq, skv, d, cv) in [] if TRACE_ONLY else [(1, 2, 256, 256, 128, b16), (1, 1, 200, 333, 128, b16), (1, 2, 1024, 1024, 128, b16), (2, 2, 640, 640, 64, b16),
                               (3, 5, 130, 130, 128, f16), (2, 3, 1, 1, 128, b16), (1, 160, 384, 384, 128, b16), (2, 70, 513, 257, 64, f16),
                               (1, 300, 100, 700, 128, b16), (1, 2, 4096, 4096, 128, b16)]:
    q, k, v = (cv(rng.uniform(-1, 1, s)) for s in ((b, h, sq, d), (b, h, skv, d), (b, h, skv, d)))
    gq, gk, gv = g(q), g(k), g(v)
    o_p, l_p = run("pers", gq, gk, gv)
    o_c, l_c = run("cta", gq, gk, gv)
    same = np.array_equal(o_p, o_c) and np.array_equal(l_p, l_c)
    o_k, l_k = run("k64", gq, gk, gv)
    dk = float(np.abs(o_k - o_c).max())
    dl = float(np.abs(l_k - l_c).max())
    msg = f"{b}x{h}x{sq}x{skv}x{d} {cv.__name__}: persistent == per-CTA bitwise: {same}  k64 vs per-CTA max |d out| {dk:.3g} |d lse| {dl:.3g}"
    if not (dk < 2e-2 and dl < 1e-3):
        bad += 1
    if b * h * sq * skv <= 4 << 20:
        eo = O.causal_attention(q, k, v)
        err = np.abs(o_p - eo).max()
        msg += f"  max_abs_err vs oracle {err:.3g}"
        if not err < 1e-2:
            bad += 1
    if not same:
        bad += 1
    print(msg, flush=True)
if not TRACE_ONLY:
    print("PARITY", "FAIL" if bad else "OK", flush=True)


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = Event(), Event()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_ms(e1) / iters


H, D = 32, 128
for S in () if TRACE_ONLY else (1024, 2048, 4096, 8192, 16384):
    B = max(1, 8 * 4096 // S)
    mk = lambda: g(b16(rng.uniform(-1, 1, (1, H, S, D))))
    Q, K, V = (kf.cat([t] * B, 0) if B > 1 else t for t in (mk(), mk(), mk()))
    fl = 4 * B * H * S * S * D / 2
    # the two kernels alternate call by call (the GPU is power-capped: clocks drift by 10 - 20 % over tens of milliseconds, so
    # back-to-back windows of one variant each are not comparable); per variant: minimum and median over the calls
    times = {f"{m}:{hg}": [] for m in ("pers", "cta", "k64") for hg in (("auto",) if "--hg" not in sys.argv else ("1", "auto", "all"))}  # heads per scheduling group (KF_ATTN_HG)
    evs = [Event() for _ in range(3)]
    for rep in range(16):
        for mode in times:
            os.environ["KF_ATTN_FWD"] = mode.split(":")[0]
            hg = mode.split(":")[1]
            if hg == "auto":
                os.environ.pop("KF_ATTN_HG", None)
            else:
                os.environ["KF_ATTN_HG"] = "1" if hg == "1" else "1000000"
            if rep < 2:
                kf.causal_attention_fwd(Q, K, V)
                continue
            evs[0].record()
            kf.causal_attention_fwd(Q, K, V)
            evs[1].record()
            evs[1].synchronize()
            times[mode].append(evs[0].elapsed_ms(evs[1]))
    res = [f"{m} min {min(t):.3f} med {sorted(t)[len(t) // 2]:.3f} ms {fl / sorted(t)[len(t) // 2] / 1e9:7.1f} TFLOP/s" for m, t in times.items()]
    print(f"B={B} S={S}:\n    " + "\n    ".join(res), flush=True)
    os.environ.pop("KF_ATTN_HG", None)
    if "--all" in sys.argv:
        for m, t in times.items():
            print(f"    {m}: " + " ".join(f"{x:.3f}" for x in t), flush=True)
    del Q, K, V

if TRACE_ONLY:
    os.environ["KF_ATTN_FWD"] = "pers"
    B, S = 8, 4096
    mk = lambda: g(b16(rng.uniform(-1, 1, (1, H, S, D))))
    Q, K, V = (kf.cat([t] * B, 0) for t in (mk(), mk(), mk()))
    kf.causal_attention_fwd(Q, K, V)
