"""SM clock and board power while one kernel runs back to back for ~2 s (nvidia-smi sampled every 10 ms in the background):
the evidence behind "the GPU is power-capped" in DESIGN 5.  Usage: python tools/gpu_clock_probe.py [gemm|attn_fwd|attn_bwd|add]"""
import os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kfunca_b200 as kf
from kfunca_b200.runtime import Event

what = sys.argv[1] if len(sys.argv) > 1 else "gemm"
if what == "gemm":
    n = 8192
    A = kf.empty([n, n], kf.bfloat16, 0); B = kf.empty([n, n], kf.bfloat16, 0)
    A.random_uniform_(1, -1.0, 1.0); B.random_uniform_(2, -1.0, 1.0)
    fn, flop = (lambda: kf.gemm(A, B, 1.0, 0.0)), 2.0 * n ** 3
elif what == "add":
    A = kf.empty([8192, 8192], kf.float, 0); B = kf.empty([8192, 8192], kf.float, 0)
    A.random_uniform_(1, -1.0, 1.0); B.random_uniform_(2, -1.0, 1.0)
    fn, flop = (lambda: A + B), 0.0
else:
    Bq, H, S, D = 8, 32, 4096, 128
    q, k, v, do = (kf.empty([Bq, H, S, D], kf.bfloat16, 0) for _ in range(4))
    for i, t in enumerate((q, k, v, do)): t.random_uniform_(10 + i, -1.0, 1.0)
    o, lse = kf.causal_attention_fwd(q, k, v)
    f = 4.0 * Bq * H * S * S * D / 2
    if what == "attn_fwd": fn, flop = (lambda: kf.causal_attention_fwd(q, k, v)), f
    else: fn, flop = (lambda: kf.causal_attention_bwd(do, q, k, v, o, lse)), 2.5 * f
for _ in range(3): fn()
kf.synchronize()
time.sleep(1.0)  # idle: let the clock recover
smi = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "10"], stdout=subprocess.PIPE, text=True)
time.sleep(0.3)
t0 = time.time()
evs = []
while time.time() - t0 < 2.0:
    e0, e1 = Event(), Event(); e0.record()
    for _ in range(8): fn()
    e1.record(); e1.synchronize()
    evs.append((time.time() - t0, e0.elapsed_ms(e1) / 8))
time.sleep(0.2)
smi.terminate()
rows = [l.strip().split(",") for l in smi.stdout.read().strip().splitlines() if "," in l]
clk = [int(r[0]) for r in rows]; pw = [float(r[1]) for r in rows]
print(f"{what}: {len(rows)} nvidia-smi samples: SM clock min {min(clk)} / median {sorted(clk)[len(clk)//2]} / max {max(clk)} MHz, power min {min(pw):.0f} / median {sorted(pw)[len(pw)//2]:.0f} / max {max(pw):.0f} W")
for lo in (0.0, 0.25, 0.5, 1.0, 1.5):
    seg = [ms for (t, ms) in evs if lo <= t < lo + 0.25]
    if seg:
        ms = sum(seg) / len(seg)
        print(f"  t in [{lo:.2f}, {lo + 0.25:.2f}) s: {ms:.4f} ms per call" + (f" = {flop / ms / 1e9:.0f} TFLOP/s" if flop else ""))
print("  clock samples (every 10th):", clk[::10])
print("  power samples (every 10th):", [round(x) for x in pw[::10]])
