"""Layer-norm backward A/B in one process (fp32 and bf16, 4096 x 4096 and the block's 32768 x 4096): the one-pass kernel with
128 threads x 8 vectors (KF_LN_BWD=narrow) against 256 threads x 4 vectors (default), resident CTAs per SM 2 / 3 (KF_LN_BWD_CTAS);
the variants alternate call by call, median over the calls; dx and the gain gradient are compared bit for bit between variants."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import kfunca_b200 as kf
from kfunca_b200.runtime import Event

rng = np.random.default_rng(3)
NSETS = 4


def run_case(rows, E, dt, name):
    xs = [kf.empty([rows, E], dt, 0) for _ in range(NSETS)]
    gs = [kf.empty([rows, E], dt, 0) for _ in range(NSETS)]
    for i, (x, g) in enumerate(zip(xs, gs)):
        x.random_uniform_(10 + i, -3.0, 3.0)
        g.random_uniform_(20 + i, -1.0, 1.0)
        x.set_requires_grad(True)
    gain = kf.empty([1, E], dt, 0)
    gain.random_uniform_(7, 0.5, 1.5)
    gain.set_requires_grad(True)
    ys = [kf.layer_norm(x, gain, 1e-5) for x in xs]
    variants = {"narrow/2": ("narrow", "2"), "regs/2": ("regs", "2"), "ring/2": (None, "2")}
    times = {k: [] for k in variants}
    outs = {}
    for rep in range(26):
        for k, (mode, per_sm) in variants.items():
            if mode is None:
                os.environ.pop("KF_LN_BWD", None)
            else:
                os.environ["KF_LN_BWD"] = mode
            os.environ["KF_LN_BWD_CTAS"] = per_sm
            i = rep % NSETS
            e0, e1 = Event(), Event()
            e0.record()
            for j in range(NSETS):  # four calls back to back: the stream never waits for the host
                xs[j].zero_grad()
                gain.zero_grad()
                ys[j].backward(gs[j])
            e1.record()
            e1.synchronize()
            if rep >= 2:
                times[k].append(e0.elapsed_ms(e1) * 1e3 / NSETS)
            if rep == 0:
                outs[k] = (xs[0].grad().float().numpy().copy(), gain.grad().float().numpy().copy())
    ref = outs["narrow/2"]
    byts = 3 * rows * E * (4 if dt == kf.float else 2)
    for k, t in times.items():
        t.sort()
        med = t[len(t) // 2]
        same_dx = np.array_equal(outs[k][0], ref[0])
        dg_err = float(np.abs(outs[k][1] - ref[1]).max() / max(1e-30, np.abs(ref[1]).max()))
        print(f"{name} {k}: median {med:7.1f} us  min {t[0]:7.1f} us  {byts / med / 1e3:7.1f} GB/s (x, dy in + dx out)  dx identical: {same_dx}  dgain rel diff {dg_err:.2e}",
              flush=True)


run_case(4096, 4096, kf.float, "fp32 4096x4096")
run_case(4096, 4096, kf.bfloat16, "bf16 4096x4096")
run_case(32768, 4096, kf.bfloat16, "bf16 32768x4096")
