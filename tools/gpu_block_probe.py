"""C5 block step time at one GPU (global batch 8) — quick probe; env hooks are read per call so variants can be compared
in one process.  `--once` runs one warm step + one step (for an ncu launch list)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import kfunca_b200 as kf
from kfunca_b200.runtime import Event
from kfunca_b200.block import Block

B, S, E, H = int(os.environ.get("KF_PROBE_B", 8)), 4096, 4096, 32
blk = Block(E, H, dtype=kf.bfloat16, device=0, seed=7)
x = kf.from_numpy(np.random.default_rng(100).uniform(-1, 1, (B, S, E)).astype(np.float32), 0).to(kf.bfloat16)


def run(n):
    e0, e1 = Event(), Event()
    e0.record()
    for _ in range(n):
        blk.step(x)
    e1.record()
    e1.synchronize()
    return e0.elapsed_ms(e1) / n


if "--once" in sys.argv:
    run(1)
    run(1)
    kf.synchronize()
    sys.exit(0)
for name, env in [("stream=1", {"KF_RED_STREAM": "1"}), ("stream=0", {"KF_RED_STREAM": "0"}), ("stream=0 pdl=0", {"KF_RED_STREAM": "0", "KF_PDL": "0"})]:
    for k in ("KF_RED_STREAM", "KF_PDL"):
        os.environ.pop(k, None)
    os.environ.update(env)
    run(2)
    print(f"{name}: {run(3):.3f} ms/step  ({blk.flops_per_sample(S) * B / run(3) / 1e9:.1f} TFLOP/s)", flush=True)
