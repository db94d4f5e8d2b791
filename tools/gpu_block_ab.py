"""A/B of the C5 block step at one GPU: fused layer norm vs the composed one, interleaved in one process (same box, same
thermal state) so box-to-box power-cap differences cancel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import kfunca_b200 as kf
from kfunca_b200.runtime import Event
from kfunca_b200.block import Block

B, S, E, H = 8, 4096, 4096, 32
x = kf.from_numpy(np.random.default_rng(100).uniform(-1, 1, (B, S, E)).astype(np.float32), 0).to(kf.bfloat16)
blocks = {"fused": Block(E, H, dtype=kf.bfloat16, device=0, seed=7, fused_norm=True),
          "composed": Block(E, H, dtype=kf.bfloat16, device=0, seed=7, fused_norm=False)}


def run(blk, n):
    e0, e1 = Event(), Event()
    e0.record()
    for _ in range(n):
        blk.step(x)
    e1.record()
    e1.synchronize()
    return e0.elapsed_ms(e1) / n


for b in blocks.values():
    run(b, 2)
res = {k: [] for k in blocks}
for rep in range(4):
    for k, b in blocks.items():
        res[k].append(run(b, 3))
for k, v in res.items():
    print(f"{k:9s} ms/step per round {[round(t, 2) for t in v]}  median {sorted(v)[len(v) // 2]:.2f}")
