"""Why is the C5 block slower inside bench.py's extras than alone?  Replays the extras sequence, then times single block steps
with pool statistics (in_use, reserved, arena mallocs) after each."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import kfunca_b200 as kf
from kfunca_b200.runtime import Event
from kfunca_b200.block import Block
import bench

peaks = bench.measured_peaks()
which = sys.argv[1] if len(sys.argv) > 1 else "extras"
print("mem before:", kf.mem_stats(), flush=True)
if which == "extras":
    real = bench.time_block
    bench.time_block = lambda *a, **k: {}
    r = bench.extras(kf, Event, peaks)
    bench.time_block = real
    print({k: v.get("ms") for k, v in r.items() if isinstance(v, dict)})
print("mem after extras:", kf.mem_stats(), flush=True)
B, S, E, H = 8, 4096, 4096, 32
blk = Block(E, H, dtype=kf.bfloat16, device=0, seed=7)
x = kf.from_numpy(np.random.default_rng(100).uniform(-1, 1, (B, S, E)).astype(np.float32), 0).to(kf.bfloat16)
import time
for i in range(8):
    e0, e1 = Event(), Event()
    t0 = time.perf_counter()
    e0.record()
    blk.step(x)
    t1 = time.perf_counter()
    e1.record()
    e1.synchronize()
    print(f"step {i}: {e0.elapsed_ms(e1):8.3f} ms (host issue {1e3 * (t1 - t0):7.2f} ms)  mem {kf.mem_stats()}", flush=True)
