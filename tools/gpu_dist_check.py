"""N-GPU check of the data-parallel block (run under torchrun, which only provides RANK / WORLD_SIZE / MASTER_*): every rank steps
its batch shard, gradients go through the library's native NCCL all-reduce (AVG) — once after the backward pass and once
overlapped with it — and rank 0 compares them with a full-batch step it computes alone.  Also checks the cross-shard mean.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/gpu_dist_check.py [E S]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import kfunca_b200 as kf
from kfunca_b200 import dist as kd
from kfunca_b200.block import Block

rank, world, local = kd.init_from_env()
E = int(sys.argv[1]) if len(sys.argv) > 1 else 256
S = int(sys.argv[2]) if len(sys.argv) > 2 else 256
B, H = 2 * world, max(2, E // 128)
x_all = np.random.default_rng(5).uniform(-1, 1, (B, S, E)).astype(np.float32)
lo, hi = kd.shard_bounds(B, rank, world)
blk = Block(E, H, dtype=kf.bfloat16, device=local, seed=3)
x = kf.from_numpy(x_all[lo:hi], local).to(kf.bfloat16)
results = {}
for mode in ("after", "overlapped"):
    if mode == "after":
        loss = blk.step(x)
        kd.all_reduce_grads(blk.params)
    else:
        with kd.OverlappedGradAllReduce(blk.params) as ar:
            loss = blk.step(x)
        assert ar.count == len(blk.params), (ar.count, len(blk.params))
    gl = kd.all_reduce_mean_scalar(loss)
    kf.synchronize()
    results[mode] = ({n: p.grad().float().numpy().astype(np.float64) for n, p in blk.params.items()}, float(gl.float().numpy().reshape(-1)[0]))
ok = True
# both schedules must give bit-identical averaged gradients (same collectives, same operands)
for n in results["after"][0]:
    ok &= bool(np.array_equal(results["after"][0][n], results["overlapped"][0][n]))
if rank == 0:
    print("nccl", kf.dist_info(), "overlapped == after:", ok)
    ref = Block(E, H, dtype=kf.bfloat16, device=local, seed=3)
    lf = ref.step(kf.from_numpy(x_all, local).to(kf.bfloat16))
    kf.synchronize()
    want_loss = float(lf.float().numpy().reshape(-1)[0])
    got, got_loss = results["overlapped"]
    print(f"loss: sharded {got_loss:.6e}  full-batch {want_loss:.6e}")
    ok &= abs(got_loss - want_loss) <= 3e-2 * max(abs(want_loss), 1e-6) + 1e-6
    for n, p in ref.params.items():
        w = p.grad().float().numpy().astype(np.float64)
        err = np.abs(got[n] - w).max() / max(np.abs(w).max(), 1e-30)
        print(f"grad {n:5s}: max rel err vs full batch {err:.3e}")
        ok &= err <= 3e-2
    print("DIST CHECK", "OK" if ok else "FAILED")
kd.barrier()
kd.finalize()
sys.exit(0 if ok else 1)
