"""N-GPU check of the data-parallel block (run under torchrun): every rank steps its batch shard, gradients go through the NCCL
all-reduce (AVG) on the library stream, and rank 0 compares them with a full-batch step it computes alone.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/gpu_dist_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import kfunca_b200 as kf
from kfunca_b200.block import Block
from kfunca_b200.dist import all_reduce_grads, all_reduce_mean_scalar, shard_bounds

kf.set_device(local)
B, S, E, H = 2 * world, 256, 256, 2
x_all = np.random.default_rng(5).uniform(-1, 1, (B, S, E)).astype(np.float32)
lo, hi = shard_bounds(B, rank, world)
blk = Block(E, H, dtype=kf.bfloat16, device=local, seed=3)
x = kf.from_numpy(x_all[lo:hi], local).to(kf.bfloat16)
loss = blk.step(x)
all_reduce_grads(blk.params, world, dist)
gl = all_reduce_mean_scalar(loss, world, dist)
kf.synchronize()
got = {n: p.grad().float().numpy().astype(np.float64) for n, p in blk.params.items()}
got_loss = float(gl.float().numpy().reshape(-1)[0])
ok = True
if rank == 0:
    ref = Block(E, H, dtype=kf.bfloat16, device=local, seed=3)
    lf = ref.step(kf.from_numpy(x_all, local).to(kf.bfloat16))
    kf.synchronize()
    want_loss = float(lf.float().numpy().reshape(-1)[0])
    print(f"loss: sharded {got_loss:.6e}  full-batch {want_loss:.6e}")
    ok &= abs(got_loss - want_loss) <= 3e-2 * max(abs(want_loss), 1e-6) + 1e-6
    for n, p in ref.params.items():
        w = p.grad().float().numpy().astype(np.float64)
        err = np.abs(got[n] - w).max() / max(np.abs(w).max(), 1e-30)
        print(f"grad {n:5s}: max rel err vs full batch {err:.3e}")
        ok &= err <= 3e-2
    print("DIST CHECK", "OK" if ok else "FAILED")
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
