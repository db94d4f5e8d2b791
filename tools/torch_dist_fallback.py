"""FALLBACK ONLY — the round-1 torch.distributed plumbing, kept outside the product package for one purpose: if the native NCCL
layer (kfunca_b200/dist.py -> csrc/dist.cpp) cannot initialise on some box, `bench.py --gpus N` still produces its numbers through
this module and says so in the JSON line (`dist_backend: "torch.distributed (fallback)"`).  Same interface as kfunca_b200.dist."""
from __future__ import annotations

import os


import numpy as np

import kfunca_b200 as kf

_TYPESTR = {kf.float: "<f4", kf.double: "<f8", kf.half: "<f2", kf.bfloat16: "<i2", kf.int: "<i4", kf.long: "<i8",
            kf.short: "<i2", kf.char: "|i1", kf.byte: "|u1", kf.bool: "|b1"}


class _CAI:
    def __init__(self, t):
        assert t.is_contiguous()
        self.__cuda_array_interface__ = {"shape": tuple(t.sizes()), "typestr": _TYPESTR[t.dtype()], "data": (t.data_ptr(), False),
                                         "version": 2, "strides": None}
        self._keep = t


def as_torch(t):
    """Zero-copy torch alias of a contiguous kfunca_b200 tensor (bf16 travels as int16 and is re-viewed)."""
    import torch

    out = torch.as_tensor(_CAI(t), device=f"cuda:{t.device()}")
    if t.dtype() == kf.bfloat16:
        out = out.view(torch.bfloat16)
    return out


def library_stream():
    """torch view of the library's compute stream, so collectives are ordered with our kernels."""
    import torch

    return torch.cuda.ExternalStream(kf.stream())


def shard_bounds(total: int, rank: int, world: int) -> tuple[int, int]:
    """Rows / samples [lo, hi) owned by `rank` when `total` independent units are split over `world` ranks
    (rank r owns [r*B/n, (r+1)*B/n)); the first total % world ranks take one extra unit."""
    if not (0 <= rank < world) or total < 0:
        raise ValueError(f"bad shard request total={total} rank={rank} world={world}")
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def global_mean_from_partials(local_sum, local_count: int, dist):
    """Cross-shard mean = all-reduce(sum of local sums) / all-reduce(sum of local counts) — NOT the mean of
    per-shard means, which is wrong for unequal shards.  `local_sum` is a torch tensor on any
    device the process group's backend supports (NCCL: cuda, gloo: cpu); it is reduced in place."""
    import torch

    cnt = torch.tensor([float(local_count)], dtype=torch.float64, device=local_sum.device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(local_sum)
        dist.all_reduce(cnt)
    return local_sum / cnt.to(local_sum.dtype)


def average_gradients_(grads, dist) -> None:
    """In-place sum-all-reduce of a list of torch gradient tensors followed by 1/world (data-parallel average)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return
    world = dist.get_world_size()
    for g in grads:
        dist.all_reduce(g)
        g.mul_(1.0 / world)


def all_reduce_grads(params, world: int, dist) -> None:
    """average every parameter gradient over the data-parallel group: one NCCL all-reduce per parameter with the AVG
    reduction (the 1/world scale happens inside the collective, no extra pass over the gradients), issued back to back on
    the library stream so they queue behind the backward kernels that produced them."""
    import torch

    if world == 1:
        return
    avg = dist.ReduceOp.AVG if dist.get_backend() == "nccl" else None
    with torch.cuda.stream(library_stream()):
        for p in params.values():
            g = p.grad()
            if not g.defined():
                continue
            tg = as_torch(g)
            if avg is not None:
                dist.all_reduce(tg, op=avg)
            else:  # gloo has no AVG
                dist.all_reduce(tg)
                g *= 1.0 / world


def all_reduce_mean_scalar(t, world: int, dist):
    """cross-shard mean of per-shard means with equal shard sizes: all-reduce(AVG)"""
    import torch

    if world == 1:
        return t
    f = t.float()
    with torch.cuda.stream(library_stream()):
        if dist.get_backend() == "nccl":
            dist.all_reduce(as_torch(f), op=dist.ReduceOp.AVG)
        else:
            dist.all_reduce(as_torch(f))
            f *= 1.0 / world
    return f


class OverlappedGradAllReduce:
    """Start the all-reduce (AVG) of every parameter gradient the moment the backward pass has enqueued it, on a side stream, so
    NCCL over NVLink runs under the rest of the backward pass instead of after it.

        with OverlappedGradAllReduce(params, world, dist):
            loss = block.step(x)          # backward() fires the leaf-gradient hook once per parameter
        # on exit the library stream waits for the side stream: gradients are averaged for whoever reads them next

    Mechanics: the hook (kf.set_leaf_grad_hook) records an event on the library stream, the side stream waits for it and the
    collective is issued there.  Gradient memory is owned by the parameter until the next zero_grad(), which is ordered after
    the join, so the pool's single-stream free rule still holds."""

    def __init__(self, params, world: int, dist):
        import torch

        self.world, self.dist, self.torch = world, dist, torch
        self.ptrs = {p.data_ptr() for p in params.values()}
        self.active = world > 1
        if self.active:
            self.lib = library_stream()
            self.side = torch.cuda.Stream()
            self.avg = dist.ReduceOp.AVG if dist.get_backend() == "nccl" else None
        self.count = 0

    def _hook(self, leaf, grad):
        if leaf.data_ptr() not in self.ptrs:
            return
        torch = self.torch
        ev = torch.cuda.Event()
        ev.record(self.lib)
        self.side.wait_event(ev)
        with torch.cuda.stream(self.side):
            tg = as_torch(grad)
            if self.avg is not None:
                self.dist.all_reduce(tg, op=self.avg)
            else:
                self.dist.all_reduce(tg)
                tg.mul_(1.0 / self.world)
        self.count += 1

    def __enter__(self):
        if self.active:
            self.count = 0
            kf.set_leaf_grad_hook(self._hook)
        return self

    def __exit__(self, *exc):
        if self.active:
            kf.set_leaf_grad_hook(None)
            self.lib.wait_stream(self.side)
        return False


# ---- the interface of kfunca_b200.dist, bound to one torch process group
_dist = None
_world = 1


def init_from_env():
    global _dist, _world
    import torch
    import torch.distributed as dist

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    kf.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    _dist, _world = dist, world
    return rank, world, local


def barrier():
    import torch

    _dist.barrier()
    torch.cuda.synchronize()
    kf.synchronize()


def max_over_ranks(value):
    import torch

    t = torch.tensor([value], dtype=torch.float64, device="cuda")
    _dist.all_reduce(t, op=_dist.ReduceOp.MAX)
    return float(t.item())


_all_reduce_grads3, _all_reduce_mean_scalar3, _Overlapped3 = all_reduce_grads, all_reduce_mean_scalar, OverlappedGradAllReduce


def all_reduce_grads(params):  # noqa: F811
    _all_reduce_grads3(params, _world, _dist)


def all_reduce_mean_scalar(t):  # noqa: F811
    return _all_reduce_mean_scalar3(t, _world, _dist)


class OverlappedGradAllReduce(_Overlapped3):  # noqa: F811
    def __init__(self, params):
        super().__init__(params, _world, _dist)
