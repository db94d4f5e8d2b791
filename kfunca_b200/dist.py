"""Data-parallel plumbing: one process per GPU, NCCL over NVLink through torch.distributed.
Only two exchanges exist on this path: the weight-gradient all-reduce and the cross-shard
full-tensor sum/mean.  Device memory stays owned by kfunca_b200's pool; torch only sees aliases of it."""
from __future__ import annotations

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


def all_reduce_grads(params, world: int, dist) -> None:
    """sum-all-reduce every parameter gradient over the data-parallel group, then scale by 1/world."""
    import torch

    if world == 1:
        return
    with torch.cuda.stream(library_stream()):
        for p in params.values():
            g = p.grad()
            if not g.defined():
                continue
            tg = as_torch(g)
            dist.all_reduce(tg)
            g *= 1.0 / world


def all_reduce_mean_scalar(t, world: int, dist):
    """cross-shard mean of per-shard means with equal shard sizes: all-reduce(sum) / world"""
    import torch

    if world == 1:
        return t
    f = t.float()
    with torch.cuda.stream(library_stream()):
        dist.all_reduce(as_torch(f))
    f *= 1.0 / world
    return f
