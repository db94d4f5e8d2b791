"""Data-parallel plumbing: one process per GPU, NCCL over NVLink through the library's own native layer
(csrc/dist.cpp, C ABI kf_dist_*).  No torch anywhere on this path: the rendezvous is a few lines of std-lib TCP (rank 0 hands the
128-byte NCCL id to the other ranks on MASTER_ADDR), the collectives are ncclAllReduce calls issued by the library on its own
streams, and the gradient all-reduce is started from inside backward() by a native leaf-gradient hook.
Only two exchanges exist on this path: the weight-gradient all-reduce and the cross-shard full-tensor sum/mean."""
from __future__ import annotations

import os
import socket
import struct
import time

import kfunca_b200 as kf

SUM, AVG, MAX = 0, 1, 2
_ID_BYTES = 128
_PORT_OFFSET = 137  # the id exchange listens on MASTER_PORT + this (MASTER_PORT itself belongs to the launcher's store)


def shard_bounds(total: int, rank: int, world: int) -> tuple[int, int]:
    """Rows / samples [lo, hi) owned by `rank` when `total` independent units are split over `world` ranks
    (rank r owns [r*B/n, (r+1)*B/n)); the first total % world ranks take one extra unit."""
    if not (0 <= rank < world) or total < 0:
        raise ValueError(f"bad shard request total={total} rank={rank} world={world}")
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def global_mean_from_partials(local_sum: float, local_count: float, all_reduce_sum) -> float:
    """Cross-shard mean = all-reduce(sum of local sums) / all-reduce(sum of local counts) — NOT the mean of per-shard means,
    which is wrong for unequal shards.  `all_reduce_sum(list[float]) -> list[float]` is the collective (NCCL through
    kf.dist_all_reduce on the GPU path, anything else in the CPU tests)."""
    s, n = all_reduce_sum([float(local_sum), float(local_count)])
    return s / n


def exchange_id(payload: bytes | None, rank: int, world: int, addr: str, port: int, timeout: float = 120.0) -> bytes:
    """rank 0 serves `payload` (the NCCL unique id) to the other world - 1 ranks over TCP; they return what they received."""
    if world == 1:
        return payload
    if rank == 0:
        assert payload is not None and len(payload) == _ID_BYTES
        srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        srv.bind((addr, port))
        srv.listen(world)
        srv.settimeout(timeout)
        seen = set()
        try:
            while len(seen) < world - 1:
                conn, _ = srv.accept()
                with conn:
                    conn.settimeout(timeout)
                    (peer,) = struct.unpack("<i", _recv_exact(conn, 4))
                    conn.sendall(payload)
                    seen.add(peer)
        finally:
            srv.close()
        return payload
    deadline = time.time() + timeout
    while True:
        try:
            with socket.create_connection((addr, port), timeout=5.0) as c:
                c.sendall(struct.pack("<i", rank))
                return _recv_exact(c, _ID_BYTES)
        except OSError:
            if time.time() > deadline:
                raise
            time.sleep(0.05)


def _recv_exact(conn, n: int) -> bytes:
    buf = b""
    while len(buf) < n:
        chunk = conn.recv(n - len(buf))
        if not chunk:
            raise ConnectionError("peer closed during the id exchange")
        buf += chunk
    return buf


def init_from_env() -> tuple[int, int, int]:
    """(rank, world, local_rank) from the launcher's environment (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_ADDR / MASTER_PORT);
    selects the GPU, exchanges the NCCL id and creates the communicator.  world == 1 needs no launcher and loads no NCCL."""
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    kf.set_device(local)
    if world > 1:
        bind_to_gpu_numa_node()
    if world > 1 and not kf.dist_info()[0]:
        addr = os.environ.get("MASTER_ADDR", "127.0.0.1")
        port = int(os.environ.get("MASTER_PORT", "29500")) + _PORT_OFFSET
        uid = exchange_id(kf.dist_unique_id() if rank == 0 else None, rank, world, addr, port)
        kf.dist_init(uid, rank, world)
    return rank, world, local


def _parse_cpulist(s: str) -> set[int]:
    cpus: set[int] = set()
    for part in filter(None, s.strip().split(",")):
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa_node() -> int:
    """rank -> core affinity: run this process (and the driver / NCCL proxy threads it spawns later) on the CPUs of its GPU's NUMA
    node, intersected with the cpuset the launcher gave it.  Returns the node (-1: unknown, nothing changed).  KF_NUMA=0 disables."""
    if os.environ.get("KF_NUMA", "1") == "0" or not hasattr(os, "sched_setaffinity"):
        return -1
    try:
        node, cpulist = kf.numa_info()
        want = _parse_cpulist(cpulist) & os.sched_getaffinity(0)
        if node >= 0 and want:
            os.sched_setaffinity(0, want)
            return node
    except (OSError, ValueError, RuntimeError):
        pass
    return -1


def finalize() -> None:
    kf.dist_finalize()


def all_reduce_(t, op: int = SUM):
    """in place on the library stream (ordered with the kernels around it); no-op at world 1"""
    return kf.dist_all_reduce(t, op)


def barrier() -> None:
    """every rank has enqueued and finished everything before this point"""
    if kf.dist_info()[0]:
        tok = kf.zeros([1], kf.float, kf.get_device())
        kf.dist_all_reduce(tok, SUM)
    kf.synchronize()


def max_over_ranks(value: float) -> float:
    if not kf.dist_info()[0]:
        return value
    import numpy as np

    t = kf.from_numpy(np.array([value], dtype=np.float64), kf.get_device())
    kf.dist_all_reduce(t, MAX)
    return float(t.numpy()[0])


def all_reduce_grads(params) -> None:
    """average every parameter gradient over the data-parallel group AFTER the backward pass: one ncclAllReduce(AVG) per
    parameter, back to back on the library stream (the non-overlapped baseline of OverlappedGradAllReduce)."""
    if not kf.dist_info()[0]:
        return
    for p in params.values():
        g = p.grad()
        if g.defined():
            kf.dist_all_reduce(g, AVG)


def all_reduce_mean_scalar(t):
    """cross-shard mean of per-shard means with equal shard sizes: all-reduce(AVG) of an fp32 copy"""
    if not kf.dist_info()[0]:
        return t
    f = t.float()
    kf.dist_all_reduce(f, AVG)
    return f


class OverlappedGradAllReduce:
    """Start the all-reduce (AVG) of every parameter gradient the moment the backward pass has enqueued it, on the library's
    communication stream, so NCCL over NVLink runs under the rest of the backward pass instead of after it.

        with OverlappedGradAllReduce(params):
            loss = block.step(x)          # backward() fires the native leaf-gradient hook once per parameter
        # on exit the library stream waits for the communication stream: gradients are averaged for whoever reads them next

    Everything between the hook and the collective is native (csrc/dist.cpp): event on the library stream, wait on the
    communication stream, ncclAllReduce.  zero_grad() the parameters before each backward inside the context (block.step does)."""

    def __init__(self, params):
        self.params = list(params.values())
        self.count = 0

    def __enter__(self):
        kf.dist_overlap_begin(self.params)
        return self

    def __exit__(self, *exc):
        self.count = kf.dist_overlap_end()
        return False
