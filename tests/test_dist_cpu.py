"""N>1 host logic on CPU, world_size 2.  The product's data-parallel layer is native NCCL (csrc/dist.cpp) and needs GPUs; what
runs here is everything around it that does not: the TCP exchange of the NCCL id between two real processes, the shard
arithmetic, and the cross-shard mean / gradient-average formulas driven through a `gloo` all-reduce in place of ncclAllReduce.
kfunca_b200.dist imports the compiled extension at module import, but nothing here launches a kernel."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total_rows, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from kfunca_b200.dist import exchange_id, global_mean_from_partials, shard_bounds

        # the rendezvous the GPU path uses for the NCCL unique id: rank 0 serves 128 bytes, the others fetch them
        payload = bytes(range(128)) if rank == 0 else None
        got = exchange_id(payload, rank, world, "127.0.0.1", port + 1)

        def all_reduce_sum(vals):  # gloo stands in for ncclAllReduce(sum)
            t = torch.tensor(vals, dtype=torch.float64)
            dist.all_reduce(t)
            return t.tolist()

        rng = np.random.default_rng(7)
        x = rng.uniform(-10, 10, (total_rows, 33))  # every rank generates the same global batch
        lo, hi = shard_bounds(total_rows, rank, world)
        local = x[lo:hi]
        mean = global_mean_from_partials(local.sum(), local.size, all_reduce_sum)
        g = all_reduce_sum([float(rank + 1)] * 5)
        q.put((rank, lo, hi, float(mean), [v / world for v in g], got == bytes(range(128))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total_rows", [8, 7])  # equal and ragged shards
def test_two_rank_shard_mean_and_grad_average(total_rows):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total_rows, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(7)
    x = rng.uniform(-10, 10, (total_rows, 33))
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == total_rows  # shards tile [0, total)
    for _, _, _, mean, g, id_ok in res:
        assert abs(mean - x.mean()) < 1e-12
        assert g == [1.5] * 5
        assert id_ok


def test_shard_bounds_cover_without_overlap():
    from kfunca_b200.dist import shard_bounds

    for total in (0, 1, 7, 8, 65536):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    with pytest.raises(ValueError):
        shard_bounds(8, 2, 2)


def test_dist_layer_is_inert_at_world_one():
    """single-process use never loads NCCL: dist_info reports 'not initialised' and the C ABI exports the whole layer"""
    import ctypes

    import kfunca_b200 as kf

    assert kf.dist_info()[0] is False
    lib = ctypes.CDLL(kf.LIB_PATH)
    for name in ("kf_dist_unique_id", "kf_dist_init", "kf_dist_finalize", "kf_dist_info", "kf_dist_all_reduce", "kf_dist_overlap_begin",
                 "kf_dist_overlap_end"):
        assert hasattr(lib, name)
    import kfunca_b200.dist as d

    src = open(d.__file__).read()
    assert "import torch" not in src  # PyTorch is not a runtime dependency of the product's multi-GPU path


def test_cpulist_parser_for_numa_binding():
    from kfunca_b200.dist import _parse_cpulist
    assert _parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert _parse_cpulist("") == set()
    assert _parse_cpulist("5") == {5}
