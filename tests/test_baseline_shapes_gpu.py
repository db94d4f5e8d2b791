"""GPU parity AT THE BASELINE.json SHAPES (VERDICT r1 "What's weak" #1): the full-size configs are generated on the device by the
counter-based generator (`tensor.random_uniform_`, host twin `oracle.counter_uniform`) and checked against the float64 / NumPy
oracle on samples that include the first and last tiles, the diagonal blocks and the rows beyond 2^31 elements.

  C2  bf16 matmul 8192^3          : 64 full rows and 64 full columns against the fp64 product (2e-2 relative, L1-mass floor)
  C3  causal attention fwd + bwd  : B=8 H=32 S=4096 D=128 bf16; three (b, h) heads checked in full (every query tile, every
                                    diagonal block) against the fp64 oracle, forward and dQ / dK / dV
  C4  top-k k=64 over 65536x32768 : values AND int64 indices bit-exact on 1024 sampled rows incl. rows 0 and 65535
  C5  transformer block, S=E=4096 : sharded-vs-full gradient identity on one GPU (two 1-sample shards averaged == 2-sample step)
"""
import numpy as np
import pytest

import kfunca_b200 as kf
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def dev_uniform(shape, dtype, seed, lo=-1.0, hi=1.0):
    t = kf.empty(list(shape), dtype, 0)
    t.random_uniform_(seed, lo, hi)
    return t


def host_uniform_bf16(shape, seed, lo=-1.0, hi=1.0, start=0):
    n = int(np.prod(shape))
    return O.counter_uniform(start, n, seed, lo, hi).astype(O.bfloat16).reshape(shape)


def test_device_generator_matches_host_twin():
    for dt, npdt in ((kf.float, np.float32), (kf.bfloat16, O.bfloat16), (kf.half, np.float16)):
        t = dev_uniform((3, 1000), dt, 42, -7.0, 9.0)
        ref = O.counter_uniform(0, 3000, 42, -7.0, 9.0).astype(npdt).reshape(3, 1000)
        assert np.array_equal(t.numpy().view(np.uint8), ref.view(np.uint8))


def test_c2_bf16_gemm_8192_full_size():
    n = 8192
    A, B = dev_uniform((n, n), kf.bfloat16, 1), dev_uniform((n, n), kf.bfloat16, 2)
    C = kf.gemm(A, B, 1.0, 0.0)
    assert C.sizes() == [n, n] and C.dtype() == kf.bfloat16
    a = host_uniform_bf16((n, n), 1).astype(np.float32).astype(np.float64)
    b = host_uniform_bf16((n, n), 2).astype(np.float32).astype(np.float64)
    rng = np.random.default_rng(0)
    rows = np.unique(np.concatenate([[0, 127, 128, 255, 256, n - 1], rng.integers(0, n, 58)]))
    cols = np.unique(np.concatenate([[0, 127, 128, 255, 256, n - 1], rng.integers(0, n, 58)]))
    got = C.float().numpy().astype(np.float64)
    ex_r, mass_r = a[rows] @ b, np.abs(a[rows]) @ np.abs(b)
    ex_c, mass_c = a @ b[:, cols], np.abs(a) @ np.abs(b[:, cols])
    assert np.all(np.abs(got[rows] - ex_r) <= 2e-2 * np.abs(ex_r) + 2e-3 * mass_r)
    assert np.all(np.abs(got[:, cols] - ex_c) <= 2e-2 * np.abs(ex_c) + 2e-3 * mass_c)


def test_c3_attention_fwd_bwd_full_size():
    B, H, S, D = 8, 32, 4096, 128
    q, k, v, do = (dev_uniform((B, H, S, D), kf.bfloat16, 10 + i) for i in range(4))
    out, lse = kf.causal_attention_fwd(q, k, v)
    dq, dk, dv = kf.causal_attention_bwd(do, q, k, v, out, lse)
    kf.synchronize()
    head = S * D
    for (b, h) in ((0, 0), (3, 17), (7, 31)):
        start = (b * H + h) * head
        hq, hk, hv, hdo = (host_uniform_bf16((1, 1, S, D), 10 + i, start=start) for i in range(4))
        eo, elogit = O.causal_attention(hq, hk, hv, return_lse=True)
        edq, edk, edv = O.causal_attention_bwd(hq, hk, hv, hdo)
        pick = lambda t: t[b][h].contiguous().float().numpy().astype(np.float64)
        go = pick(out)
        # 2e-2 relative with an absolute floor of the output scale (rows deep in the sequence average ~ S values of |v| <= 1)
        assert np.all(np.abs(go - eo[0, 0]) <= 2e-2 * np.abs(eo[0, 0]) + 4e-3), ("fwd", b, h, float(np.abs(go - eo[0, 0]).max()))
        glse = lse[b][h].contiguous().numpy().astype(np.float64)
        assert np.all(np.abs(glse - elogit[0, 0]) <= 1e-3 * np.maximum(1.0, np.abs(elogit[0, 0])))
        for name, got, ex in (("dq", pick(dq), edq[0, 0]), ("dk", pick(dk), edk[0, 0]), ("dv", pick(dv), edv[0, 0])):
            scale = np.abs(ex).max()
            err = np.abs(got - ex)
            assert np.all(err <= 2e-2 * np.abs(ex) + 1e-2 * scale), (name, b, h, float(err.max()), float(scale))
            # and in aggregate the error must be far below the band (guards against a systematically wrong tile)
            assert np.linalg.norm(got - ex) <= 1e-2 * np.linalg.norm(ex), (name, b, h)


def test_c3_attention_bwd_is_bit_reproducible_full_size():
    B, H, S, D = 2, 4, 4096, 128
    q, k, v, do = (dev_uniform((B, H, S, D), kf.bfloat16, 20 + i) for i in range(4))
    out, lse = kf.causal_attention_fwd(q, k, v)
    a = [t.numpy().view(np.uint16) for t in kf.causal_attention_bwd(do, q, k, v, out, lse)]
    b = [t.numpy().view(np.uint16) for t in kf.causal_attention_bwd(do, q, k, v, out, lse)]
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_c4_topk_65536_rows_full_size():
    rows, cols, k, seed = 65536, 32768, 64, 1234
    x = dev_uniform((rows, cols), kf.float, seed, -1e5, 1e5)  # 8.6 GB, > 2^31 elements
    vals, idx = x.topk(k, 1, True)
    assert vals.sizes() == [rows, k] and idx.sizes() == [rows, k] and idx.dtype() == kf.long
    gv, gi = vals.numpy(), idx.numpy()
    del x
    rng = np.random.default_rng(7)
    sample = np.unique(np.concatenate([[0, 1, 32767, 32768, 65534, 65535], rng.integers(0, rows, 1020)]))
    assert sample.size >= 1000
    for r in sample:
        row = O.counter_uniform(int(r) * cols, cols, seed, -1e5, 1e5).reshape(1, cols)
        ev, ei = O.topk(row, k, 1, True)
        assert np.array_equal(gv[r].view(np.uint32), ev[0].view(np.uint32)), int(r)
        assert np.array_equal(gi[r], ei[0]), int(r)
    # whole-output properties: sorted descending, indices in range and pointing at the values they claim
    assert np.all(gv[:, :-1] >= gv[:, 1:])
    assert gi.min() >= 0 and gi.max() < cols


def test_c5_block_full_size_sharded_equals_full_batch():
    """The data-parallel identity at the real C5 shape on one GPU: the mean of the gradients of two 1-sample shards equals the
    gradient of the 2-sample step (what the NCCL AVG all-reduce computes across ranks; tools/gpu_dist_check.py does the same
    across 8 real ranks).  Also checks the fused-epilogue block against the reference-API block at full size."""
    from kfunca_b200.block import Block

    S = E = 4096
    H = 32
    x = dev_uniform((2, S, E), kf.bfloat16, 77)
    blk = Block(E, H, dtype=kf.bfloat16, device=0, seed=7)
    loss_full = float(blk.step(x).float().numpy().reshape(-1)[0])
    full = {n: p.grad().float().numpy().astype(np.float64) for n, p in blk.params.items()}
    acc, losses = None, []
    for s in range(2):
        xs = x[s:s + 1].contiguous()
        losses.append(float(blk.step(xs).float().numpy().reshape(-1)[0]))
        g = {n: p.grad().float().numpy().astype(np.float64) for n, p in blk.params.items()}
        acc = g if acc is None else {n: acc[n] + g[n] for n in g}
    assert np.isfinite(loss_full)
    assert abs(loss_full - 0.5 * (losses[0] + losses[1])) <= 2e-2 * max(1e-3, abs(loss_full))
    for n in full:
        avg = 0.5 * acc[n]
        scale = np.abs(full[n]).max() + 1e-30
        assert np.linalg.norm(avg - full[n]) <= 2e-2 * np.linalg.norm(full[n]) + 1e-12, n
        assert np.abs(avg - full[n]).max() <= 5e-2 * scale, n
    # fused (gemm_residual / gemm_glu / qkv_attention) vs the composition from the reference's own operator names
    blk_ref = Block(E, H, dtype=kf.bfloat16, device=0, seed=7, fused=False)
    xs = x[0:1].contiguous()
    l1 = float(blk.step(xs).float().numpy().reshape(-1)[0])
    g1 = {n: p.grad().float().numpy().astype(np.float64) for n, p in blk.params.items()}
    l2 = float(blk_ref.step(xs).float().numpy().reshape(-1)[0])
    g2 = {n: p.grad().float().numpy().astype(np.float64) for n, p in blk_ref.params.items()}
    assert abs(l1 - l2) <= 2e-2 * max(1e-3, abs(l2))
    for n in g1:
        assert np.linalg.norm(g1[n] - g2[n]) <= 2e-2 * np.linalg.norm(g2[n]) + 1e-12, n
