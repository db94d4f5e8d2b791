"""GPU parity for causal attention forward/backward against the float64 oracle and the reference build's outputs.
fp32: 1e-5-class (abs/rel on O(1) outputs); bf16/fp16: 2e-2."""
import os

import numpy as np
import pytest

import kfunca_b200 as kf
from oracle import oracle as O
from oracle.golden_cases import cases

pytestmark = pytest.mark.gpu
RNG = np.random.default_rng(2026)
GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_outputs.npz"))


def g(a):
    return kf.from_numpy(a, 0)


def to16(x, dt):
    return x.astype(np.float32).astype(O.bfloat16 if dt == "bf16" else np.float16)


def qkv(b, h, sq, skv, d, lo=-1.0, hi=1.0, dtype=np.float32):
    mk = lambda s: RNG.uniform(lo, hi, s).astype(dtype)
    return mk((b, h, sq, d)), mk((b, h, skv, d)), mk((b, h, skv, d))


@pytest.mark.parametrize("shape", [(2, 4, 32, 256, 128), (3, 5, 64, 32, 64), (5, 16, 65, 33, 123), (1, 2, 100, 100, 40), (1, 1, 1, 1, 8)])
def test_causal_attention_fp32(shape):  # ref: test/test_nn.py:11-33 (same tuples, U(-10,10) inputs)
    q, k, v = qkv(*shape, lo=-10, hi=10)
    out = kf.causal_attention(g(q), g(k), g(v))
    assert out.sizes() == list(q.shape) and out.dtype() == kf.float
    # U(-10,10) gives logits with sigma ~33 (near one-hot softmax): the reference's own tolerance (test/common.py:6-11)
    np.testing.assert_allclose(out.numpy(), O.causal_attention(q, k, v), rtol=1e-3, atol=1e-3)


def test_causal_attention_fp32_tight_and_fp64():
    q, k, v = qkv(2, 3, 96, 96, 64)
    out = kf.causal_attention(g(q), g(k), g(v)).numpy()
    np.testing.assert_allclose(out, O.causal_attention(q, k, v), rtol=1e-5, atol=1e-5)
    q, k, v = qkv(1, 2, 50, 70, 32, dtype=np.float64)
    out = kf.causal_attention(g(q), g(k), g(v)).numpy()
    np.testing.assert_allclose(out, O.causal_attention(q, k, v), rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("shape,lo,hi,tol", [((1, 2, 256, 256, 128), -1, 1, 1e-5), ((1, 1, 200, 333, 128), -1, 1, 1e-5), ((2, 2, 640, 640, 64), -1, 1, 1e-5),
                                             ((3, 5, 130, 130, 128), -1, 1, 1e-5), ((1, 3, 513, 257, 64), -1, 1, 1e-5), ((1, 2, 100, 700, 128), -1, 1, 1e-5),
                                             ((1, 1, 2048, 2048, 128), 0, 1, 1e-5), ((2, 4, 32, 256, 128), -10, 10, 1e-3),
                                             ((1, 64, 1, 1024, 128), -1, 1, 1e-5), ((1, 2, 100, 5000, 64), -1, 1, 1e-5), ((1, 4, 129, 129, 128), -1, 1, 1e-5)])
def test_causal_attention_fp32_tensor_core_path(shape, lo, hi, tol, monkeypatch):
    """fp32 forward on tcgen05 (three bf16 planes per operand, attention_f32_tc.cu) against the float64 oracle at the fp32 band,
    incl. an all-positive 2048-key case (every truncating addition of the tensor core has the same sign) and the reference's
    U(-10, 10) inputs at its own tolerance; the FFMA kernel (KF_ATTN_F32=simt) must agree with it within the same band."""
    q, k, v = qkv(*shape, lo=lo, hi=hi)
    gq, gk, gv = g(q), g(k), g(v)
    launches = kf.launch_count()
    out, lse = kf.causal_attention_fwd(gq, gk, gv)
    assert kf.launch_count() - launches == 4  # three plane splits + the tensor-core kernel: the path under test really ran
    exact = O.causal_attention(q, k, v)
    np.testing.assert_allclose(out.numpy(), exact, rtol=tol, atol=tol)
    monkeypatch.setenv("KF_ATTN_F32", "simt")
    out_s, lse_s = kf.causal_attention_fwd(gq, gk, gv)
    np.testing.assert_allclose(out_s.numpy(), exact, rtol=tol, atol=tol)
    np.testing.assert_allclose(lse.numpy(), lse_s.numpy(), rtol=1e-5, atol=1e-4 if hi > 1 else 1e-5)


@pytest.mark.parametrize("name", ["attn_0", "attn_1", "attn_2"])
def test_attention_against_reference_outputs(name):
    _, inp, _ = next((k, i, p) for n, k, i, p in cases() if n == name)
    out = kf.causal_attention(g(inp["q"]), g(inp["k"]), g(inp["v"])).numpy()
    np.testing.assert_allclose(out, GOLD[f"{name}.out"], rtol=2e-5, atol=2e-5)


@pytest.mark.parametrize("kernel", ["pers", "cta"])
@pytest.mark.parametrize("hg", ["1", "4", "7", "1000000"])
@pytest.mark.parametrize("shape", [(3, 5, 130, 130, 128), (2, 9, 513, 257, 64), (1, 30, 300, 700, 128), (1, 149, 200, 200, 64)])
def test_forward_kernels_and_item_orders_agree(kernel, hg, shape, monkeypatch):
    """Both forward kernels (persistent work queue / one CTA per query pair) under every scheduling order (heads per group: one,
    partial last group, all) give the bits of the default path: the order only decides which CTA computes an item.  The 149-head
    case has one work item more than the persistent grid has CTAs (148): exactly one CTA claims a second item from the queue."""
    q, k, v = (to16(t, "bf16") for t in qkv(*shape))
    gq, gk, gv = g(q), g(k), g(v)
    ref_o, ref_l = kf.causal_attention_fwd(gq, gk, gv)
    ref_o, ref_l = ref_o.float().numpy(), ref_l.numpy()
    monkeypatch.setenv("KF_ATTN_FWD", kernel)
    monkeypatch.setenv("KF_ATTN_HG", hg)
    o, l = kf.causal_attention_fwd(gq, gk, gv)
    assert np.array_equal(o.float().numpy(), ref_o) and np.array_equal(l.numpy(), ref_l)
    np.testing.assert_allclose(ref_o, O.causal_attention(q, k, v), rtol=2e-2, atol=1e-2)


@pytest.mark.parametrize("dt", ["bf16", "fp16"])
@pytest.mark.parametrize("shape", [(3, 5, 130, 130, 128), (2, 9, 513, 257, 64), (1, 30, 300, 700, 128), (1, 2, 1024, 1024, 128)])
def test_forward_k64_variant(dt, shape, monkeypatch):
    """KF_ATTN_FWD=k64 (64-key blocks, the next S queued before P V, double-buffered P in tensor memory): an opt-in experiment that
    measured slower than the default, kept correct — same band against the oracle, LSE within 1e-3 of the default kernel's."""
    q, k, v = (to16(t, dt) for t in qkv(*shape))
    gq, gk, gv = g(q), g(k), g(v)
    _, ref_l = kf.causal_attention_fwd(gq, gk, gv)
    monkeypatch.setenv("KF_ATTN_FWD", "k64")
    o, l = kf.causal_attention_fwd(gq, gk, gv)
    np.testing.assert_allclose(o.float().numpy(), O.causal_attention(q, k, v), rtol=2e-2, atol=1e-2)
    np.testing.assert_allclose(l.numpy(), ref_l.numpy(), rtol=0, atol=1e-3)


@pytest.mark.parametrize("dt", ["bf16", "fp16"])
@pytest.mark.parametrize("shape", [(1, 2, 128, 128, 128), (2, 3, 256, 256, 64), (1, 2, 384, 384, 128), (2, 2, 200, 333, 128),
                                   (1, 1, 130, 70, 64), (1, 4, 1024, 1024, 128)])
def test_causal_attention_tc_16bit(dt, shape):
    q, k, v = (to16(t, dt) for t in qkv(*shape))
    out, lse = kf.causal_attention_fwd(g(q), g(k), g(v))
    assert out.dtype() == (kf.bfloat16 if dt == "bf16" else kf.half)
    exact, lse_exact = O.causal_attention(q, k, v, return_lse=True)
    got = out.float().numpy().astype(np.float64)
    assert np.all(np.abs(got - exact) <= 2e-2 * np.abs(exact) + 8e-3), float(np.abs(got - exact).max())
    np.testing.assert_allclose(lse.numpy(), lse_exact, rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize("dt", ["bf16", "fp16"])
@pytest.mark.parametrize("shape", [(1, 2, 512, 512, 128), (1, 2, 700, 900, 64), (2, 1, 257, 257, 128)])
def test_causal_attention_tc_large_logits(dt, shape):
    """wide logit range (sigma ~ 10 in the exp2 domain): the lazily-moved reference max must rescale O and l"""
    q, k, v = (to16(t, dt) for t in qkv(*shape, lo=-4.0, hi=4.0))
    out, lse = kf.causal_attention_fwd(g(q), g(k), g(v))
    exact, lse_exact = O.causal_attention(q, k, v, return_lse=True)
    got = out.float().numpy().astype(np.float64)
    assert np.all(np.abs(got - exact) <= 2e-2 * np.abs(exact) + 3e-2), float(np.abs(got - exact).max())
    np.testing.assert_allclose(lse.numpy(), lse_exact, rtol=1e-2, atol=2e-2)


def test_attention_16bit_odd_head_dim_uses_simt():
    q, k, v = (to16(t, "bf16") for t in qkv(1, 2, 40, 40, 24))
    got = kf.causal_attention(g(q), g(k), g(v)).float().numpy()
    np.testing.assert_allclose(got, O.causal_attention(q, k, v), rtol=2e-2, atol=1e-2)


def test_attention_errors():
    q, k, v = qkv(1, 2, 8, 8, 16)
    with pytest.raises(RuntimeError):
        kf.causal_attention(g(q), g(k[:, :1]), g(v))
    with pytest.raises(RuntimeError):
        kf.causal_attention(g(q), g(k), g(v).half())


@pytest.mark.parametrize("shape,dtype", [((2, 3, 48, 48, 32), np.float32), ((1, 2, 33, 50, 24), np.float32), ((1, 2, 40, 40, 16), np.float64)])
def test_attention_backward_fp32(shape, dtype):
    q, k, v = qkv(*shape, dtype=dtype)
    do = RNG.uniform(-1, 1, q.shape).astype(dtype)
    gq, gk, gv = g(q), g(k), g(v)
    for t in (gq, gk, gv):
        t.set_requires_grad(True)
    kf.causal_attention(gq, gk, gv).backward(g(do))
    dq, dk, dv = O.causal_attention_bwd(q, k, v, do)
    tol = 1e-4 if dtype == np.float32 else 1e-10
    np.testing.assert_allclose(gq.grad().numpy(), dq, rtol=tol, atol=tol)
    np.testing.assert_allclose(gk.grad().numpy(), dk, rtol=tol, atol=tol)
    np.testing.assert_allclose(gv.grad().numpy(), dv, rtol=tol, atol=tol)


@pytest.mark.parametrize("dt", ["bf16", "fp16"])
@pytest.mark.parametrize("shape", [(1, 2, 128, 128, 128), (2, 2, 256, 256, 64), (1, 2, 200, 200, 128), (1, 2, 300, 200, 64), (1, 1, 100, 333, 128),
                                   (1, 2, 1024, 1024, 128), (2, 1, 640, 640, 64), (1, 1, 1, 1, 64), (1, 1, 65, 129, 128)])
def test_attention_backward_16bit(dt, shape):
    """tcgen05 backward (attention_bwd_tc.cu): dQ, dK, dV against the float64 oracle of the same 16-bit inputs"""
    q, k, v = (to16(t, dt) for t in qkv(*shape))
    do = to16(RNG.uniform(-1, 1, q.shape), dt)
    out, lse = kf.causal_attention_fwd(g(q), g(k), g(v))
    dq, dk, dv = kf.causal_attention_bwd(g(do), g(q), g(k), g(v), out, lse)
    eq, ek, ev = O.causal_attention_bwd(q, k, v, do)
    for got, exp in ((dq, eq), (dk, ek), (dv, ev)):
        gotf = got.float().numpy().astype(np.float64)
        scale = np.abs(exp).max()
        # 1e-6 floor: a single-key row has exactly zero dQ/dK in exact arithmetic, the kernel leaves fp32 rounding noise (~4e-8)
        assert np.all(np.abs(gotf - exp) <= 2e-2 * np.abs(exp) + 1e-2 * scale + 1e-6), float(np.abs(gotf - exp).max() / max(scale, 1e-30))


def test_attention_backward_16bit_is_deterministic():
    q, k, v = (to16(t, "bf16") for t in qkv(1, 2, 512, 512, 128))
    do = to16(RNG.uniform(-1, 1, q.shape), "bf16")
    out, lse = kf.causal_attention_fwd(g(q), g(k), g(v))
    a = [t.float().numpy() for t in kf.causal_attention_bwd(g(do), g(q), g(k), g(v), out, lse)]
    b = [t.float().numpy() for t in kf.causal_attention_bwd(g(do), g(q), g(k), g(v), out, lse)]
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
