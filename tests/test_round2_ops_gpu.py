"""GPU parity for the round-2 kernels: fp32 GEMM on tensor cores (bf16 x 3 split), GEMM epilogue fusions (residual, GLU), rms_norm,
one-launch column statistics (norm_stat / mean_var), embedding gather + deterministic scatter-add, autograd fixes."""
import os

import numpy as np
import pytest

import kfunca_b200 as kf
from oracle import oracle as O

pytestmark = pytest.mark.gpu
RNG = np.random.default_rng(2024)


def g(a):
    return kf.from_numpy(np.ascontiguousarray(a), 0)


class env:
    def __init__(self, **kw):
        self.kw = kw

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kw}
        os.environ.update(self.kw)

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


# ------------------------------------------------------------------------------------------------ fp32 GEMM on tcgen05
@pytest.mark.parametrize("mode", ["x6", "x9"])
@pytest.mark.parametrize("shape", [(256, 256, 256), (123, 457, 234), (1000, 72, 520), (300, 1031, 264), (512, 4096, 384)])
def test_gemm_f32_tensor_core_1e5_of_l1_mass(mode, shape):
    m, k, n = shape
    a, b = RNG.uniform(-10, 10, (m, k)).astype(np.float32), RNG.uniform(-10, 10, (k, n)).astype(np.float32)
    with env(KF_GEMM_F32=mode):
        l0 = kf.launch_count()
        out = kf.gemm(g(a), g(b), 1.0, 0.0)
        assert kf.launch_count() - l0 == 3  # split A, split B, tcgen05 kernel: the tensor-core path really ran
    exact = O.gemm(a, b)
    mass = np.abs(a.astype(np.float64)) @ np.abs(b.astype(np.float64))
    err = np.abs(out.numpy().astype(np.float64) - exact)
    assert np.all(err <= 1e-5 * mass), float((err / mass).max())
    with env(KF_GEMM_F32="simt"):
        ref = kf.gemm(g(a), g(b), 1.0, 0.0).numpy()
    # and it is at least as accurate as the FFMA kernel on the same data (both far below the band)
    assert (err / mass).max() <= max(4.0 * (np.abs(ref.astype(np.float64) - exact) / mass).max(), 2e-7)


def test_gemm_f32_tensor_core_positive_long_k():
    """all-positive operands and K = 8192: the worst case for a truncating accumulator (every partial sum has the same sign)"""
    m, k, n = 256, 8192, 256
    a, b = RNG.uniform(0, 1, (m, k)).astype(np.float32), RNG.uniform(0, 1, (k, n)).astype(np.float32)
    out = kf.gemm(g(a), g(b), 1.0, 0.0).numpy().astype(np.float64)
    exact = O.gemm(a, b)
    rel = np.abs(out - exact) / exact  # == L1 mass here
    assert rel.max() <= 1e-5, float(rel.max())


@pytest.mark.parametrize("ta,tb", [(False, False), (True, False), (False, True), (True, True)])
def test_matmul_f32_tensor_core_layouts_and_batch(ta, tb):
    bsz, m, k, n = 3, 200, 328, 136
    a = RNG.uniform(-1, 1, (bsz, k, m) if ta else (bsz, m, k)).astype(np.float32)
    b = RNG.uniform(-1, 1, (bsz, n, k) if tb else (bsz, k, n)).astype(np.float32)
    with env(KF_GEMM_F32="x6"):
        out = kf.matmul(g(a), ta, g(b), tb, 0.5).numpy()
    a64 = np.swapaxes(a, 1, 2).astype(np.float64) if ta else a.astype(np.float64)
    b64 = np.swapaxes(b, 1, 2).astype(np.float64) if tb else b.astype(np.float64)
    exact, mass = 0.5 * (a64 @ b64), 0.5 * (np.abs(a64) @ np.abs(b64))
    assert np.all(np.abs(out - exact) <= 1e-5 * mass)
    # broadcast B over the batch
    with env(KF_GEMM_F32="x6"):
        out2 = kf.matmul(g(a), ta, g(b[0]), tb, 1.0).numpy()
    assert np.all(np.abs(out2 - a64 @ b64[0]) <= 1e-5 * (np.abs(a64) @ np.abs(b64[0])))


def test_gemm_f32_beta_and_extreme_magnitudes():
    m, k, n = 256, 512, 256
    a = (RNG.uniform(-1, 1, (m, k)) * 10.0 ** RNG.integers(-12, 12, (m, 1))).astype(np.float32)
    b = (RNG.uniform(-1, 1, (k, n)) * 10.0 ** RNG.integers(-12, 12, (1, n))).astype(np.float32)
    c0 = RNG.uniform(-1, 1, (m, n)).astype(np.float32)
    out = g(c0)
    with env(KF_GEMM_F32="x6"):
        kf.gemm_out(out, g(a), g(b), 2.0, 0.5)
    exact = 2.0 * O.gemm(a, b) + 0.5 * c0.astype(np.float64)
    mass = 2.0 * (np.abs(a.astype(np.float64)) @ np.abs(b.astype(np.float64))) + 0.5 * np.abs(c0)
    assert np.all(np.abs(out.numpy() - exact) <= 1e-5 * mass)


# ------------------------------------------------------------------------------------------------ fused epilogues
@pytest.mark.parametrize("dt", ["bfloat16", "half", "float"])
@pytest.mark.parametrize("shape", [(512, 256, 384), (130, 264, 1032), (2, 300, 128, 256)])
def test_gemm_residual_is_bit_identical_to_gemm_then_add(dt, shape):
    *lead, k, n = shape
    kdt = getattr(kf, dt)
    a = g(RNG.uniform(-1, 1, (*lead, k)).astype(np.float32)).to(kdt)
    b = g(RNG.uniform(-1, 1, (k, n)).astype(np.float32)).to(kdt)
    r = g(RNG.uniform(-1, 1, (*lead, n)).astype(np.float32)).to(kdt)
    fused = kf.gemm_residual(a, b, r, 1.0)
    composed = r + kf.gemm(a, b, 1.0, 0.0)
    assert fused.sizes() == composed.sizes()
    assert np.array_equal(fused.numpy().view(np.uint8), composed.numpy().view(np.uint8))


@pytest.mark.parametrize("dt", ["bfloat16", "half", "float", "double"])
@pytest.mark.parametrize("shape", [(512, 256, 384), (130, 264, 1032), (2, 300, 128, 768), (3, 7, 16, 24)])
def test_qkv_linear_bias_in_the_epilogue(dt, shape):
    """qkv_linear(x, w, bias) (ref: README.md:32, the reference's next planned operator): the bias row added in the GEMM epilogue is
    bit-identical to gemm followed by a broadcast add, matches the float64 oracle, and without a bias it IS gemm."""
    *lead, k, n = shape
    kdt = getattr(kf, dt)
    x = g(RNG.uniform(-1, 1, (*lead, k)).astype(np.float32)).to(kdt)
    w = g(RNG.uniform(-1, 1, (k, n)).astype(np.float32)).to(kdt)
    brow = RNG.uniform(-1, 1, (n,)).astype(np.float32)
    bias = g(brow).to(kdt)
    fused = kf.qkv_linear(x, w, bias)
    plain = kf.gemm(x, w, 1.0, 0.0)
    composed = plain + bias.view(*([1] * len(lead)), n)
    assert fused.sizes() == [*lead, n]
    assert np.array_equal(fused.numpy().view(np.uint8), composed.numpy().view(np.uint8))
    assert np.array_equal(kf.qkv_linear(x, w, None).numpy().view(np.uint8), plain.numpy().view(np.uint8))
    xs, ws, bs = (t.double().numpy() if dt == "double" else t.float().numpy().astype(np.float64) for t in (x, w, bias))
    exact = xs @ ws + bs
    tol = {"bfloat16": 2e-2, "half": 2e-2, "float": 1e-5, "double": 1e-12}[dt]
    mass = np.abs(xs) @ np.abs(ws) + np.abs(bs)
    got = fused.double().numpy() if dt == "double" else fused.float().numpy().astype(np.float64)
    assert np.all(np.abs(got - exact) <= tol * mass + (2e-2 * np.abs(exact) if tol > 1e-3 else 0))


def test_qkv_linear_gradients():
    m, k, n = 3 * 128, 256, 384
    mk = lambda shape: g(RNG.uniform(-1, 1, shape).astype(np.float32))
    x, w, b, go = mk((2, m // 2, k)), mk((k, n)), mk((n,)), mk((2, m // 2, n))
    for t in (x, w, b):
        t.set_requires_grad(True)
    kf.qkv_linear(x, w, b).backward(go)
    xn, wn, gn = x.numpy().reshape(m, k).astype(np.float64), w.numpy().astype(np.float64), go.numpy().reshape(m, n).astype(np.float64)
    np.testing.assert_allclose(x.grad().numpy().reshape(m, k), gn @ wn.T, rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(w.grad().numpy(), xn.T @ gn, rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(b.grad().numpy().reshape(n), gn.sum(0), rtol=1e-4, atol=1e-4)
    assert b.grad().sizes() == [n]


@pytest.mark.parametrize("dt", ["bfloat16", "half"])
@pytest.mark.parametrize("shape", [(512, 256, 384), (300, 264, 1032), (2, 300, 128, 256), (64, 64, 64)])
def test_gemm_glu_is_bit_identical_to_two_gemms_and_a_multiply(dt, shape):
    *lead, k, n = shape
    kdt = getattr(kf, dt)
    a = g(RNG.uniform(-1, 1, (*lead, k)).astype(np.float32)).to(kdt)
    b1 = g(RNG.uniform(-1, 1, (k, n)).astype(np.float32)).to(kdt)
    b3 = g(RNG.uniform(-1, 1, (k, n)).astype(np.float32)).to(kdt)
    fused = kf.gemm_glu(a, b1, b3)
    composed = kf.gemm(a, b1, 1.0, 0.0) * kf.gemm(a, b3, 1.0, 0.0)
    assert np.array_equal(fused.numpy().view(np.uint16), composed.numpy().view(np.uint16))


def test_fused_epilogue_gradients_match_composed_graph():
    m, k, n = 384, 256, 512
    mk = lambda shape: g(RNG.uniform(-1, 1, shape).astype(np.float32)).bfloat16()
    a, b1, b3, w2, r = mk((m, k)), mk((k, n)), mk((k, n)), mk((n, k)), mk((m, k))
    go = mk((m, k))
    grads = []
    for fused in (True, False):
        for t in (a, b1, b3, w2, r):
            t.set_requires_grad(True)
            t.zero_grad()
        if fused:
            y = kf.gemm_residual(kf.gemm_glu(a, b1, b3), w2, r, 1.0)
        else:
            y = r + kf.gemm(kf.gemm(a, b1, 1.0, 0.0) * kf.gemm(a, b3, 1.0, 0.0), w2, 1.0, 0.0)
        y.backward(go)
        grads.append([t.grad().float().numpy().astype(np.float64) for t in (a, b1, b3, w2, r)])
    for gf, gc in zip(*grads):
        assert np.linalg.norm(gf - gc) <= 1e-2 * np.linalg.norm(gc)


# ------------------------------------------------------------------------------------------------ norms
@pytest.mark.parametrize("dt,tol", [("float", 2e-5), ("bfloat16", 2e-2)])
def test_rms_norm_forward_backward(dt, tol):
    rows, E = 96, 1024
    x32 = RNG.uniform(-3, 3, (rows, E)).astype(np.float32)
    g32 = RNG.uniform(0.5, 1.5, (1, E)).astype(np.float32)
    dy32 = RNG.uniform(-1, 1, (rows, E)).astype(np.float32)
    kdt = getattr(kf, dt)
    x, gain, dy = g(x32).to(kdt), g(g32).to(kdt), g(dy32).to(kdt)
    x.set_requires_grad(True)
    gain.set_requires_grad(True)
    y = kf.rms_norm(x, gain, 1e-5)
    y.backward(dy)
    xr, gr, dyr = (t.float().numpy() for t in (x, gain, dy))
    ey = O.rms_norm(xr, gr, 1e-5)
    edx, edg = O.rms_norm_bwd(xr, gr, dyr, 1e-5)
    assert np.all(np.abs(y.float().numpy() - ey) <= tol * np.maximum(1.0, np.abs(ey)))
    assert np.all(np.abs(x.grad().float().numpy() - edx) <= tol * np.maximum(1.0, np.abs(edx)))
    gg = gain.grad().float().numpy().reshape(-1)
    assert np.all(np.abs(gg - edg) <= 5 * tol * np.maximum(1.0, np.abs(edg)))


@pytest.mark.parametrize("shape", [(64, 64), (1024, 2048), (4096, 4096), (4096 * 4 + 3, 4096 * 4 + 3), (7, 3), (1, 130)])
def test_norm_stat_reference_shapes_one_launch(shape):  # ref: test/test_tensor.py:134-146 (incl. its 4096^2 and 16387^2 shapes)
    x = RNG.uniform(-10, 10, shape).astype(np.float32)
    t = g(x)
    l0 = kf.launch_count()
    mean, invstd = t.norm_stat(0)
    assert kf.launch_count() - l0 == 1
    em, ei = O.norm_stat(x)
    assert mean.sizes() == [1, shape[1]] and invstd.sizes() == [1, shape[1]]
    assert np.all(np.abs(mean.numpy() - em) <= 1e-5 * np.abs(x.astype(np.float64)).mean(0, keepdims=True) + 1e-6)
    if shape[0] > 1:
        assert np.allclose(invstd.numpy(), ei, rtol=2e-5, atol=0)


@pytest.mark.parametrize("shape,dim", [((13, 325, 127), 1), ((40, 33, 8), 0), ((5, 1000, 4), 1)])
def test_mean_var_non_last_dim_one_launch(shape, dim):  # ref: test/test_tensor.py:120-132
    x = RNG.uniform(-10, 10, shape).astype(np.float32)
    t = g(x)
    for take_sqrt in (False, True):
        l0 = kf.launch_count()
        m, v = t.mean_var(dim, take_sqrt)
        assert kf.launch_count() - l0 == 1
        em, ev = O.mean_var(x, dim, take_sqrt)
        assert np.allclose(m.numpy(), em, rtol=1e-5, atol=1e-5)
        assert np.allclose(v.numpy(), ev, rtol=2e-5, atol=1e-6)


# ------------------------------------------------------------------------------------------------ indexing
@pytest.mark.parametrize("dt", [np.float32, np.float64, np.int32, np.int64, np.float16])
def test_embedding_gather_bit_exact(dt):
    V, E = 1000, 130 if dt != np.float16 else 128
    w = RNG.uniform(-100, 100, (V, E)).astype(dt)
    idx = RNG.integers(-V, V, (7, 33)).astype(np.int64)
    out = kf.embedding(g(w), g(idx))
    assert out.sizes() == [7, 33, E]
    assert np.array_equal(out.numpy().view(np.uint8), O.embedding(w, idx).view(np.uint8))


def test_embedding_backward_deterministic_scatter_add():
    V, E, n = 50, 256, 4000  # many repeats of every id
    w = g(RNG.uniform(-1, 1, (V, E)).astype(np.float32))
    w.set_requires_grad(True)
    idx = RNG.integers(0, V, (n,)).astype(np.int64)
    go = RNG.uniform(-1, 1, (n, E)).astype(np.float32)
    outs = []
    for _ in range(2):
        w.zero_grad()
        kf.embedding(w, g(idx)).backward(g(go))
        outs.append(w.grad().numpy())
    assert np.array_equal(outs[0].view(np.uint32), outs[1].view(np.uint32))
    assert np.array_equal(outs[0].view(np.uint32), O.embedding_bwd(idx, go, V).view(np.uint32))  # same order, same precision
    # ids that never occur get exact zeros
    w2 = g(np.zeros((V + 5, E), np.float32))
    w2.set_requires_grad(True)
    kf.embedding(w2, g(idx)).backward(g(go))
    assert not w2.grad().numpy()[V:].any()


def test_index_put_against_oracle():  # ref: test/test_tensor.py:273-284
    x = RNG.uniform(-1e4, 1e4, (13, 15)).astype(np.float32)
    i0, i1 = np.array([0, 5, 1, 2, -1]).astype(np.int64), np.array([0, 11, 1, 0, -2]).astype(np.int64)
    vals = RNG.uniform(-1e4, 1e4, 5).astype(np.float32)
    t = g(x)
    t.index_put_([g(i0), g(i1)], g(vals))
    assert np.array_equal(t.numpy(), O.index_put(x, [i0, i1], vals))


# ------------------------------------------------------------------------------------------------ autograd fixes
def test_getitem_int_index_routes_gradient_to_the_source():
    x = g(RNG.uniform(-1, 1, (4, 6)).astype(np.float32))
    x.set_requires_grad(True)
    go = RNG.uniform(-1, 1, (6,)).astype(np.float32)
    (x[2] * 3.0).backward(g(go))
    want = np.zeros((4, 6), np.float32)
    want[2] = 3.0 * go
    assert np.array_equal(x.grad().numpy(), want)


def test_matmul_broadcast_batch_gradient_is_summed():
    a = g(RNG.uniform(-1, 1, (3, 20, 16)).astype(np.float64))  # [B, K, M], used transposed
    b = g(RNG.uniform(-1, 1, (20, 8)).astype(np.float64))      # shared 2-D operand
    b.set_requires_grad(True)
    go = RNG.uniform(-1, 1, (3, 16, 8))
    kf.matmul(a, True, b, False, 1.0).backward(g(go))
    want = sum(a.numpy()[i] @ go[i] for i in range(3))
    assert b.grad().sizes() == [20, 8]
    assert np.allclose(b.grad().numpy(), want, rtol=1e-12, atol=1e-12)
