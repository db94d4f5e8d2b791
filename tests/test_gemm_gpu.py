"""GPU parity for gemm / matmul: fp32/fp64 SIMT path and bf16/fp16 tcgen05 path against the float64 oracle.
Tolerances: fp32 1e-5 and fp64 1e-13 of the L1 mass sum|a||b|; 16-bit 2e-2 relative."""
import os

import numpy as np
import pytest

import kfunca_b200 as kf
from oracle import oracle as O
from oracle.golden_cases import cases

pytestmark = pytest.mark.gpu
RNG = np.random.default_rng(99)
GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_outputs.npz"))


def g(a):
    return kf.from_numpy(a, 0)


def to16(x, dt):
    return x.astype(np.float32).astype(O.bfloat16 if dt == "bf16" else np.float16)


def check16(got, exact, mass, what):
    got = got.astype(np.float32).astype(np.float64)
    err = np.abs(got - exact)
    # 2e-2 relative on well-conditioned entries; an L1-mass floor covers cancellation
    tol = 2e-2 * np.abs(exact) + 2e-3 * mass + 1e-6
    assert np.all(err <= tol), (what, float(err.max()), float((err / tol).max()))


def test_gemm_base_fp64():  # ref: test/test_gemm.py:9-17
    a, b = RNG.uniform(-10, 10, (123, 457)), RNG.uniform(-10, 10, (457, 234))
    out = kf.gemm(g(a), g(b), 1.0, 0.0).numpy()
    mass = np.abs(a) @ np.abs(b)
    assert O.l1_tolerance_ok(out, a @ b, mass, 1e-13)


@pytest.mark.parametrize("shape", [(123, 457, 234), (1, 1, 1), (257, 130, 129), (4, 1000, 3), (512, 512, 512)])
def test_gemm_fp32(shape):
    m, k, n = shape
    a, b = RNG.uniform(-10, 10, (m, k)).astype(np.float32), RNG.uniform(-10, 10, (k, n)).astype(np.float32)
    out = kf.gemm(g(a), g(b), 1.0, 0.0)
    assert out.dtype() == kf.float and out.sizes() == [m, n]
    mass = np.abs(a.astype(np.float64)) @ np.abs(b.astype(np.float64))
    assert O.l1_tolerance_ok(out.numpy(), O.gemm(a, b), mass, 1e-5)
    out2 = kf.gemm(g(a), g(b), 0.5, 0.0).numpy()
    assert O.l1_tolerance_ok(out2, O.gemm(a, b, 0.5), mass, 1e-5)


def test_gemm_leading_dims_fold_into_m():  # ref: gemm_kernel.cu:10-15
    a, b = RNG.uniform(-1, 1, (3, 41, 130)).astype(np.float32), RNG.uniform(-1, 1, (130, 77)).astype(np.float32)
    out = kf.gemm(g(a), g(b), 1.0, 0.0)
    assert out.sizes() == [3, 41, 77]
    np.testing.assert_allclose(out.numpy(), a @ b, rtol=1e-4, atol=1e-4)
    with pytest.raises(RuntimeError):
        kf.gemm(g(a), g(b.T.copy()), 1.0, 0.0)  # K mismatch
    with pytest.raises(RuntimeError):
        kf.gemm(g(a), g(b).half(), 1.0, 0.0)  # dtype mismatch


@pytest.mark.parametrize("name", ["gemm_f64", "gemm_f32"])
def test_gemm_against_reference_outputs(name):
    kind, inp, _ = next((k, i, p) for n, k, i, p in cases() if n == name)
    out = kf.gemm(g(inp["a"]), g(inp["b"]), 1.0, 0.0).numpy()
    ref = GOLD[f"{name}.out"]
    mass = np.abs(inp["a"].astype(np.float64)) @ np.abs(inp["b"].astype(np.float64))
    assert O.l1_tolerance_ok(out, ref.astype(np.float64), mass, 2e-5 if name.endswith("f32") else 1e-13)


@pytest.mark.parametrize("dt", ["bf16", "fp16"])
@pytest.mark.parametrize("shape", [(128, 64, 256), (256, 128, 512), (123, 456, 232), (1000, 72, 40), (384, 8192, 128), (130, 264, 1032)])
def test_gemm_tc_16bit(dt, shape):
    m, k, n = shape
    a, b = to16(RNG.uniform(-1, 1, (m, k)), dt), to16(RNG.uniform(-1, 1, (k, n)), dt)
    out = kf.gemm(g(a), g(b), 1.0, 0.0)
    assert out.dtype() == (kf.bfloat16 if dt == "bf16" else kf.half) and out.sizes() == [m, n]
    exact = O.gemm(a, b)
    mass = np.abs(a.astype(np.float64)) @ np.abs(b.astype(np.float64))
    check16(out.float().numpy(), exact, mass, (dt, shape))


@pytest.mark.parametrize("ta,tb", [(False, False), (False, True), (True, False), (True, True)])
def test_matmul_tc_all_layouts_and_batches(ta, tb):
    for (bsz, m, k, n) in [(1, 256, 192, 384), (3, 200, 136, 72), (2, 128, 64, 128)]:
        a = to16(RNG.uniform(-1, 1, (bsz, k, m) if ta else (bsz, m, k)), "bf16")
        b = to16(RNG.uniform(-1, 1, (bsz, n, k) if tb else (bsz, k, n)), "bf16")
        out = kf.matmul(g(a), ta, g(b), tb, 1.0)
        af = a.astype(np.float64).transpose(0, 2, 1) if ta else a.astype(np.float64)
        bf = b.astype(np.float64).transpose(0, 2, 1) if tb else b.astype(np.float64)
        exact = af @ bf
        mass = np.abs(af) @ np.abs(bf)
        assert out.sizes() == [bsz, m, n]
        check16(out.float().numpy(), exact, mass, (ta, tb, bsz, m, k, n))


@pytest.fixture
def force_cta_pair():
    os.environ["KF_GEMM_CTA_GROUP"] = "2"
    yield
    os.environ.pop("KF_GEMM_CTA_GROUP", None)


@pytest.mark.parametrize("ta,tb", [(False, False), (False, True), (True, False), (True, True)])
def test_matmul_tc_cta_pair_all_layouts(ta, tb, force_cta_pair):
    """the cta_group::2 kernel (256 x 256 tiles over a CTA pair), forced on small / ragged / batched shapes"""
    for (bsz, m, k, n) in [(1, 256, 64, 256), (1, 512, 192, 768), (3, 200, 136, 328), (2, 129, 72, 136), (1, 1000, 520, 264)]:
        a = to16(RNG.uniform(-1, 1, (bsz, k, m) if ta else (bsz, m, k)), "bf16")
        b = to16(RNG.uniform(-1, 1, (bsz, n, k) if tb else (bsz, k, n)), "bf16")
        out = kf.matmul(g(a), ta, g(b), tb, 1.0)
        af = a.astype(np.float64).transpose(0, 2, 1) if ta else a.astype(np.float64)
        bf = b.astype(np.float64).transpose(0, 2, 1) if tb else b.astype(np.float64)
        check16(out.float().numpy(), af @ bf, np.abs(af) @ np.abs(bf), ("pair", ta, tb, bsz, m, k, n))


def test_gemm_tc_cta_pair_beta_fp16_and_many_tiles(force_cta_pair):
    a, b = to16(RNG.uniform(-1, 1, (384, 128)), "fp16"), to16(RNG.uniform(-1, 1, (128, 512)), "fp16")
    c = to16(RNG.uniform(-1, 1, (384, 512)), "fp16")
    gc = g(c)
    kf.gemm_out(gc, g(a), g(b), 2.0, 0.5)
    exact = 2.0 * O.gemm(a, b) + 0.5 * c.astype(np.float64)
    check16(gc.float().numpy(), exact, 2 * np.abs(a.astype(np.float64)) @ np.abs(b.astype(np.float64)) + 1, "pair beta")
    # more tiles than CTA pairs: every cluster loops over several tiles, both accumulator buffers and all ring phases cycle
    m, k, n = 2304 + 40, 1096, 5120 + 8
    a, b = to16(RNG.uniform(-1, 1, (m, k)), "bf16"), to16(RNG.uniform(-1, 1, (k, n)), "bf16")
    out = kf.gemm(g(a), g(b), 1.0, 0.0).float().numpy()
    rows = [0, 127, 128, 255, 256, 1000, 2303, 2304, m - 1]
    af, bf = a[rows].astype(np.float64), b.astype(np.float64)
    check16(out[rows], af @ bf, np.abs(af) @ np.abs(bf), "pair rows")
    cols = [0, 127, 128, 255, 256, 5119, 5120, n - 1]
    af, bf = a.astype(np.float64), b[:, cols].astype(np.float64)
    check16(out[:, cols], af @ bf, np.abs(af) @ np.abs(bf), "pair cols")


def test_gemm_tc_default_dispatch_uses_pair_on_big_shapes():
    """no env override: >= one wave of 256 x 256 tiles takes the pair kernel; result must match the single-CTA kernel bit for bit
    (same k order, same fp32 accumulation)"""
    m, k, n = 2560, 512, 2560
    a, b = to16(RNG.uniform(-1, 1, (m, k)), "bf16"), to16(RNG.uniform(-1, 1, (k, n)), "bf16")
    ga, gb = g(a), g(b)
    o_default = kf.gemm(ga, gb, 1.0, 0.0).float().numpy()
    os.environ["KF_GEMM_CTA_GROUP"] = "1"
    try:
        o_single = kf.gemm(ga, gb, 1.0, 0.0).float().numpy()
    finally:
        os.environ.pop("KF_GEMM_CTA_GROUP", None)
    assert np.array_equal(o_default, o_single)


def test_gemm_tc_unaligned_falls_back_to_simt():
    a, b = to16(RNG.uniform(-1, 1, (33, 77)), "bf16"), to16(RNG.uniform(-1, 1, (77, 45)), "bf16")  # ld not multiple of 8
    out = kf.gemm(g(a), g(b), 1.0, 0.0)
    check16(out.float().numpy(), O.gemm(a, b), np.abs(a.astype(np.float64)) @ np.abs(b.astype(np.float64)), "unaligned")


def test_gemm_out_beta():
    a, b = to16(RNG.uniform(-1, 1, (256, 128)), "bf16"), to16(RNG.uniform(-1, 1, (128, 256)), "bf16")
    c = to16(RNG.uniform(-1, 1, (256, 256)), "bf16")
    gc = g(c)
    kf.gemm_out(gc, g(a), g(b), 2.0, 0.5)
    exact = 2.0 * O.gemm(a, b) + 0.5 * c.astype(np.float64)
    check16(gc.float().numpy(), exact, 2 * np.abs(a.astype(np.float64)) @ np.abs(b.astype(np.float64)) + 1, "beta")
    a32, b32 = RNG.uniform(-1, 1, (50, 60)).astype(np.float32), RNG.uniform(-1, 1, (60, 70)).astype(np.float32)
    c32 = RNG.uniform(-1, 1, (50, 70)).astype(np.float32)
    gc32 = g(c32)
    kf.gemm_out(gc32, g(a32), g(b32), 1.0, 1.0)
    np.testing.assert_allclose(gc32.numpy(), a32 @ b32 + c32, rtol=1e-5, atol=1e-5)


def test_gemm_large_linearity_property():
    """full-size-style check without a CPU GEMM: (A1 + A2) B == A1 B + A2 B within bf16 rounding, and row spot checks."""
    m = k = n = 2048
    a1, a2 = to16(RNG.uniform(-1, 1, (m, k)), "bf16"), to16(RNG.uniform(-1, 1, (m, k)), "bf16")
    b = to16(RNG.uniform(-1, 1, (k, n)), "bf16")
    ga1, gb = g(a1), g(b)
    o1 = kf.gemm(ga1, gb, 1.0, 0.0).float().numpy()
    rows = [0, 1, 127, 128, 1000, 2047]
    exact = a1[rows].astype(np.float64) @ b.astype(np.float64)
    mass = np.abs(a1[rows].astype(np.float64)) @ np.abs(b.astype(np.float64))
    check16(o1[rows], exact, mass, "rows")
    cols = [0, 255, 256, 2047]
    exact_c = a1.astype(np.float64) @ b[:, cols].astype(np.float64)
    check16(o1[:, cols], exact_c, np.abs(a1.astype(np.float64)) @ np.abs(b[:, cols].astype(np.float64)), "cols")


def test_gemm_backward():
    torch = pytest.importorskip("torch")
    a = RNG.uniform(-1, 1, (2, 24, 40)).astype(np.float32)
    w = RNG.uniform(-1, 1, (40, 16)).astype(np.float32)
    go = RNG.uniform(-1, 1, (2, 24, 16)).astype(np.float32)
    ga, gw = g(a), g(w)
    ga.set_requires_grad(True)
    gw.set_requires_grad(True)
    kf.gemm(ga, gw, 1.0, 0.0).backward(g(go))
    ta, tw = torch.tensor(a, requires_grad=True), torch.tensor(w, requires_grad=True)
    (ta @ tw).backward(torch.tensor(go))
    np.testing.assert_allclose(ga.grad().numpy(), ta.grad.numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(gw.grad().numpy(), tw.grad.numpy(), rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("dtype,M,N,K,slab", [("bfloat16", 1000, 384, 512, 256), ("bfloat16", 2048, 1024, 768, 0), ("float", 300, 200, 100, 128),
                                              ("half", 129, 256, 64, 64)])
def test_gemm_host_streamed_equals_device_gemm(dtype, M, N, K, slab):
    """kf_gemm_host (pinned host operands, slab-pipelined upload / GEMM / download) gives the same bits as from_numpy + gemm +
    numpy: each slab runs the very same kernel on the same rows."""
    from kfunca_b200.runtime import PinnedBuffer, gemm_host
    kdt = getattr(kf, dtype)
    rng = np.random.default_rng(3)
    a32 = rng.uniform(-1, 1, (M, K)).astype(np.float32)
    b32 = rng.uniform(-1, 1, (K, N)).astype(np.float32)
    ta, tb = kf.from_numpy(a32, 0).to(kdt), kf.from_numpy(b32, 0).to(kdt)
    want = kf.gemm(ta, tb, 1.0, 0.0)
    raw = {"bfloat16": np.uint16, "half": np.float16, "float": np.float32}[dtype]
    bits = (lambda t: t.numpy().view(raw)) if dtype == "bfloat16" else (lambda t: t.numpy())
    pa, pb, pc = PinnedBuffer((M, K), raw), PinnedBuffer((K, N), raw), PinnedBuffer((M, N), raw)
    pa.array[:] = bits(ta)
    pb.array[:] = bits(tb)
    pc.array[:] = 0
    for _ in range(2):  # twice: the recycled device slabs and events of the first call must not leak into the second
        gemm_host(pc, pa, pb, kdt, 1.0, slab)
        kf.synchronize()
        assert np.array_equal(pc.array, bits(want))
