"""GPU parity tests: the CUDA path (through the pybind module -> C ABI) against the oracle on seeded inputs.
Mirrors the reference's test/test_tensor.py case by case (cited per test) and adds dtype / edge coverage.
Integer, copy, permute, concat/split and index results are compared bit-exactly; floating point within the
north-star tolerances (fp32 1e-5, 16-bit 2e-2), reductions in the L1-mass form."""
import copy

import numpy as np
import pytest

import kfunca_b200 as kf
from oracle import oracle as O

pytestmark = pytest.mark.gpu

RNG = np.random.default_rng(1234)
NP_DT = {kf.bool: np.bool_, kf.byte: np.uint8, kf.char: np.int8, kf.short: np.int16, kf.int: np.int32,
         kf.long: np.int64, kf.half: np.float16, kf.float: np.float32, kf.double: np.float64}


def rand(shape, dtype=np.float32, lo=-10, hi=10):
    if np.dtype(dtype) == np.bool_:
        return RNG.integers(0, 2, size=shape).astype(np.bool_)
    return RNG.uniform(lo, hi, size=shape).astype(dtype)


def g(arr):
    return kf.from_numpy(arr, 0)


def assert_close(got, exp, rtol=1e-5, atol=1e-5):
    got = got.contiguous().numpy() if not isinstance(got, np.ndarray) else got
    assert got.shape == tuple(exp.shape), (got.shape, exp.shape)
    np.testing.assert_allclose(got.astype(np.float64), np.asarray(exp).astype(np.float64), rtol=rtol, atol=atol)


def assert_equal(got, exp):
    got = got.contiguous().numpy() if not isinstance(got, np.ndarray) else got
    assert got.dtype == exp.dtype, (got.dtype, exp.dtype)
    np.testing.assert_array_equal(got, exp)


# ---------------------------------------------------------------- host I/O
def test_from_numpy_roundtrip_all_dtypes():  # ref: test_tensor.py:10-13
    for dt in (np.bool_, np.uint8, np.int8, np.int16, np.int32, np.int64, np.float16, np.float32, np.float64):
        arr = rand((7, 33), dt)
        assert_equal(g(arr), arr)
    arr = rand((5, 9), np.float32).astype(O.bfloat16)
    t = g(arr)
    assert t.dtype() == kf.bfloat16
    assert_equal(t, arr)
    # non-contiguous numpy input is made contiguous (the reference silently misreads it)
    arr = rand((6, 8)).T
    assert_equal(g(arr), np.ascontiguousarray(arr))
    # empty tensor
    e = g(np.zeros((0, 4), np.float32))
    assert e.numel() == 0 and e.numpy().shape == (0, 4)


def test_zeros_empty_fill_item():
    z = kf.zeros([3, 5], kf.int, 0)
    assert_equal(z, np.zeros((3, 5), np.int32))
    e = kf.empty([4, 4], kf.float, 0)
    e.fill_(2.5)
    assert_equal(e, np.full((4, 4), 2.5, np.float32))
    i = kf.empty([9], kf.int, 0)
    i.fill_(2.9)  # double -> int64 -> int32 truncation (ref: nullary_ops_kernel.cu:19-25)
    assert_equal(i, np.full((9,), 2, np.int32))
    arr = rand((3, 4))
    assert g(arr).item([2, 1]) == pytest.approx(float(arr[2, 1]))
    assert "shape=[3,4]" in repr(g(arr))


# ---------------------------------------------------------------- elementwise
def test_tensor_add():  # ref: test_tensor.py:15-27
    for shape in ((2, 3), (1000,), (12, 11, 3331)):
        a = rand(shape)
        assert_equal(g(a) + g(a), a + a)
        i, f = rand(shape, np.int32), rand(shape)
        assert_equal(g(i) + g(f), O.binary("+", i, f))


def test_inplace_ops_keep_data_ptr():  # ref: test_tensor.py:29-68
    a, b = rand((5, 7, 11)), rand((5, 1, 11))
    ga, gb = g(a), g(b)
    addr = ga.data_ptr()
    for op in ("+", "-", "*", "/"):
        a = O.binary(op, a, np.broadcast_to(b, a.shape))
        if op == "+":
            ga += gb
        elif op == "-":
            ga -= gb
        elif op == "*":
            ga *= gb
        else:
            ga /= gb
        assert ga.data_ptr() == addr
        assert_equal(ga, a)
    for op, s in (("+", 2), ("-", 3), ("*", 4), ("/", 5)):
        a = O.binary_scalar(op, a, s)
        if op == "+":
            ga += s
        elif op == "-":
            ga -= s
        elif op == "*":
            ga *= s
        else:
            ga /= s
        assert ga.data_ptr() == addr
        assert_equal(ga, a)


def test_refcounts_on_device():  # ref: test_tensor.py:70-84
    x = g(rand((3, 4)))
    y = copy.deepcopy(x)
    assert x.data_ptr() == y.data_ptr()
    assert x.storage_ref_count() == y.storage_ref_count() == 1
    assert x.impl_ref_count() == y.impl_ref_count() == 2
    del x
    assert y.impl_ref_count() == 1


@pytest.mark.parametrize("shapes", [([16, 1], [1, 6]), ([162, 1, 345], [162, 6, 1]), ([123, 1, 567], [123, 127, 567]),
                                    ([3, 1, 64, 8], [3, 5, 1, 8]), ([1], [1]), ([0, 3], [1, 3])])
def test_broadcast_binary(shapes):  # ref: test_tensor.py:86-108 ("easy" table + extras)
    for op in "+-*/":
        a, b = rand(shapes[0]), rand(shapes[1], lo=1, hi=10)
        assert_equal(eval(f"g(a) {op} g(b)"), O.binary(op, a, b))
        i = rand(shapes[0], np.int32)
        assert_equal(eval(f"g(i) {op} g(b)"), O.binary(op, i, b))


def test_broadcast_hard_large():  # ref: test_tensor.py:91-92 (1 Gi-element shapes, '+' only)
    a = rand((2, 1024, 1024, 128))
    b = rand((2, 1024, 1, 128))
    assert_equal(g(a) + g(b), a + b)
    assert_equal(g(a) + g(a), a + a)


def test_binary_dtype_matrix():
    """every (dtype, dtype) pair x every op against the bit-exact oracle (promotion + acc-type + cast rules)"""
    dts = [np.bool_, np.uint8, np.int8, np.int16, np.int32, np.int64, np.float16, O.bfloat16, np.float32, np.float64]
    shape = (37, 24)
    for da in dts:
        for db in dts:
            a = rand(shape, np.float32, 1, 9).astype(da) if da != np.bool_ else rand(shape, np.bool_)
            b = rand(shape, np.float32, 1, 9).astype(db) if db != np.bool_ else np.ones(shape, np.bool_)
            for op in "+-*/":
                got = eval(f"g(a) {op} g(b)").numpy()
                exp = O.binary(op, a, b)
                assert got.dtype == exp.dtype, (da, db, op)
                # IEEE ops are exactly rounded on both sides and integer ops wrap identically -> identical bytes
                np.testing.assert_array_equal(got.view(np.uint8), exp.view(np.uint8), err_msg=f"{da} {op} {db}")


def test_scalar_ops_all_dtypes():
    for dt in (np.uint8, np.int16, np.int32, np.int64, np.float16, np.float32, np.float64):
        a = rand((19, 8), np.float32, 1, 9).astype(dt)
        for op, s in (("+", 2.5), ("-", 3.0), ("*", 1.5), ("/", 2.0)):
            got = eval(f"g(a) {op} s")
            assert_equal(got, O.binary_scalar(op, a, s))


def test_unaligned_and_strided_operands():
    a, b = rand((64, 67)), rand((64, 67))
    ga, gb = g(a), g(b)
    assert_equal(ga[:, 1:] + gb[:, 1:], a[:, 1:] + b[:, 1:])          # misaligned rows -> scalar kernel
    assert_equal(ga[::2] * gb[::2], a[::2] * b[::2])                  # row-strided, inner contiguous
    assert_equal(ga.permute(1, 0) - gb.permute(1, 0), a.T - b.T)      # both permuted
    assert_equal(ga[:, ::3] / gb[:, ::3], a[:, ::3] / b[:, ::3])      # inner stride 3
    x = g(a)
    v = x[10:20, 5:50]
    v += 1.0                                                          # in-place through a view
    a2 = a.copy()
    a2[10:20, 5:50] += 1.0
    assert_equal(x, a2)


def test_convert_half_bf16():  # ref: test_tensor.py:148-160
    arr = rand((2, 3), np.float64)
    t = g(arr)
    h = t.half()
    assert h.dtype() == kf.half
    assert_equal(h, arr.astype(np.float32).astype(np.float16))      # double -> float -> half (double rounding)
    t *= t
    h *= h
    assert_close(t, h.float().numpy(), rtol=2e-2, atol=2e-2)
    arr = rand((257, 65), np.float64)
    b = g(arr).bfloat16()
    assert_equal(b, arr.astype(np.float32).astype(O.bfloat16))
    assert_equal(b.float(), arr.astype(np.float32).astype(O.bfloat16).astype(np.float32))
    i = g(rand((33, 9), np.int32, -1000, 1000))
    assert_equal(i.float(), i.numpy().astype(np.float32))
    f = g(rand((33, 9), np.float32, -100, 100))
    assert_equal(f.to(kf.int), np.trunc(f.numpy()).astype(np.int32))


# ---------------------------------------------------------------- copy / permute / views
def test_permute_contiguous():  # ref: test_tensor.py:162-167
    arr = rand((16, 8, 64, 11), np.float64)
    assert_equal(g(arr).permute(2, 1, 0, 3).contiguous(), np.ascontiguousarray(arr.transpose(2, 1, 0, 3)))
    for dt in (np.uint8, np.int16, np.float32, np.float64):
        a = rand((130, 257), np.float32, 0, 100).astype(dt)
        assert_equal(g(a).permute(1, 0).contiguous(), np.ascontiguousarray(a.T))          # tiled transpose
        b = rand((5, 70, 3, 66), np.float32, 0, 100).astype(dt)
        assert_equal(g(b).permute(3, 2, 0, 1).contiguous(), np.ascontiguousarray(b.transpose(3, 2, 0, 1)))
        assert_equal(g(b).permute(0, 3, 2, 1).contiguous(), np.ascontiguousarray(b.transpose(0, 3, 2, 1)))
    big = rand((4096, 4096))
    assert_equal(g(big).permute(1, 0).contiguous(), np.ascontiguousarray(big.T))           # C1 shape


def test_slice_view_cat_split():  # ref: test_tensor.py:233-271
    arr = rand((11, 155, 33, 5), lo=-10000, hi=10000)
    t = g(arr)
    assert_equal(t[3, 3:8, 4:11:2].contiguous(), np.ascontiguousarray(arr[3, 3:8, 4:11:2]))
    arr2 = rand((5, 2, 11, 23))
    assert_equal(g(arr2).view(5, -1, 23).contiguous() + 1, arr2.reshape(5, -1, 23) + np.float32(1))
    a1, a2, a3 = rand((5, 11, 23)), rand((5, 13, 23)), rand((5, 1, 23))
    assert_equal(kf.cat([g(a1), g(a2), g(a3)], 1), np.concatenate([a1, a2, a3], 1))
    assert_equal(kf.cat([g(a1), g(a1)], -1), np.concatenate([a1, a1], -1))
    arr3 = rand((5, 25, 23))
    parts = g(arr3).split([11, 13, 1], 1)
    for p, e in zip(parts, np.split(arr3, [11, 24], axis=1)):
        assert_equal(p, np.ascontiguousarray(e))
    # cat casts to the first tensor's dtype (ref: tensor_shape.cpp:41-70)
    i1 = rand((4, 3), np.int32)
    assert_equal(kf.cat([g(i1), g(rand((4, 2)))], 1)[:, :3].contiguous(), i1)


def test_index_put():  # ref: test_tensor.py:273-284
    arr = rand((13, 15))
    t = g(arr)
    i0 = np.array([0, 5, 1, -1], dtype=np.int64)
    i1 = np.array([0, 11, 1, 0], dtype=np.int64)
    vals = rand((4,))
    t.index_put_([g(i0), g(i1)], g(vals))
    exp = arr.copy()
    exp[i0, i1] = vals
    assert_equal(t, exp)


# ---------------------------------------------------------------- reductions
@pytest.mark.parametrize("op", ["sum", "mean"])
def test_reduce_fp32(op):  # ref: test_tensor.py:110-118 (shape and dims), tolerance: L1-mass 1e-5
    arr = rand((223, 23, 3213))
    t = g(arr)
    for dim in (0, 1, 2, -1):
        got = getattr(t, op)(dim).numpy()
        exact = O.reduce_exact(op, arr, dim)
        assert got.shape == exact.shape
        mass = np.abs(arr.astype(np.float64)).sum(axis=dim, keepdims=True) / (arr.shape[dim] if op == "mean" else 1)
        assert O.l1_tolerance_ok(got, exact, mass, 1e-5), (op, dim, np.abs(got - exact).max())


@pytest.mark.parametrize("shape", [(4096, 4096), (1, 1 << 22), (1 << 22, 1), (3, 5, 7), (64, 1), (1000, 1000, 3), (2, 65536 + 3), (70000, 33)])
def test_reduce_shapes_and_paths(shape):
    arr = rand(shape)
    t = g(arr)
    for dim in range(len(shape)):
        for op in ("sum", "mean"):
            got = getattr(t, op)(dim).numpy()
            exact = O.reduce_exact(op, arr, dim)
            mass = np.abs(arr.astype(np.float64)).sum(axis=dim, keepdims=True) / (arr.shape[dim] if op == "mean" else 1)
            assert O.l1_tolerance_ok(got, exact, mass, 1e-5), (shape, op, dim)
    flat = g(arr).view(-1)
    got = flat.sum(0).numpy()
    assert got.shape == (1,)
    assert abs(got[0] - arr.astype(np.float64).sum()) <= 1e-5 * np.abs(arr).astype(np.float64).sum()


@pytest.mark.parametrize("shape,dt", [((4096, 4096), np.float32), ((3001, 2048), np.float32), ((5000, 1024), np.float32), ((2500, 1000 * 4), np.float32),
                                      ((4096, 4096), "bf16"), ((2100, 2048), np.float16), ((4100, 1024), np.float64)])
def test_reduce_dim0_streaming_kernel(shape, dt):
    """outer == 1 column reduce over whole-row bands with the in-kernel grid barrier (reduce_cols_stream_kernel); called three
    times in a row so the self-resetting arrive / depart counters are exercised, and compared with the cluster kernel"""
    import os
    if dt == "bf16":
        arr = rand(shape, np.float32, -1, 1).astype(O.bfloat16)
    elif dt == np.float16:
        arr = rand(shape, np.float32, -1, 1).astype(np.float16)
    else:
        arr = rand(shape, dt)
    t = g(arr)
    a64 = arr.astype(np.float32).astype(np.float64) if arr.dtype.itemsize == 2 else arr.astype(np.float64)
    mass = np.abs(a64).sum(axis=0, keepdims=True)
    tol = 1e-5 if dt == np.float32 else (1e-12 if dt == np.float64 else 4e-3)  # 16-bit: one output rounding
    for op in ("sum", "mean", "sum"):
        os.environ["KF_RED_STREAM"] = "1"  # opt-in variant (the cluster kernel is the default)
        try:
            got_t = getattr(t, op)(0)
        finally:
            os.environ.pop("KF_RED_STREAM", None)
        got = got_t.float().numpy().astype(np.float64) if arr.dtype.itemsize == 2 else got_t.numpy().astype(np.float64)
        exact = a64.sum(0, keepdims=True) / (shape[0] if op == "mean" else 1)
        m = mass / (shape[0] if op == "mean" else 1)
        assert got.shape == exact.shape
        assert np.all(np.abs(got - exact) <= tol * np.maximum(m, 1e-30)), (shape, dt, op, float(np.abs(got - exact).max()))
    other = t.sum(0)  # default: cluster kernel with the push fold
    os.environ["KF_RED_PUSH"] = "0"
    try:
        pull = t.sum(0)  # cluster kernel, pull fold (cluster.sync + remote reads)
    finally:
        os.environ.pop("KF_RED_PUSH", None)
    pl = pull.float().numpy() if arr.dtype.itemsize == 2 else pull.numpy()
    ot = other.float().numpy() if arr.dtype.itemsize == 2 else other.numpy()
    assert np.array_equal(pl, ot)  # same summation order: bit-identical
    o = other.float().numpy().astype(np.float64) if arr.dtype.itemsize == 2 else other.numpy().astype(np.float64)
    assert np.all(np.abs(o - a64.sum(0, keepdims=True)) <= tol * mass)


def test_reduce_other_dtypes():
    arr = rand((37, 513), np.float64)
    assert_close(g(arr).sum(1), arr.sum(1, keepdims=True), rtol=1e-12, atol=1e-9)
    assert_close(g(arr).mean(0), arr.mean(0, keepdims=True), rtol=1e-12, atol=1e-12)
    for dt in (np.uint8, np.int8, np.int16, np.int32, np.int64, np.bool_):
        a = rand((129, 300), np.float32, -100, 100).astype(dt) if dt != np.bool_ else rand((129, 300), np.bool_)
        for dim in (0, 1):
            assert_equal(g(a).sum(dim), O.reduce_int("sum", a, dim))    # wraps in the input dtype
            assert_equal(g(a).mean(dim), O.reduce_int("mean", a, dim))  # integer-divided factor (reference quirk)
    for dt in (np.float16, O.bfloat16):
        a = rand((64, 4096), np.float32, -1, 1).astype(dt)
        for dim in (0, 1):
            got = g(a).sum(dim).float().numpy().astype(np.float64)
            exact = O.reduce_exact("sum", a, dim)
            mass = np.abs(a.astype(np.float32).astype(np.float64)).sum(axis=dim, keepdims=True)
            assert np.all(np.abs(got - exact) <= 2e-2 * np.maximum(np.abs(exact), 1e-3 * mass) + 1e-2)
    # non-contiguous input
    b = rand((40, 50))
    assert_close(g(b).permute(1, 0).sum(0), b.T.sum(0, keepdims=True), rtol=1e-5, atol=1e-4)


def test_mean_var_and_norm_stat():  # ref: test_tensor.py:120-146
    arr = rand((13, 325, 127), np.float64)
    m, v = g(arr).mean_var(1, False)
    assert_close(m, arr.mean(1, keepdims=True), rtol=1e-10, atol=1e-10)
    assert_close(v, arr.var(1, ddof=1, keepdims=True), rtol=1e-10, atol=1e-10)
    m, s = g(arr).mean_var(2, True)
    assert_close(s, arr.std(2, ddof=1, keepdims=True), rtol=1e-10, atol=1e-10)
    # fp32 over the last dim: the one-pass row-statistics kernel (and the composed path for rows it does not cover)
    for shape in ([37, 512], [5, 7, 4096], [3, 1000], [2, 16384 * 3], [4, 30]):
        a = rand(shape)
        a64 = a.astype(np.float64)
        m, v = g(a).mean_var(-1, False)
        assert_close(m, a64.mean(-1, keepdims=True), rtol=1e-5, atol=1e-5)
        assert_close(v, a64.var(-1, ddof=1, keepdims=True), rtol=1e-5, atol=1e-6)
        m, sd = g(a).mean_var(-1, True)
        assert_close(sd, a64.std(-1, ddof=1, keepdims=True), rtol=1e-5, atol=1e-6)
    for shape in ([64, 64], [1024, 2048]):
        a = rand(shape)
        m, inv = g(a).norm_stat(0)
        assert_close(m, a.mean(0, keepdims=True), rtol=1e-4, atol=1e-5)
        assert_close(inv, 1.0 / np.sqrt(a.astype(np.float64).var(0, keepdims=True)), rtol=1e-4, atol=1e-5)


# ---------------------------------------------------------------- autograd
def test_basic_backward():  # ref: test_tensor.py:286-309
    grad = rand((2, 3))
    a, b, c = g(rand((2, 3))), g(rand((2, 3))), g(rand((2, 3)))
    a.set_requires_grad(True)
    b.set_requires_grad(True)
    ca = c + a
    ab = a + b
    accb = ca + ab
    accba = accb + a
    accba.backward(g(grad))
    assert_equal(a.grad(), grad * 3)
    assert_equal(b.grad(), grad)
    assert not c.grad().defined()
    accba.backward(g(grad))  # leaf grads accumulate across calls (ref: tensor.cpp:75-84)
    assert_equal(a.grad(), grad * 6)


def test_backward_mul_div_sum_mean_broadcast():
    torch = pytest.importorskip("torch")
    x, w, bias = rand((6, 5), lo=1, hi=2), rand((6, 5), lo=1, hi=2), rand((1, 5))
    gx, gw, gb = g(x), g(w), g(bias)
    for t in (gx, gw, gb):
        t.set_requires_grad(True)
    y = ((gx * gw) / (gw + gb) - gx).mean(1) * 3.0
    go = rand((6, 1))
    y.backward(g(go))
    tx, tw, tb = (torch.tensor(v, dtype=torch.float64, requires_grad=True) for v in (x, w, bias))
    ty = ((tx * tw) / (tw + tb) - tx).mean(1, keepdim=True) * 3.0
    ty.backward(torch.tensor(go, dtype=torch.float64))
    assert_close(y, ty.detach().numpy(), rtol=1e-5, atol=1e-5)
    assert_close(gx.grad(), tx.grad.numpy(), rtol=1e-4, atol=1e-5)
    assert_close(gw.grad(), tw.grad.numpy(), rtol=1e-4, atol=1e-5)
    assert_close(gb.grad(), tb.grad.numpy(), rtol=1e-4, atol=1e-5)


def test_pool_reuses_memory_stream_ordered():
    kf.synchronize()
    a = g(rand((1024, 1024)))
    _, _, mallocs0 = kf.mem_stats()
    ptrs = set()
    for _ in range(50):
        t = a + a
        ptrs.add(t.data_ptr())
        del t
    _, _, mallocs1 = kf.mem_stats()
    assert mallocs1 == mallocs0 or mallocs1 == mallocs0 + 1
    assert len(ptrs) <= 2
