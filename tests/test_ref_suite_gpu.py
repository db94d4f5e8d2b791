"""The reference's own Python test-suite, re-stated test for test against the drop-in module.

Every test below carries the name, shapes, value ranges, comparison partner (NumPy / torch-CPU) and tolerance of the test it
restates in /root/reference/test/{test_tensor,test_gemm,test_nn}.py (cited per test).  The module under test is imported exactly
the way a kfunca user would after switching: `import kfunca_b200 as kfunca` — nothing else in the test bodies knows about this
repo.  Differences from the originals: inputs are seeded, loops are pytest parameters, and the 4 GiB "hard" broadcast case keeps
only the fp32 operand pair (the int32 variant doubles a 4 GiB host allocation without reaching new code).
"""
import copy

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import kfunca_b200 as kfunca

pytestmark = pytest.mark.gpu
RNG = np.random.default_rng(314159)


def assert_allclose(a, b, atol=1e-3, rtol=1e-3):  # ref: test/common.py:6-11 (its rtol / atol swap is harmless: both 1e-3)
    if not isinstance(a, np.ndarray):
        a = a.contiguous().numpy()
    if not isinstance(b, np.ndarray):
        b = b.contiguous().numpy()
    assert np.allclose(a, b, rtol=atol, atol=rtol)


def U(lo, hi, shape, dtype=None):
    x = RNG.uniform(lo, hi, size=shape)
    return x if dtype is None else x.astype(dtype)


def test_tensor_impl():  # ref: test_tensor.py:10-13
    arr = U(-10, 10, (2, 3))
    assert_allclose(arr, kfunca.from_numpy(arr, 0))


@pytest.mark.parametrize("shape", [(2, 3), (1000,), (12, 11, 3331)])
def test_tensor_add(shape):  # ref: test_tensor.py:15-28
    arr = U(-10, 10, shape, np.float32)
    a = kfunca.from_numpy(arr, 0)
    assert_allclose(arr + arr, (a + a).numpy())
    arr1, arr2 = U(-10, 10, shape, np.int32), U(-10, 10, shape, np.float32)
    assert_allclose(arr1 + arr2, kfunca.from_numpy(arr1, 0) + kfunca.from_numpy(arr2, 0))


def test_inplace_op():  # ref: test_tensor.py:30-68 — in-place ops keep data_ptr()
    arr1, arr2 = U(-10, 10, (5, 7, 11), np.float32), U(-10, 10, (5, 1, 11), np.float32)
    t1, t2 = kfunca.from_numpy(arr1, 0), kfunca.from_numpy(arr2, 0)
    addr = t1.data_ptr()
    for op in ("+", "-", "*", "/"):
        if op == "+":
            arr1 += arr2; t1 += t2
        elif op == "-":
            arr1 -= arr2; t1 -= t2
        elif op == "*":
            arr1 *= arr2; t1 *= t2
        else:
            arr1 /= arr2; t1 /= t2
        assert addr == t1.data_ptr()
        assert_allclose(arr1, t1)
    arr1 += 2; t1 += 2
    assert addr == t1.data_ptr(); assert_allclose(arr1, t1)
    arr1 -= 3; t1 -= 3
    assert addr == t1.data_ptr(); assert_allclose(arr1, t1)
    arr1 *= 4; t1 *= 4
    assert addr == t1.data_ptr(); assert_allclose(arr1, t1)
    arr1 /= 5; t1 /= 5
    assert addr == t1.data_ptr(); assert_allclose(arr1, t1)


def test_data_ptr():  # ref: test_tensor.py:70-84 — handle copies share the impl, storage is counted once
    arr = U(-10, 10, (3, 4), np.float32)
    x = kfunca.from_numpy(arr, 0)
    y = kfunca.from_numpy(arr, 0)
    y = x
    z = copy.deepcopy(x)
    assert x.data_ptr() == y.data_ptr() == z.data_ptr()
    assert x.storage_ref_count() == y.storage_ref_count() == z.storage_ref_count() == 1
    assert x.impl_ref_count() == y.impl_ref_count() == z.impl_ref_count() == 2
    del x
    assert z.impl_ref_count() == 2 and y.impl_ref_count() == 2
    del y
    assert z.impl_ref_count() == 1


EASY = [([16, 1], [1, 6]), ([162, 1, 345], [162, 6, 1]), ([123, 1, 567], [123, 127, 567])]


@pytest.mark.parametrize("op", ["+", "-", "*", "/"])
@pytest.mark.parametrize("sa,sb", EASY)
def test_broadcast_basic_binary_easy(sa, sb, op):  # ref: test_tensor.py:86-108 ('easy' rows, fp32 and int32 x fp32)
    for dta in (np.float32, np.int32):
        a, b = U(-10, 10, sa, dta), U(-10, 10, sb, np.float32)
        ta, tb = kfunca.from_numpy(a, 0), kfunca.from_numpy(b, 0)
        want = {"+": a + b, "-": a - b, "*": a * b, "/": a / b}[op]
        got = {"+": ta + tb, "-": ta - tb, "*": ta * tb, "/": ta / tb}[op]
        assert_allclose(want, got)


@pytest.mark.parametrize("sb", [[2, 1024, 1, 512], [2, 1024, 1024, 512]])
def test_broadcast_basic_binary_hard(sb):  # ref: test_tensor.py:90-91 ('hard' rows run '+' only; 4 GiB operands)
    sa = [2, 1024, 1024, 512]
    a = RNG.random(sa, dtype=np.float32) * 20 - 10
    b = RNG.random(sb, dtype=np.float32) * 20 - 10
    got = (kfunca.from_numpy(a, 0) + kfunca.from_numpy(b, 0)).numpy()
    a += b
    assert np.allclose(a, got, rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("op", ["sum", "mean"])
@pytest.mark.parametrize("dim", [0, 1, 2])
def test_reduce(op, dim):  # ref: test_tensor.py:110-118
    arr = U(-10, 10, [223, 23, 3213], np.float32)
    want = getattr(np, op)(arr, axis=dim, keepdims=True)
    got = getattr(kfunca.from_numpy(arr, 0), op)(dim)
    assert_allclose(want, got, atol=1e-2, rtol=1e-2)


def test_mean_std():  # ref: test_tensor.py:120-132 (fp64, dim 1)
    shape, dim = (13, 325, 127), 1
    t = kfunca.from_numpy(U(-10, 10, shape), 0)
    mean = t.mean(dim)
    var = ((t - mean) * (t - mean)).sum(dim) / (shape[dim] - 1)
    m2, v2 = t.mean_var(dim, False)
    assert_allclose(mean, m2, atol=1e-2, rtol=1e-2)
    assert_allclose(var, v2, atol=1e-2, rtol=1e-2)
    kfunca.memstat()


@pytest.mark.parametrize("shape", [[64, 64], [1024, 2048], [4096, 4096], [4096 * 4 + 3, 4096 * 4 + 3]])
def test_norm_stat(shape):  # ref: test_tensor.py:134-146
    arr = RNG.random(shape, dtype=np.float32) * 20 - 10
    mean = np.mean(arr, axis=0, keepdims=True)
    var = np.sum((arr - mean) * (arr - mean), axis=0, keepdims=True)
    invstd = 1.0 / np.sqrt(var / shape[0])
    m, i = kfunca.from_numpy(arr, 0).norm_stat(0)
    assert_allclose(mean, m)
    assert_allclose(invstd, i)


def test_convert():  # ref: test_tensor.py:148-160
    # The reference draws six unseeded values and asks (x_f64)^2 ~ half(x)^2 within 1e-3 + 1e-3 |.|, which correctly rounded fp16
    # arithmetic itself misses for ~1 in 3 draws (|x| near 4: 2^-11 input rounding, doubled by the square, plus the output rounding).
    # Keep its assertion, on a draw that exact IEEE half arithmetic satisfies, and pin our result to that arithmetic bit for bit.
    while True:
        x = U(-10, 10, (2, 3))
        exact_half = (x.astype(np.float16) * x.astype(np.float16)).astype(np.float32)
        if np.allclose(x * x, exact_half, rtol=1e-3, atol=1e-3):
            break
    t = kfunca.from_numpy(x, 0)
    h = t.half()
    t *= t
    h *= h
    assert np.array_equal(h.float().numpy(), exact_half)
    assert_allclose(t, h.float())
    t = kfunca.from_numpy(U(-10, 10, (2, 3)), 0)
    b = t.bfloat16()
    t *= t
    b *= b
    assert_allclose(t, b.float(), atol=1e-1, rtol=1e-1)


def test_permute():  # ref: test_tensor.py:162-167
    arr = U(-10, 10, (16, 8, 64, 11))
    assert_allclose(kfunca.from_numpy(arr, 0).permute(2, 1, 0, 3).contiguous(), arr.transpose(2, 1, 0, 3))


SORT_SHAPES = [[2, 3, 4], [23, 11, 23], [11, 23, 64], [13, 65, 1049], [5, 11, 22223]]


@pytest.mark.parametrize("dtype", [np.float32, np.double, np.int32])
@pytest.mark.parametrize("descending", [False, True])
@pytest.mark.parametrize("dim", [2, 1, 0])
def test_sort_small_slice(dtype, descending, dim):  # ref: test_tensor.py:169-192 — values AND indices vs torch.sort(stable=True)
    for shape in SORT_SHAPES:
        arr = U(-1000, 1000, shape, dtype)
        res, ind = torch.sort(torch.from_numpy(arr), dim=dim, descending=descending, stable=True)
        gres, gind = kfunca.from_numpy(arr, 0).sort(dim, descending)
        assert np.array_equal(gres.numpy(), res.numpy())
        assert np.array_equal(gind.numpy(), ind.numpy())


def test_sort_large_slice():  # ref: test_tensor.py:194-201
    arr = U(-1000, 1000, (4, 1024000), np.float32)
    gres, gind = kfunca.from_numpy(arr, 0).sort(1, False)
    assert np.array_equal(gres.numpy(), np.sort(arr, axis=1))
    assert np.array_equal(gind.numpy(), np.argsort(arr, axis=1, kind="stable"))


@pytest.mark.parametrize("dtype", [np.float32, np.double, np.int32])
@pytest.mark.parametrize("largest", [False, True])
@pytest.mark.parametrize("dim", [2, 1, 0])
def test_topk_small(dtype, largest, dim):  # ref: test_tensor.py:203-221 (values vs torch.topk, k = 8)
    for shape in ([13, 65, 1049], [33, 22, 22223]):
        arr = U(-100000, 100000, shape, dtype)
        res, _ = torch.topk(torch.from_numpy(arr), 8, dim=dim, largest=largest)
        gres, gind = kfunca.from_numpy(arr, 0).topk(8, dim, largest)
        assert np.array_equal(gres.numpy(), res.numpy())
        assert np.array_equal(np.take_along_axis(arr, gind.numpy(), axis=dim), res.numpy())


@pytest.mark.parametrize("k", [2049, 22223])
def test_topk_large(k):  # ref: test_tensor.py:223-230
    arr = U(-10000, 10000, (4, 1024000), np.float32)
    res, _ = torch.topk(torch.from_numpy(arr), k, dim=1, largest=True)
    gres, _ = kfunca.from_numpy(arr, 0).topk(k, 1, True)
    assert np.array_equal(gres.numpy(), res.numpy())


def test_tensor_slice():  # ref: test_tensor.py:232-238
    arr = U(-10000, 10000, (11, 155, 33, 5), np.float32)
    assert_allclose(torch.from_numpy(arr)[3, 3:8, 4:11:2], kfunca.from_numpy(arr, 0)[3, 3:8, 4:11:2].contiguous())


def test_view():  # ref: test_tensor.py:240-246
    arr = U(-10000, 10000, (5, 2, 11, 23), np.float32)
    want = torch.from_numpy(arr).view(5, -1, 23).contiguous() + 1
    got = kfunca.from_numpy(arr, 0).view(5, -1, 23).contiguous() + 1
    assert_allclose(want, got)


def test_cat():  # ref: test_tensor.py:248-260
    parts = [U(-10000, 10000, (5, n, 23), np.float32) for n in (11, 13, 1)]
    got = kfunca.cat([kfunca.from_numpy(p, 0) for p in parts], 1)
    assert_allclose(torch.cat([torch.from_numpy(p) for p in parts], 1), got)


def test_split():  # ref: test_tensor.py:262-271
    arr = U(-10000, 10000, (5, 25, 23), np.float32)
    want = torch.from_numpy(arr).split([11, 13, 1], 1)
    got = kfunca.from_numpy(arr, 0).split([11, 13, 1], 1)
    for w, g_ in zip(want, got):
        assert_allclose(w, g_)


def test_index_put():  # ref: test_tensor.py:273-284
    arr = U(-10000, 10000, (13, 15), np.float32)
    t = kfunca.from_numpy(arr, 0)
    indices = [kfunca.from_numpy(np.array([0, 5, 1, 2]).astype("q"), 0), kfunca.from_numpy(np.array([0, 11, 1, 0]).astype("q"), 0)]
    values = kfunca.from_numpy(U(-10000, 10000, (4,), np.float32), 0)
    t.index_put_(indices, values)
    want = torch.from_numpy(arr)
    want.index_put_([torch.from_numpy(i.numpy()) for i in indices], torch.from_numpy(values.numpy()))
    assert_allclose(t, want)


def test_basic_backward():  # ref: test_tensor.py:286-309 — fan-in of three uses of `a`
    grad_ = U(-10, 10, (2, 3), np.float32)
    grad = kfunca.from_numpy(grad_, 0)
    a, b, c = (kfunca.from_numpy(U(-10, 10, (2, 3), np.float32), 0) for _ in range(3))
    a.set_requires_grad(True)
    b.set_requires_grad(True)
    out = ((c + a) + (a + b)) + a
    out.backward(grad)
    assert_allclose(a.grad(), grad * 3)
    assert_allclose(b.grad(), grad)


def test_gemm_base():  # ref: test_gemm.py:9-17 (fp64)
    a, b = U(-10, 10, (123, 457)), U(-10, 10, (457, 234))
    assert_allclose(np.matmul(a, b), kfunca.gemm(kfunca.from_numpy(a, 0), kfunca.from_numpy(b, 0), 1.0, 0.0))


@pytest.mark.parametrize("B,H,Sq,Skv,D", [(2, 4, 32, 256, 128), (3, 5, 64, 32, 64), (5, 16, 65, 33, 123)])
def test_causal_attention(B, H, Sq, Skv, D):  # ref: test_nn.py:11-33 (fp32, U(-10, 10), vs torch SDPA is_causal)
    q_, k_, v_ = U(-10, 10, (B, H, Sq, D), np.float32), U(-10, 10, (B, H, Skv, D), np.float32), U(-10, 10, (B, H, Skv, D), np.float32)
    out = kfunca.causal_attention(kfunca.from_numpy(q_, 0), kfunca.from_numpy(k_, 0), kfunca.from_numpy(v_, 0)).numpy()
    ref = F.scaled_dot_product_attention(torch.from_numpy(q_), torch.from_numpy(k_), torch.from_numpy(v_), is_causal=True).numpy()
    assert_allclose(out, ref)
