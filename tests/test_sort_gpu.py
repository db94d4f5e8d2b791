"""GPU parity for sort / top-k: values AND int64 indices bit-exact against the oracle (stable order, ties keep
ascending index) and against the reference build's own outputs (tests/golden/ref_outputs.npz)."""
import os

import numpy as np
import pytest

import kfunca_b200 as kf
from oracle import oracle as O
from oracle.golden_cases import cases

pytestmark = pytest.mark.gpu
RNG = np.random.default_rng(4321)
GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_outputs.npz"))


def g(a):
    return kf.from_numpy(a, 0)


def check_sort(x, dim, desc):
    v, i = g(x).sort(dim, desc)
    ev, ei = O.sort(x, dim, desc)
    assert i.dtype() == kf.long
    np.testing.assert_array_equal(v.numpy().view(np.uint8), ev.view(np.uint8))
    np.testing.assert_array_equal(i.numpy(), ei)


def check_topk(x, k, dim, largest):
    v, i = g(x).topk(k, dim, largest)
    ev, ei = O.topk(x, k, dim, largest)
    np.testing.assert_array_equal(v.numpy().view(np.uint8), ev.view(np.uint8))
    np.testing.assert_array_equal(i.numpy(), ei)


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int32])
def test_sort_small_slice(dtype):  # ref: test_tensor.py:169-192
    for shape in ([2, 3, 4], [23, 11, 23], [11, 23, 64], [13, 65, 1049], [5, 11, 22223]):
        for dim in (2, 1, 0):
            for desc in (False, True):
                check_sort(RNG.uniform(-1000, 1000, size=shape).astype(dtype), dim, desc)


def test_sort_ties_and_special_values():
    x = np.round(RNG.uniform(-5, 5, (7, 3000))).astype(np.float32)  # heavy ties, includes -0.0
    x[0, :10] = [np.inf, -np.inf, np.nan, -0.0, 0.0, np.nan, 1.0, -1.0, np.inf, -np.inf]
    for desc in (False, True):
        check_sort(x, 1, desc)
        check_sort(x[:, :300].copy(), 1, desc)
        check_sort(np.ascontiguousarray(x.T), 0, desc)


@pytest.mark.parametrize("dtype", [np.uint8, np.int8, np.int16, np.int64, np.float16, "bf16"])
def test_sort_other_dtypes(dtype):
    dt = O.bfloat16 if dtype == "bf16" else dtype
    for n in (100, 5000):
        x = RNG.uniform(-100, 100, (9, n)).astype(np.float32).astype(dt)
        check_sort(x, 1, False)
        check_sort(x, 1, True)


def test_sort_large_slice():  # ref: test_tensor.py:194-201
    x = RNG.uniform(-1000, 1000, size=(4, 1024000)).astype(np.float32)
    v, i = g(x).sort(1, False)
    np.testing.assert_array_equal(v.numpy(), np.sort(x, axis=1))
    np.testing.assert_array_equal(i.numpy(), np.argsort(x, axis=1, kind="stable"))


def test_sort_edge_shapes():
    check_sort(RNG.uniform(-1, 1, (5, 1)).astype(np.float32), 1, False)
    check_sort(RNG.uniform(-1, 1, (1, 4096)).astype(np.float32), 1, True)
    check_sort(RNG.uniform(-1, 1, (3, 4097)).astype(np.float32), 1, True)
    v, i = g(np.zeros((0, 7), np.float32)).sort(1, False)
    assert v.numpy().shape == (0, 7) and i.numpy().shape == (0, 7)
    with pytest.raises(RuntimeError):
        g(np.zeros((3, 3), np.bool_)).sort(0, False)  # ref: sort_ops_kernel.cu:565-566


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int32])
def test_topk_small(dtype):  # ref: test_tensor.py:203-222 — we check indices too, the reference test cannot
    for shape in ([13, 65, 1049], [33, 22, 22223]):
        for dim in (2, 1, 0):
            for largest in (False, True):
                check_topk(RNG.uniform(-100000, 100000, size=shape).astype(dtype), 8, dim, largest)


def test_topk_large():  # ref: test_tensor.py:224-231
    x = RNG.uniform(-10000, 10000, size=(4, 1024000)).astype(np.float32)
    for k in (64, 2049, 22223):
        check_topk(x, k, 1, True)


def test_topk_select_path_ties_and_overflow():
    # vocab-style rows that take the register-resident select kernel (2048 <= n <= 32768)
    for n, k in ((32768, 64), (32768, 1), (32768, 1024), (30000, 64), (2048, 100), (5001, 33)):
        check_topk(RNG.uniform(-1e5, 1e5, (67, n)).astype(np.float32), k, 1, True)
        check_topk(RNG.uniform(-1e5, 1e5, (5, n)).astype(np.float32), k, 1, False)
    ties = np.round(RNG.uniform(-3, 3, (9, 32768))).astype(np.float32)      # ~4700 copies of each value -> overflow rows
    check_topk(ties, 64, 1, True)
    check_topk(ties, 700, 1, False)
    const = np.full((3, 8192), 1.5, np.float32)                               # all equal: indices must be 0..k-1
    v, i = g(const).topk(64, 1, True)
    np.testing.assert_array_equal(i.numpy(), np.tile(np.arange(64), (3, 1)))
    check_topk(RNG.integers(-50, 50, (11, 16384)).astype(np.int32), 64, 1, True)
    check_topk(RNG.uniform(-9, 9, (4, 4096)).astype(np.float64), 17, 1, True)
    check_topk(RNG.uniform(-9, 9, (4, 9000)).astype(np.float16), 17, 1, False)


def test_topk_full_size_property():
    """C4-shaped rows at reduced row count: the k-th value bounds every unselected element, indices point at the values."""
    rows, n, k = 2048, 32768, 64
    x = RNG.uniform(-1e5, 1e5, (rows, n)).astype(np.float32)
    v, i = g(x).topk(k, 1, True)
    v, i = v.numpy(), i.numpy()
    assert np.all(np.take_along_axis(x, i, 1) == v)
    assert np.all(np.diff(v, axis=1) <= 0)
    kth = v[:, -1:]
    assert np.all((x > kth).sum(1) <= k - 1)
    ev, ei = O.topk(x[:64], k, 1, True)
    np.testing.assert_array_equal(i[:64], ei)


@pytest.mark.parametrize("name", [n for n, k, _, _ in cases() if k in ("sort", "topk")])
def test_against_reference_outputs(name):
    kind, inp, prm = next((k, i, p) for n, k, i, p in cases() if n == name)
    t = g(inp["x"])
    v, i = t.sort(prm["dim"], prm["descending"]) if kind == "sort" else t.topk(prm["k"], prm["dim"], prm["largest"])
    np.testing.assert_array_equal(v.numpy(), GOLD[f"{name}.values"])
    np.testing.assert_array_equal(i.numpy(), GOLD[f"{name}.indices"])
