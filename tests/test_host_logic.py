"""CPU-only tests of the host logic behind the C ABI: symbols, view algebra, planner, pool allocator.
No compute calls (there is no CPU compute path)."""
import ctypes
import os
import re

import numpy as np
import pytest

import kfunca_b200 as kf
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "kfunca_b200.h")).read()
    names = sorted(set(re.findall(r"\b(kf_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) > 70
    lib = ctypes.CDLL(kf.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    lib.kf_last_error.restype = ctypes.c_char_p
    major, minor = ctypes.c_int(), ctypes.c_int()
    assert lib.kf_version(ctypes.byref(major), ctypes.byref(minor)) == 0


def test_c_abi_error_reporting():
    lib = ctypes.CDLL(kf.LIB_PATH)
    lib.kf_last_error.restype = ctypes.c_char_p
    out = ctypes.c_void_p()
    shape = (ctypes.c_int64 * 1)(4)
    assert lib.kf_empty(shape, 1, 99, -1, ctypes.byref(out)) != 0
    assert b"bad dtype" in lib.kf_last_error()
    n = ctypes.c_int()
    assert lib.kf_dim(None, ctypes.byref(n)) != 0
    assert b"null tensor handle" in lib.kf_last_error()


def test_dtype_enum_matches_reference_order():
    # ref: scalar_type.h:9-27 / register.cpp:64-75
    assert [int(x) for x in (kf.bool, kf.byte, kf.char, kf.short, kf.int, kf.long, kf.half, kf.bfloat16, kf.float, kf.double)] == list(range(10))


def test_promotion_table_matches_oracle():
    codes = [kf.bool, kf.byte, kf.char, kf.short, kf.int, kf.long, kf.half, kf.bfloat16, kf.float, kf.double]
    for a in codes:
        for b in codes:
            assert int(kf.promote_types(a, b)) == O.promote(int(a), int(b)), (a, b)
    # the reference's quirks: unsigned (+) signed -> the signed type; half + bf16 -> bf16
    assert kf.promote_types(kf.byte, kf.char) == kf.char
    assert kf.promote_types(kf.half, kf.bfloat16) == kf.bfloat16
    assert kf.promote_types(kf.int, kf.float) == kf.float
    assert kf.promote_types(kf.long, kf.half) == kf.half


def meta(shape, dt=None):
    return kf.empty(list(shape), dt or kf.float, -1)


def np_like(t, base):
    """numpy view with the same shape/strides/offset over `base` (flat array)"""
    return np.lib.stride_tricks.as_strided(base[t.storage_offset():], shape=t.sizes(), strides=[s * base.itemsize for s in t.strides()])


def test_view_algebra_matches_numpy():
    shape = (11, 155, 33, 5)
    base = np.arange(np.prod(shape), dtype=np.float32)
    arr = base.reshape(shape)
    t = meta(shape)
    assert t.is_contiguous() and t.sizes() == list(shape)
    cases = [
        (t[3, 3:8, 4:11:2], arr[3, 3:8, 4:11:2]),
        (t[-1], arr[-1]),
        (t[:, ::7], arr[:, ::7]),
        (t[2:100, 150:999], arr[2:100, 150:999]),
        (t.permute(2, 1, 0, 3), arr.transpose(2, 1, 0, 3)),
        (t.permute(3, -2, 0, 1)[1:, 5], arr.transpose(3, 2, 0, 1)[1:, 5]),
        (t.view(11, -1, 5), arr.reshape(11, -1, 5)),
        (t.view(-1), arr.reshape(-1)),
    ]
    for got, exp in cases:
        assert got.sizes() == list(exp.shape)
        np.testing.assert_array_equal(np_like(got, base), exp)
    parts = t.split([11, 13, 131], 1)
    exp_parts = np.split(arr, [11, 24], axis=1)
    for g, e in zip(parts, exp_parts):
        np.testing.assert_array_equal(np_like(g, base), e)
    # views share storage, each is a new impl (ref: tensor.cpp:167-174)
    assert t.storage_ref_count() >= 2
    assert not t.permute(1, 0, 2, 3).is_contiguous()
    assert t[0].is_contiguous() and t[:, 0].is_contiguous() is False


def test_view_errors():
    t = meta((4, 6))
    with pytest.raises(RuntimeError):
        t.permute(0, 0)
    with pytest.raises(RuntimeError):
        t.permute(0)
    with pytest.raises(RuntimeError):
        t.view(5, -1)
    with pytest.raises(RuntimeError):
        t.permute(1, 0).view(24)  # view needs contiguity (ref: tensor.cpp:269-270)
    with pytest.raises(RuntimeError):
        t.split([1, 2], 1)  # sizes must sum to the dim (ref: tensor_shape.cpp:88)
    with pytest.raises(RuntimeError):
        t[7]
    with pytest.raises(RuntimeError):
        kf.empty([1] * 13, kf.float, -1)  # MAX_TENSOR_DIMS = 12


def test_handle_refcounts():
    import copy

    x = meta((3, 4))
    y = copy.deepcopy(x)
    assert x.storage_ref_count() == y.storage_ref_count() == 1
    assert x.impl_ref_count() == y.impl_ref_count() == 2
    del x
    assert y.impl_ref_count() == 1


def test_planner_broadcast_and_collapse():
    # contiguous same-shape operands collapse to 1-D (ref: coalesce_dimensions, tensor_iterator.cpp:263-307)
    a, b = meta((4096, 4096)), meta((4096, 4096))
    shape, strides, common = kf.debug_plan_binary(a, b)
    assert shape == [4096 * 4096] and strides == [[4], [4], [4]] and common == kf.float
    # inner broadcast: [2,1024,1024,512] + [2,1024,1,512]
    a, b = meta((2, 1024, 1024, 512)), meta((2, 1024, 1, 512))
    shape, strides, _ = kf.debug_plan_binary(a, b)
    assert shape == [512, 1024, 2048]
    assert strides[2] == [4, 0, 2048]
    # mixed dtype promotes, strides are in bytes of each operand
    a, b = meta((5, 7), kf.int), meta((5, 1), kf.double)
    shape, strides, common = kf.debug_plan_binary(a, b)
    assert common == kf.double and shape == [7, 5] and strides[1] == [4, 28] and strides[2] == [0, 8]
    # permuted input keeps two dims, output fastest dim first
    a = meta((8, 16)).permute(1, 0)
    shape, strides, _ = kf.debug_plan_binary(a, meta((16, 8)))
    assert shape == [8, 16] and strides[1] == [64, 4]
    with pytest.raises(RuntimeError):
        kf.debug_plan_binary(meta((3, 4)), meta((4,)))  # equal ndim required (ref: tensor_iterator.cpp:17-30)
    with pytest.raises(RuntimeError):
        kf.debug_plan_binary(meta((3, 4)), meta((2, 4)))


def test_pool_allocator_reuse_split_coalesce():
    MiB = 1 << 20
    # alloc A(1000) B(3000); free A; alloc 600 must reuse A's slot (best fit), no new arena
    offs, (in_use, reserved, mallocs) = kf.debug_pool_trace([1000, 3000, -1, 600])
    assert mallocs == 1 and reserved == 2 * MiB
    assert offs[3] == offs[0]
    assert in_use == 3072 + 1024
    # neighbours coalesce: after freeing both, one block of the full arena serves a 2 MiB-class request
    offs, (in_use, reserved, mallocs) = kf.debug_pool_trace([1000, 3000, -1, -2, 1 * MiB])
    assert mallocs == 1 and offs[4] == offs[0]
    # large blocks: 64 MiB x3 like C1, free all, re-allocate -> no new device mallocs
    big = 64 * MiB
    offs, (in_use, reserved, mallocs) = kf.debug_pool_trace([big, big, big, -1, -2, -3, big, big, big])
    assert mallocs == 3 and sorted(offs[6:]) == sorted(offs[:3]) and in_use == 3 * big
    # empty_cache returns only fully-free arenas
    offs, (in_use, reserved, mallocs) = kf.debug_pool_trace([big, 1000, -1, 0])
    assert reserved == 2 * MiB and in_use == 1024


def test_pool_fences_side_streams_on_release():
    """record_stream: a block used on other streams is fenced (library stream waits for each such stream, once) when it is released;
    blocks nobody announced are released without a fence; the record dies with the block (its next tenant starts clean)."""
    A, F, R = 0, 1, 2
    log, n = kf.debug_pool_fences([A, R, R, R, F], [4096, 0, 0, 0, 0], [0, 11, 22, 11, 0])
    assert n == 2 and log == [11, 22]
    log, n = kf.debug_pool_fences([A, A, R, F, F], [4096, 4096, 1, 0, 1], [0, 0, 7, 0, 0])
    assert log == [7]  # only the announced block is fenced
    # released + reallocated at the same address: the new tenant carries no stale streams
    log, n = kf.debug_pool_fences([A, R, F, A, F], [4096, 0, 0, 4096, 3], [0, 5, 0, 0, 0])
    assert log == [5] and n == 1
    with pytest.raises(RuntimeError):
        kf.debug_pool_fences([R], [0], [1])


def test_pool_grows_in_slabs_once_it_is_large():
    """>= 1 GiB reserved: a miss reserves a quarter of the pool (not the exact request), so a repeating allocation pattern
    stops calling the driver (steady state of a training step)."""
    MiB = 1 << 20
    big = 64 * MiB
    offs, (in_use, reserved, mallocs) = kf.debug_pool_trace([big] * 16)
    assert mallocs == 16 and reserved == 16 * big  # below 1 GiB: exact-size arenas
    offs, (in_use, reserved, mallocs) = kf.debug_pool_trace([big] * 17)
    assert mallocs == 17 and reserved == 16 * big + 256 * MiB  # the 17th miss takes 1024 / 4 MiB
    offs, (in_use, reserved, mallocs) = kf.debug_pool_trace([big] * 20)
    assert mallocs == 17 and in_use == 20 * big  # ... which serves the next three requests without the driver
    # a repeating "step" (allocate a mixed set, free it all) settles: the second and third pass add no driver calls
    step = [big, 300 * MiB, 3 * MiB, big, 700 * MiB, 1000, big, 128 * MiB]
    ops, base = [], 0
    counts = []
    for rep in range(3):
        ops += step
        ops += [-(base + i + 1) for i in range(len(step))]
        base += 2 * len(step)
        counts.append(kf.debug_pool_trace(ops)[1][2])
    assert counts[1] == counts[0] == counts[2]


def test_compute_without_gpu_fails_loudly():
    a = meta((2, 2))
    for fn in (lambda: a + a, lambda: a.sum(0), lambda: a.contiguous() if False else a.permute(1, 0).contiguous(),
               lambda: a.sort(0, False), lambda: kf.gemm(a, a, 1.0, 0.0)):
        with pytest.raises(RuntimeError):
            fn()


def test_leaf_grad_hook_registration_needs_no_gpu():
    """installing / removing the data-parallel leaf-gradient hook is pure host state (the hook only fires inside backward())"""
    calls = []
    kf.set_leaf_grad_hook(lambda leaf, grad: calls.append(1))
    kf.set_leaf_grad_hook(None)
    kf.set_leaf_grad_hook(None)  # idempotent
    assert calls == []
