"""TEST INFRASTRUCTURE ONLY — the seeded inputs behind tests/golden/ref_*.npz.

Shapes and value ranges follow the reference's own tests (test/test_tensor.py, test_gemm.py, test_nn.py);
the reference draws them unseeded, here they are seeded so that the outputs of the reference build can be
stored once (oracle/make_golden_from_ref.py, run on a B200) and replayed anywhere."""
from __future__ import annotations

import numpy as np


def cases():
    """yield (name, kind, inputs: dict[str, ndarray], params: dict)"""
    rng = np.random.default_rng(20261017)
    u = lambda shape, lo=-10, hi=10, dt=np.float32: rng.uniform(lo, hi, size=shape).astype(dt)
    # elementwise incl. int32 + fp32 promotion and broadcast (test_tensor.py:15-27,86-108)
    for op in "+-*/":
        yield f"bin_f32_{ord(op)}", "binary", {"a": u((162, 1, 45)), "b": u((162, 6, 1), 1, 10)}, {"op": op}
        yield f"bin_i32f32_{ord(op)}", "binary", {"a": u((12, 11, 331), dt=np.int32), "b": u((12, 11, 331), 1, 10)}, {"op": op}
    yield "bin_i32_add", "binary", {"a": u((33, 65), dt=np.int32), "b": u((33, 65), dt=np.int32)}, {"op": "+"}
    yield "bin_i64_div", "binary", {"a": u((33, 65), -1000, 1000, np.int64), "b": u((33, 65), 1, 9, np.int64)}, {"op": "/"}
    # reductions (test_tensor.py:110-118)
    x = u((223, 23, 213))
    for dim in (0, 1, 2):
        yield f"sum_f32_{dim}", "reduce", {"x": x}, {"op": "sum", "dim": dim}
        yield f"mean_f32_{dim}", "reduce", {"x": x}, {"op": "mean", "dim": dim}
    xi = u((37, 300), -100, 100, np.int32)
    yield "sum_i32_1", "reduce", {"x": xi}, {"op": "sum", "dim": 1}
    yield "mean_i32_1", "reduce", {"x": xi}, {"op": "mean", "dim": 1}
    # permute (test_tensor.py:162-167)
    yield "permute_f64", "permute", {"x": u((16, 8, 64, 11), dt=np.float64)}, {"dims": (2, 1, 0, 3)}
    # sort / topk incl. ties (test_tensor.py:169-222); small integer-valued floats force ties
    xs = np.round(u((13, 65, 149), -20, 20)).astype(np.float32)
    for dim in (0, 1, 2):
        for desc in (False, True):
            yield f"sort_f32_{dim}_{int(desc)}", "sort", {"x": xs}, {"dim": dim, "descending": desc}
    yield "sort_i32", "sort", {"x": u((5, 11, 2223), -50, 50, np.int32)}, {"dim": 2, "descending": True}
    yield "sort_f64", "sort", {"x": u((11, 23, 64), -1000, 1000, np.float64)}, {"dim": 1, "descending": False}
    yield "sort_long_rows", "sort", {"x": np.round(u((3, 20000), -300, 300)).astype(np.float32)}, {"dim": 1, "descending": True}
    yield "topk_f32", "topk", {"x": np.round(u((33, 22, 2223), -50, 50)).astype(np.float32)}, {"k": 8, "dim": 2, "largest": True}
    yield "topk_f32_small", "topk", {"x": np.round(u((13, 65, 149), -50, 50)).astype(np.float32)}, {"k": 8, "dim": 1, "largest": False}
    # gemm (test_gemm.py:9-17)
    yield "gemm_f64", "gemm", {"a": u((123, 457), dt=np.float64), "b": u((457, 234), dt=np.float64)}, {}
    yield "gemm_f32", "gemm", {"a": u((3, 41, 130)), "b": u((130, 77))}, {}
    # attention (test_nn.py:11-33), incl. the odd shape that takes the reference's fallback kernel
    for i, (b, h, sq, skv, d) in enumerate([(2, 4, 32, 256, 128), (3, 5, 64, 32, 64), (2, 3, 17, 33, 24)]):
        yield f"attn_{i}", "attention", {"q": u((b, h, sq, d), -1, 1), "k": u((b, h, skv, d), -1, 1), "v": u((b, h, skv, d))}, {}


def cases_r2():
    """round-2 additions: index_put_ (test_tensor.py:273-284), norm_stat (test_tensor.py:134-146) and mean_var over a
    middle dim (test_tensor.py:120-132) -> tests/golden/ref_outputs_r2.npz"""
    rng = np.random.default_rng(20261018)
    u = lambda shape, lo=-10, hi=10, dt=np.float32: rng.uniform(lo, hi, size=shape).astype(dt)
    x = u((13, 15), -10000, 10000)
    i0, i1 = np.array([0, 5, 1, 2, 12, 7]).astype(np.int64), np.array([0, 11, 1, 0, 14, 3]).astype(np.int64)
    yield "index_put_f32", "index_put", {"x": x, "i0": i0, "i1": i1, "values": u((6,), -10000, 10000)}, {}
    xi = u((9, 4, 6), -100, 100, np.int32)
    yield "index_put_i32_3d", "index_put", {"x": xi, "i0": np.array([0, 8, 3]).astype(np.int64), "i1": np.array([3, 0, 2]).astype(np.int64),
                                          "i2": np.array([5, 5, 0]).astype(np.int64), "values": u((3,), -100, 100, np.int32)}, {}
    yield "norm_stat_64", "norm_stat", {"x": u((64, 64))}, {}
    yield "norm_stat_1024x2048", "norm_stat", {"x": u((1024, 2048))}, {}
    yield "mean_var_mid", "mean_var", {"x": u((13, 325, 127))}, {"dim": 1, "take_sqrt": False}
    yield "mean_var_last_sqrt", "mean_var", {"x": u((37, 512))}, {"dim": 1, "take_sqrt": True}
