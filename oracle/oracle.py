"""TEST INFRASTRUCTURE ONLY — CPU restatement (NumPy) of the reference's operator semantics.

Nothing under kfunca_b200/ imports this file; only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs do.  Every function cites the reference lines it restates
(paths relative to the xytpai/kfunca checkout).

Pinning: the reference stores no golden vectors; its tests recompute expectations with
NumPy / torch-CPU.  This oracle is pinned two ways:
  1. tests/test_oracle.py checks it against the same NumPy / torch-CPU calls the reference's tests use
     (np.add/.., np.sum/mean(keepdims), torch.sort(stable=True), np.matmul, F.scaled_dot_product_attention);
  2. tests/golden/ref_*.npz hold outputs of the UNMODIFIED reference build (oracle/_ref, see oracle/Makefile)
     produced on a B200 by oracle/make_golden_from_ref.py; tests/test_oracle.py replays them.
"""
from __future__ import annotations

import numpy as np

try:  # bf16 host dtype (casts only)
    import ml_dtypes

    bfloat16 = np.dtype(ml_dtypes.bfloat16)
except Exception:  # pragma: no cover
    ml_dtypes = None
    bfloat16 = None

# ref: src/core/include/scalar_type.h:9-27 — enum order is the promotion order
BOOL, BYTE, CHAR, SHORT, INT, LONG, HALF, BFLOAT16, FLOAT, DOUBLE = range(10)
NAMES = ["bool", "byte", "char", "short", "int", "long", "half", "bfloat16", "float", "double"]


def np_dtype(code: int) -> np.dtype:
    table = {
        BOOL: np.dtype(np.bool_), BYTE: np.dtype(np.uint8), CHAR: np.dtype(np.int8), SHORT: np.dtype(np.int16),
        INT: np.dtype(np.int32), LONG: np.dtype(np.int64), HALF: np.dtype(np.float16), BFLOAT16: bfloat16,
        FLOAT: np.dtype(np.float32), DOUBLE: np.dtype(np.float64),
    }
    return table[code]


def code_of(dt) -> int:
    dt = np.dtype(dt)
    for c in range(10):
        if np_dtype(c) is not None and np_dtype(c) == dt:
            return c
    raise TypeError(f"unsupported dtype {dt}")


def is_floating(c: int) -> bool:
    return c in (HALF, BFLOAT16, FLOAT, DOUBLE)


def is_unsigned_class(c: int) -> bool:
    return c in (BOOL, BYTE)


def promote(a: int, b: int) -> int:
    """ref: update_common_dtype, src/core/tensor_iterator.cpp:32-44"""
    if is_floating(a) and is_floating(b):
        return max(a, b)
    if is_floating(a) or is_floating(b):
        return a if is_floating(a) else b
    if is_unsigned_class(a) and is_unsigned_class(b):
        return max(a, b)
    if is_unsigned_class(a) or is_unsigned_class(b):
        return b if is_unsigned_class(a) else a
    return max(a, b)


def acc_dtype(c: int) -> np.dtype:
    """ref: src/core/include/accumulate_type.h:17-27"""
    if c in (HALF, BFLOAT16, FLOAT):
        return np.dtype(np.float32)
    if c == DOUBLE:
        return np.dtype(np.float64)
    if c == BOOL:
        return np.dtype(np.bool_)
    return np.dtype(np.int64)


def cast(x: np.ndarray, code: int) -> np.ndarray:
    """static_cast<dst>(src) as the device does it (ref: tensor_memory_access.h:13-37): 16-bit floats go
    through fp32 with round-to-nearest-even; integer narrowing wraps."""
    dst = np_dtype(code)
    src = x.dtype
    if src == dst:
        return x.copy()
    if code in (HALF, BFLOAT16):
        if src.kind in "iub":
            x = x.astype(np.float32)
        elif src == np.float64:
            x = x.astype(np.float32)
        elif src != np.float32:
            x = x.astype(np.float32)
        return x.astype(dst)
    if src in (np.dtype(np.float16), bfloat16) and code != FLOAT:
        x = x.astype(np.float32)
    if code == BOOL:
        return x != 0
    with np.errstate(all="ignore"):
        if dst.kind in "iu" and x.dtype.kind == "f":
            return np.trunc(x).astype(np.int64).astype(dst)  # in-range values only; out-of-range is UB in the reference
        return x.astype(dst)


def _trunc_div_int(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    q = np.abs(a) // np.abs(b)
    return np.where((a < 0) != (b < 0), -q, q)


def binary(op: str, x: np.ndarray, y: np.ndarray, out_code: int | None = None) -> np.ndarray:
    """out = x op y.  ref: add/sub/mul/div_kernel, src/device/binary_ops_kernel.cu:34-60 — operands are cast
    to acc_type(common dtype), combined, and cast to the output dtype (common dtype, or self's for in-place)."""
    cx, cy = code_of(x.dtype), code_of(y.dtype)
    common = promote(cx, cy)
    acc = acc_dtype(common)
    xa, ya = _to_acc(x, acc), _to_acc(y, acc)
    with np.errstate(all="ignore"):
        if acc == np.bool_:
            r = {"+": xa | ya, "-": xa ^ ya, "*": xa & ya, "/": xa & ya}[op]
        elif acc == np.int64:
            r = {"+": lambda: xa + ya, "-": lambda: xa - ya, "*": lambda: xa * ya, "/": lambda: _trunc_div_int(xa, ya)}[op]()
        else:
            r = {"+": lambda: xa + ya, "-": lambda: xa - ya, "*": lambda: xa * ya, "/": lambda: xa / ya}[op]()
            r = r.astype(acc)
    return cast(np.asarray(r), common if out_code is None else out_code)


def _to_acc(x: np.ndarray, acc: np.dtype) -> np.ndarray:
    if x.dtype in (np.dtype(np.float16), bfloat16):
        x = x.astype(np.float32)
    if acc == np.bool_:
        return x != 0
    if acc == np.int64 and x.dtype.kind == "f":
        return np.trunc(x).astype(np.int64)
    return x.astype(acc)


def scalar_through(code: int, v: float) -> float:
    """`t op 2.5`: the reference fills a tensor of t's dtype with the scalar first (src/register.cpp:172-206,
    nullary_ops_kernel.cu:19-25): double -> acc_t -> dtype."""
    acc = acc_dtype(code)
    a = np.array(v, dtype=np.float64)
    a = (a != 0) if acc == np.bool_ else (np.trunc(a).astype(np.int64) if acc == np.int64 else a.astype(acc))
    return cast(np.asarray(a), code)


def binary_scalar(op: str, x: np.ndarray, v: float) -> np.ndarray:
    s = scalar_through(code_of(x.dtype), v)
    return binary(op, x, np.broadcast_to(s, x.shape))


def reduce_exact(op: str, x: np.ndarray, dim: int) -> np.ndarray:
    """float64 ground truth of sum/mean over `dim`, keepdim (ref: reduce_ops.cpp:8-20 keepdim via
    tensor_iterator.cpp:60-76).  For floating inputs only; parity is judged with an L1-mass tolerance."""
    xf = x.astype(np.float32).astype(np.float64) if x.dtype in (np.dtype(np.float16), bfloat16) else x.astype(np.float64)
    return getattr(np, op)(xf, axis=dim, keepdims=True)


def reduce_int(op: str, x: np.ndarray, dim: int) -> np.ndarray:
    """Integer / bool sum & mean, bit-exact restatement.  ref: SumFunctor<scalar_t> and MeanOps<scalar_t,
    acc_t=scalar_t> (src/device/reduce_ops_kernel.cu:6-59): accumulation wraps in the input dtype; the mean
    factor is `static_cast<scalar_t>(n_out) / numel` in integer arithmetic (0 unless nothing is reduced)."""
    code = code_of(x.dtype)
    if code == BOOL:
        s = np.any(x, axis=dim, keepdims=True)
        if op == "sum":
            return s
        n_out, numel = x.size // max(x.shape[dim], 1), x.size
        factor = (1 // numel) != 0 if n_out != 0 else False
        return s & factor
    s = np.sum(x.astype(np.int64), axis=dim, keepdims=True)
    if op == "sum":
        return s.astype(x.dtype)  # wrap
    n_out, numel = s.size, x.size
    t = np.array(n_out, dtype=np.int64).astype(x.dtype).astype(np.int64)
    factor = np.array(_trunc_div_int(t, np.array(numel, dtype=np.int64)), dtype=np.int64).astype(x.dtype).astype(np.int64)
    return (s * factor).astype(x.dtype)


def sort_key(x: np.ndarray) -> np.ndarray:
    """Order-preserving unsigned key.  ref: KeyTraits<T>::convert, src/device/utils/sorting_common.h:39-238
    (floats: flip all bits of negatives, sign bit of non-negatives => -NaN < -inf < .. < -0 < +0 < .. < +inf < +NaN)."""
    dt = x.dtype
    if dt.kind == "f" or dt == bfloat16:
        bits = {2: np.uint16, 4: np.uint32, 8: np.uint64}[dt.itemsize]
        u = x.view(bits)
        sign = np.array(1, dtype=bits) << np.array(dt.itemsize * 8 - 1, dtype=bits)
        return np.where(u & sign, ~u, u | sign).astype(bits)
    if dt.kind == "i":
        bits = {1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}[dt.itemsize]
        sign = np.array(1, dtype=bits) << np.array(dt.itemsize * 8 - 1, dtype=bits)
        return x.view(bits) ^ sign
    return x


def sort(x: np.ndarray, dim: int, descending: bool):
    """Stable sort along dim: (values, int64 indices); descending keeps ascending index among ties.
    ref: sort_stable_kernel, src/device/sort_ops_kernel.cu:553-615; descending = reversed bucket order in a
    stable LSD radix sort (sorting_radix_sort.h:327-330), i.e. a stable sort on the complemented key."""
    key = sort_key(np.ascontiguousarray(x))
    if descending:
        key = ~key
    idx = np.argsort(key, axis=dim, kind="stable").astype(np.int64)
    return np.take_along_axis(x, idx, axis=dim), idx


def topk(x: np.ndarray, k: int, dim: int, largest: bool):
    """ref: topk_with_sort, src/device/sort_ops_kernel.cu:617-632 — full stable sort, first k."""
    v, i = sort(x, dim, largest)
    sl = [slice(None)] * x.ndim
    sl[dim] = slice(0, k)
    return np.ascontiguousarray(v[tuple(sl)]), np.ascontiguousarray(i[tuple(sl)])


def gemm(a: np.ndarray, b: np.ndarray, alpha: float = 1.0) -> np.ndarray:
    """float64 ground truth of alpha * a[..,K] @ b[K,N] (ref: gemm_kernel, src/device/gemm_kernel.cu:8-38)."""
    af = a.astype(np.float32).astype(np.float64) if a.dtype.itemsize == 2 else a.astype(np.float64)
    bf = b.astype(np.float32).astype(np.float64) if b.dtype.itemsize == 2 else b.astype(np.float64)
    return alpha * (af @ bf)


def causal_attention(q: np.ndarray, k: np.ndarray, v: np.ndarray, return_lse: bool = False):
    """float64 ground truth.  ref: CausalAttentionRefForwardFN, src/device/utils/causal_attention_ref.h:25-64:
    s = q.k^T / sqrt(D); keep s[m, n] where m >= n (top-left aligned) else -inf; softmax over n; p @ v."""
    def f64(t):
        return t.astype(np.float32).astype(np.float64) if t.dtype.itemsize == 2 else t.astype(np.float64)

    q, k, v = f64(q), f64(k), f64(v)
    sq, skv, d = q.shape[-2], k.shape[-2], q.shape[-1]
    s = (q @ np.swapaxes(k, -1, -2)) / np.sqrt(float(d))  # matmul = BLAS: the S = 4096 heads of the C3 test finish in seconds
    mask = np.arange(sq)[:, None] >= np.arange(skv)[None, :]
    s = np.where(mask, s, -np.inf)
    m = s.max(axis=-1, keepdims=True)
    e = np.exp(s - m)
    l = e.sum(axis=-1, keepdims=True)
    out = (e / l) @ v
    if return_lse:
        return out, (m + np.log(l))[..., 0]
    return out


def causal_attention_bwd(q, k, v, dout):
    """float64 analytic gradients of causal_attention (the reference has no backward)."""
    def f64(t):
        return t.astype(np.float32).astype(np.float64) if t.dtype.itemsize == 2 else t.astype(np.float64)

    q, k, v, dout = f64(q), f64(k), f64(v), f64(dout)
    sq, skv, d = q.shape[-2], k.shape[-2], q.shape[-1]
    scale = 1.0 / np.sqrt(float(d))
    s = (q @ np.swapaxes(k, -1, -2)) * scale
    mask = np.arange(sq)[:, None] >= np.arange(skv)[None, :]
    s = np.where(mask, s, -np.inf)
    p = np.exp(s - s.max(axis=-1, keepdims=True))
    p = p / p.sum(axis=-1, keepdims=True)
    dv = np.swapaxes(p, -1, -2) @ dout
    dp = dout @ np.swapaxes(v, -1, -2)
    delta = (p * dp).sum(axis=-1, keepdims=True)
    ds = p * (dp - delta) * scale
    dq = ds @ k
    dk = np.swapaxes(ds, -1, -2) @ q
    return dq, dk, dv


def l1_tolerance_ok(got: np.ndarray, exact: np.ndarray, l1_mass: np.ndarray, rel: float) -> bool:
    """|got - exact| <= rel * sum|terms| — the well-conditioned form of "within rel" (SURVEY §8d)."""
    return bool(np.all(np.abs(got.astype(np.float64) - exact) <= rel * l1_mass + 1e-30))


def mean_var(x: np.ndarray, dim: int, take_sqrt: bool = False):
    """mean and UNBIASED variance (correction 1) along `dim`, keepdim, in float64 — the reference's Welford result
    (src/device/reduce_ops_kernel.cu:61-153, `MeanVarOps` with correction 1; src/core/reduce_ops.cpp:22-28)."""
    x64 = np.asarray(x).astype(np.float64)
    m = x64.mean(axis=dim, keepdims=True)
    v = x64.var(axis=dim, ddof=1, keepdims=True)
    return m, (np.sqrt(v) if take_sqrt else v)


def layer_norm(x: np.ndarray, gain: np.ndarray, eps: float = 1e-5) -> np.ndarray:
    """y = (x - mean) / sqrt(var_biased + eps) * gain over the last dim, float64.  The statistics are those of the
    reference's norm_stat (biased variance, rsqrt(var + eps); src/device/norm_ops_kernel.cu:6-61,
    src/device/utils/welford_norm.h:25-355); the fused op itself is the reference's next planned op (README.md:28)."""
    x64 = np.asarray(x).astype(np.float64)
    g64 = np.asarray(gain).astype(np.float64).reshape(-1)
    m = x64.mean(axis=-1, keepdims=True)
    v = ((x64 - m) ** 2).mean(axis=-1, keepdims=True)
    return (x64 - m) / np.sqrt(v + eps) * g64


def layer_norm_bwd(x: np.ndarray, gain: np.ndarray, dy: np.ndarray, eps: float = 1e-5):
    """gradients of layer_norm in float64: dx (same shape as x) and dgain ([E])."""
    x64 = np.asarray(x).astype(np.float64)
    g64 = np.asarray(gain).astype(np.float64).reshape(-1)
    dy64 = np.asarray(dy).astype(np.float64)
    m = x64.mean(axis=-1, keepdims=True)
    v = ((x64 - m) ** 2).mean(axis=-1, keepdims=True)
    rstd = 1.0 / np.sqrt(v + eps)
    xh = (x64 - m) * rstd
    gg = dy64 * g64
    dx = rstd * (gg - gg.mean(axis=-1, keepdims=True) - xh * (gg * xh).mean(axis=-1, keepdims=True))
    dgain = (dy64 * xh).reshape(-1, x64.shape[-1]).sum(axis=0)
    return dx, dgain


def rms_norm(x: np.ndarray, gain: np.ndarray, eps: float = 1e-5) -> np.ndarray:
    """y = x / sqrt(mean(x^2) + eps) * gain over the last dim, float64 — the op the reference's README lists as its next one
    (README.md:28 `rms_norm`); statistics in the style of its mean_var / norm_stat kernels (reduce_ops_kernel.cu:61-153)."""
    x64 = np.asarray(x).astype(np.float32).astype(np.float64) if np.asarray(x).dtype.itemsize == 2 else np.asarray(x).astype(np.float64)
    g64 = np.asarray(gain).astype(np.float32).astype(np.float64).reshape(-1) if np.asarray(gain).dtype.itemsize == 2 else np.asarray(gain).astype(np.float64).reshape(-1)
    ms = (x64 * x64).mean(axis=-1, keepdims=True)
    return x64 / np.sqrt(ms + eps) * g64


def rms_norm_bwd(x: np.ndarray, gain: np.ndarray, dy: np.ndarray, eps: float = 1e-5):
    """gradients of rms_norm in float64: dx and dgain ([E])."""
    x64, g64, dy64 = (np.asarray(t).astype(np.float64) for t in (x, gain, dy))
    g64 = g64.reshape(-1)
    rstd = 1.0 / np.sqrt((x64 * x64).mean(axis=-1, keepdims=True) + eps)
    xh = x64 * rstd
    gg = dy64 * g64
    dx = rstd * (gg - xh * (gg * xh).mean(axis=-1, keepdims=True))
    return dx, (dy64 * xh).reshape(-1, x64.shape[-1]).sum(axis=0)


def norm_stat(x: np.ndarray, eps: float = 1e-12):
    """(mean, invstd) over dim 0 of a 2-D array, keepdim, float64: biased variance, invstd = 1 / sqrt(var + eps).
    ref: norm_stat_kernel (src/device/norm_ops_kernel.cu:6-61) + WelfordNormPFKernel (src/device/utils/welford_norm.h:25-355);
    the reference's own test recomputes it as 1 / sqrt(sum((x - mean)^2) / R) (test/test_tensor.py:134-146)."""
    x64 = np.asarray(x).astype(np.float64)
    m = x64.mean(axis=0, keepdims=True)
    v = ((x64 - m) ** 2).mean(axis=0, keepdims=True)
    return m, 1.0 / np.sqrt(v + eps)


def index_put(x: np.ndarray, indices, values: np.ndarray) -> np.ndarray:
    """self[idx0[i], idx1[i], ...] = values[i] on a copy; negative indices wrap.  ref: index_put_kernel /
    IndexElementwiseKernel (src/device/utils/tensor_index.h:19-143, src/core/index_ops.cpp:6-38).  Duplicate index tuples are
    unspecified in the reference (last writer wins on the GPU); tests use distinct tuples."""
    out = np.array(x, copy=True)
    idx = tuple(np.where(np.asarray(ix) < 0, np.asarray(ix) + out.shape[d], np.asarray(ix)) for d, ix in enumerate(indices))
    out[idx] = np.asarray(values).reshape(-1)
    return out


def embedding(weight: np.ndarray, idx: np.ndarray) -> np.ndarray:
    """out[..., :] = weight[idx[...], :] (the gather direction of the reference's index kernels, tensor_index.h:19-143;
    README.md:30 `embedding`).  Negative ids wrap."""
    idx = np.asarray(idx)
    return weight[np.where(idx < 0, idx + weight.shape[0], idx)]


def embedding_bwd(idx: np.ndarray, grad: np.ndarray, V: int) -> np.ndarray:
    """dW[v] = sum of grad rows whose id is v, accumulated in ascending position order in float32 (16-bit and fp32 inputs) or
    float64 — the same order and precision as the deterministic device kernel, so fp32 results can be compared bit for bit."""
    idx = np.asarray(idx).reshape(-1)
    g = np.asarray(grad).reshape(idx.size, -1)
    acc_t = np.float64 if g.dtype == np.float64 else np.float32
    dw = np.zeros((V, g.shape[1]), dtype=acc_t)
    for p in range(idx.size):  # small cases only
        dw[idx[p] if idx[p] >= 0 else idx[p] + V] += g[p].astype(acc_t)
    return dw.astype(g.dtype)


def counter_uniform(start: int, count: int, seed: int, lo: float, hi: float) -> np.ndarray:
    """Host twin of kf_random_uniform_ (kernels/misc.cu fill_random_kernel): element i = lo + (hi - lo) * u_i in float32 with a
    separately rounded multiply and add, u_i = top 24 bits of splitmix64(i + seed * golden) * 2^-24.  Returns elements
    [start, start + count) as float32.  Test infrastructure: lets the BASELINE-size inputs be generated on the device and any
    sampled row be re-created on the host."""
    with np.errstate(over="ignore"):
        z = np.arange(start, start + count, dtype=np.uint64) + np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    u = (z >> np.uint64(40)).astype(np.float32) * np.float32(5.9604644775390625e-8)
    span = np.float32(np.float32(hi) - np.float32(lo))
    return (np.float32(lo) + (span * u).astype(np.float32)).astype(np.float32)
