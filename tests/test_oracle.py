"""CPU tests that PIN the oracle (oracle/oracle.py):
  1. against outputs of the UNMODIFIED reference build run on a B200 (tests/golden/ref_outputs.npz, produced by
     oracle/make_golden_from_ref.py from the seeded cases in oracle/golden_cases.py);
  2. against the NumPy / torch-CPU calls the reference's own tests use as expectations."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from oracle.golden_cases import cases

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_outputs.npz"))
CASES = {name: (kind, inp, prm) for name, kind, inp, prm in cases()}


def test_golden_file_covers_the_cases():
    have = {k.split(".")[0] for k in GOLD.files}
    # the reference rejects NumPy's 'l' int64 (register.cpp:23,30), so that case has no reference output
    assert set(CASES) - have == {"bin_i64_div"}


@pytest.mark.parametrize("name", [n for n, (k, _, _) in CASES.items() if k == "binary" and n != "bin_i64_div"])
def test_binary_matches_reference_bit_exact(name):
    _, inp, prm = CASES[name]
    exp = GOLD[f"{name}.out"]
    got = O.binary(prm["op"], inp["a"], inp["b"])
    assert got.dtype == exp.dtype
    np.testing.assert_array_equal(got.view(np.uint8), exp.view(np.uint8))


@pytest.mark.parametrize("name", [n for n, (k, _, _) in CASES.items() if k == "reduce"])
def test_reduce_matches_reference(name):
    _, inp, prm = CASES[name]
    exp = GOLD[f"{name}.out"]
    x = inp["x"]
    if x.dtype.kind == "i":
        got = O.reduce_int(prm["op"], x, prm["dim"])
        assert got.dtype == exp.dtype
        np.testing.assert_array_equal(got, exp)  # incl. the integer-mean-is-zero quirk
    else:
        exact = O.reduce_exact(prm["op"], x, prm["dim"])
        mass = np.abs(x.astype(np.float64)).sum(axis=prm["dim"], keepdims=True) / (x.shape[prm["dim"]] if prm["op"] == "mean" else 1)
        assert exp.shape == exact.shape
        assert O.l1_tolerance_ok(exp, exact, mass, 1e-5)


def test_permute_matches_reference():
    _, inp, prm = CASES["permute_f64"]
    np.testing.assert_array_equal(np.ascontiguousarray(inp["x"].transpose(prm["dims"])), GOLD["permute_f64.out"])


@pytest.mark.parametrize("name", [n for n, (k, _, _) in CASES.items() if k == "sort"])
def test_sort_matches_reference_bit_exact(name):
    _, inp, prm = CASES[name]
    v, i = O.sort(inp["x"], prm["dim"], prm["descending"])
    np.testing.assert_array_equal(v, GOLD[f"{name}.values"])
    np.testing.assert_array_equal(i, GOLD[f"{name}.indices"])  # tie order: ascending index


@pytest.mark.parametrize("name", [n for n, (k, _, _) in CASES.items() if k == "topk"])
def test_topk_matches_reference_bit_exact(name):
    _, inp, prm = CASES[name]
    v, i = O.topk(inp["x"], prm["k"], prm["dim"], prm["largest"])
    np.testing.assert_array_equal(v, GOLD[f"{name}.values"])
    np.testing.assert_array_equal(i, GOLD[f"{name}.indices"])


@pytest.mark.parametrize("name", ["gemm_f64", "gemm_f32"])
def test_gemm_matches_reference(name):
    _, inp, _ = CASES[name]
    exact = O.gemm(inp["a"], inp["b"])
    mass = np.abs(inp["a"].astype(np.float64)) @ np.abs(inp["b"].astype(np.float64))
    exp = GOLD[f"{name}.out"]
    assert exp.shape == exact.shape
    assert O.l1_tolerance_ok(exp, exact, mass, 1e-5 if name.endswith("f32") else 1e-13)


@pytest.mark.parametrize("name", ["attn_0", "attn_1", "attn_2"])
def test_attention_matches_reference(name):
    _, inp, _ = CASES[name]
    exact = O.causal_attention(inp["q"], inp["k"], inp["v"])
    np.testing.assert_allclose(GOLD[f"{name}.out"], exact, rtol=1e-4, atol=1e-4)


# ---- the reference's own expectations: NumPy / torch-CPU (test/test_tensor.py, test_gemm.py, test_nn.py)
def test_oracle_vs_numpy_elementwise_and_reduce():
    rng = np.random.default_rng(7)
    a, b = rng.uniform(-10, 10, (12, 11, 331)).astype(np.float32), rng.uniform(1, 10, (12, 11, 331)).astype(np.float32)
    for op, fn in (("+", np.add), ("-", np.subtract), ("*", np.multiply), ("/", np.divide)):
        np.testing.assert_array_equal(O.binary(op, a, b), fn(a, b))
    i = rng.uniform(-10, 10, (12, 11, 331)).astype(np.int32)
    np.testing.assert_array_equal(O.binary("+", i, b), (i + b).astype(np.float32))  # int32 + fp32 -> fp32 (test_tensor.py:22-27)
    for dim in (0, 1, 2):
        np.testing.assert_allclose(O.reduce_exact("sum", a, dim), np.sum(a, axis=dim, keepdims=True), rtol=1e-4, atol=1e-2)
        np.testing.assert_allclose(O.reduce_exact("mean", a, dim), np.mean(a, axis=dim, keepdims=True), rtol=1e-4, atol=1e-4)


def test_oracle_sort_vs_torch_stable_sort():
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(8)
    for dt in (np.float32, np.float64, np.int32):
        x = (np.round(rng.uniform(-30, 30, (7, 33, 129))) + 0.0).astype(dt)  # many ties; +0.0 folds -0.0 (radix keys order -0 < +0, torch does not)
        for dim in (0, 1, 2):
            for desc in (False, True):
                tv, ti = torch.sort(torch.from_numpy(x), dim=dim, descending=desc, stable=True)
                v, i = O.sort(x, dim, desc)
                np.testing.assert_array_equal(v, tv.numpy())
                np.testing.assert_array_equal(i, ti.numpy())
    x = rng.uniform(-1000, 1000, (4, 102400)).astype(np.float32)
    v, i = O.sort(x, 1, False)
    np.testing.assert_array_equal(i, np.argsort(x, axis=1, kind="stable"))
    tv, _ = torch.topk(torch.from_numpy(x), 8, dim=1, largest=True)
    np.testing.assert_array_equal(O.topk(x, 8, 1, True)[0], tv.numpy())  # values only


def test_oracle_attention_vs_torch_sdpa():
    torch = pytest.importorskip("torch")
    import torch.nn.functional as F

    rng = np.random.default_rng(9)
    for (b, h, sq, skv, d) in [(2, 4, 32, 256, 128), (3, 5, 64, 32, 64), (5, 16, 65, 33, 123)]:  # test_nn.py:13-17
        q, k, v = (rng.uniform(-10, 10, s).astype(np.float32) for s in ((b, h, sq, d), (b, h, skv, d), (b, h, skv, d)))
        ref = F.scaled_dot_product_attention(torch.from_numpy(q), torch.from_numpy(k), torch.from_numpy(v), is_causal=True).numpy()
        np.testing.assert_allclose(O.causal_attention(q, k, v), ref, rtol=1e-3, atol=1e-3)


def test_oracle_attention_bwd_vs_torch_autograd():
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(10)
    q, k, v, do = (rng.uniform(-1, 1, (2, 3, 24, 16)) for _ in range(4))
    tq, tk, tv = (torch.tensor(t, requires_grad=True) for t in (q, k, v))
    out = torch.nn.functional.scaled_dot_product_attention(tq, tk, tv, is_causal=True)
    out.backward(torch.tensor(do))
    dq, dk, dv = O.causal_attention_bwd(q, k, v, do)
    np.testing.assert_allclose(dq, tq.grad.numpy(), rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(dk, tk.grad.numpy(), rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(dv, tv.grad.numpy(), rtol=1e-9, atol=1e-11)


def test_oracle_casts_16bit():
    rng = np.random.default_rng(11)
    x = rng.uniform(-10, 10, (1000,))
    np.testing.assert_array_equal(O.cast(x, O.HALF), x.astype(np.float32).astype(np.float16))
    np.testing.assert_array_equal(O.cast(x, O.BFLOAT16), x.astype(np.float32).astype(O.bfloat16))
    assert float(O.scalar_through(O.HALF, 2.5)) == 2.5 and int(O.scalar_through(O.INT, 2.9)) == 2


def test_oracle_layer_norm_and_mean_var_vs_torch():
    """the oracle's layer_norm / layer_norm_bwd / mean_var against torch-CPU float64 (the way the reference's own tests pin
    their expectations: recomputed from a library at test time, test/test_tensor.py:120-146)"""
    import torch

    rng = np.random.default_rng(21)
    x = rng.uniform(-3, 3, (5, 7, 96))
    gain = rng.uniform(0.5, 1.5, (96,))
    dy = rng.uniform(-1, 1, x.shape)
    tx = torch.tensor(x, requires_grad=True)
    tg = torch.tensor(gain, requires_grad=True)
    ty = torch.nn.functional.layer_norm(tx, (96,), weight=tg, eps=1e-5)
    ty.backward(torch.tensor(dy))
    np.testing.assert_allclose(O.layer_norm(x, gain, 1e-5), ty.detach().numpy(), rtol=1e-12, atol=1e-12)
    dx, dgain = O.layer_norm_bwd(x, gain, dy, 1e-5)
    np.testing.assert_allclose(dx, tx.grad.numpy(), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(dgain, tg.grad.numpy(), rtol=1e-10, atol=1e-12)
    m, v = O.mean_var(x, 1)
    np.testing.assert_allclose(m, torch.tensor(x).mean(1, keepdim=True).numpy(), rtol=1e-12)
    np.testing.assert_allclose(v, torch.tensor(x).var(1, unbiased=True, keepdim=True).numpy(), rtol=1e-12)
    _, sd = O.mean_var(x, 2, True)
    np.testing.assert_allclose(sd, torch.tensor(x).std(2, unbiased=True, keepdim=True).numpy(), rtol=1e-12)


def test_oracle_round2_functions_vs_torch():
    """rms_norm (+ backward), norm_stat, index_put, embedding (+ backward) against torch-CPU float64 — the same way the reference's
    tests pin their expectations (test/test_tensor.py:134-146, 273-284)."""
    import torch

    rng = np.random.default_rng(33)
    x, gain, dy = rng.uniform(-3, 3, (6, 5, 64)), rng.uniform(0.5, 1.5, (64,)), rng.uniform(-1, 1, (6, 5, 64))
    tx, tg = torch.tensor(x, requires_grad=True), torch.tensor(gain, requires_grad=True)
    ty = tx * torch.rsqrt((tx * tx).mean(-1, keepdim=True) + 1e-5) * tg
    ty.backward(torch.tensor(dy))
    np.testing.assert_allclose(O.rms_norm(x, gain, 1e-5), ty.detach().numpy(), rtol=1e-12, atol=1e-12)
    dx, dg = O.rms_norm_bwd(x, gain, dy, 1e-5)
    np.testing.assert_allclose(dx, tx.grad.numpy(), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(dg, tg.grad.numpy(), rtol=1e-10, atol=1e-12)
    a = rng.uniform(-10, 10, (300, 17))
    m, inv = O.norm_stat(a)
    np.testing.assert_allclose(m, a.mean(0, keepdims=True), rtol=1e-12)
    np.testing.assert_allclose(inv, 1.0 / np.sqrt(((a - a.mean(0)) ** 2).sum(0, keepdims=True) / 300), rtol=1e-9)  # test_tensor.py:141-143
    base = rng.uniform(-1e4, 1e4, (13, 15)).astype(np.float32)
    i0, i1 = np.array([0, 5, 1, 2]), np.array([0, 11, 1, 0])
    vals = rng.uniform(-1e4, 1e4, 4).astype(np.float32)
    want = torch.from_numpy(base.copy())
    want.index_put_([torch.from_numpy(i0), torch.from_numpy(i1)], torch.from_numpy(vals))
    np.testing.assert_array_equal(O.index_put(base, [i0, i1], vals), want.numpy())
    w = rng.uniform(-1, 1, (20, 8))
    idx = rng.integers(0, 20, (3, 7))
    tw = torch.tensor(w, requires_grad=True)
    go = rng.uniform(-1, 1, (3, 7, 8))
    te = torch.nn.functional.embedding(torch.from_numpy(idx), tw)
    te.backward(torch.tensor(go))
    np.testing.assert_array_equal(O.embedding(w, idx), te.detach().numpy())
    np.testing.assert_allclose(O.embedding_bwd(idx, go, 20), tw.grad.numpy(), rtol=1e-12, atol=1e-12)


def test_oracle_counter_uniform_properties():
    a = O.counter_uniform(0, 100000, 5, -2.0, 3.0)
    assert a.dtype == np.float32 and a.min() >= -2.0 and a.max() < 3.0
    assert abs(float(a.mean()) - 0.5) < 0.02 and np.unique(a).size > 99000
    np.testing.assert_array_equal(O.counter_uniform(1000, 50, 5, -2.0, 3.0), a[1000:1050])      # random access = sequential
    assert not np.array_equal(O.counter_uniform(0, 100, 6, -2.0, 3.0), a[:100])                   # the seed matters
    big = O.counter_uniform((1 << 31) + 7, 4, 5, -1.0, 1.0)                                        # indices beyond 2^31
    assert np.all(np.isfinite(big))


def test_oracle_against_reference_build_outputs_round2():
    """replay of tests/golden/ref_outputs_r2.npz: outputs of the UNMODIFIED reference build (oracle/_ref) on a B200 for
    index_put_, norm_stat and mean_var (oracle/make_golden_from_ref.py r2)."""
    import os

    from oracle.golden_cases import cases_r2

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_outputs_r2.npz")
    if not os.path.exists(path):
        pytest.skip("ref_outputs_r2.npz not generated yet")
    gold = np.load(path)
    seen = 0
    for name, kind, inp, prm in cases_r2():
        if kind == "index_put" and f"{name}.out" in gold:
            idx = [inp[k] for k in ("i0", "i1", "i2") if k in inp]
            np.testing.assert_array_equal(O.index_put(inp["x"], idx, inp["values"]), gold[f"{name}.out"])
            seen += 1
        elif kind == "norm_stat" and f"{name}.mean" in gold:
            m, inv = O.norm_stat(inp["x"])
            np.testing.assert_allclose(gold[f"{name}.mean"], m, rtol=1e-4, atol=1e-5)
            np.testing.assert_allclose(gold[f"{name}.invstd"], inv, rtol=1e-4)
            seen += 1
        elif kind == "mean_var" and f"{name}.mean" in gold:
            m, v = O.mean_var(inp["x"], prm["dim"], prm["take_sqrt"])
            np.testing.assert_allclose(gold[f"{name}.mean"], m, rtol=1e-4, atol=1e-5)
            np.testing.assert_allclose(gold[f"{name}.var"], v, rtol=1e-4)
            seen += 1
    assert seen >= 1
