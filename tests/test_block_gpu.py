"""Transformer block (BASELINE configs[4]) forward + backward through the operator API, against the
same graph in torch-CPU float64 autograd on the (rounded) parameters.  fp32: 1e-4-class on O(1) gradients
(the composed graph chains ~10 fp32 ops); bf16: 2e-2 of each gradient's scale."""
import numpy as np
import pytest
import torch

import kfunca_b200 as kf
from kfunca_b200.block import Block
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def torch_block(params, x, H, eps=1e-5):
    B, S, E = x.shape
    D = E // H

    def norm(t, g):
        mu = t.mean(-1, keepdim=True)
        xc = t - mu
        var = (xc * xc).mean(-1, keepdim=True)
        return xc * torch.rsqrt(var + eps) * g

    xn = norm(x, params["g1"])
    qkv = xn @ params["wqkv"]
    q, k, v = qkv.split([E, E, E], -1)
    heads = lambda t: t.reshape(B, S, H, D).permute(0, 2, 1, 3)
    o = torch.nn.functional.scaled_dot_product_attention(heads(q), heads(k), heads(v), is_causal=True)
    o = o.permute(0, 2, 1, 3).reshape(B, S, E)
    x1 = x + o @ params["wo"]
    xn2 = norm(x1, params["g2"])
    h = (xn2 @ params["w1"]) * (xn2 @ params["w3"])
    y = x1 + h @ params["w2"]
    return y.mean()


def host(t):
    return t.float().numpy().astype(np.float64) if t.dtype() != kf.double else t.numpy()


@pytest.mark.parametrize("dtype,B,S,E,H,tol", [("float", 2, 64, 128, 2, 2e-4), ("bfloat16", 2, 256, 256, 2, 3e-2)])
def test_block_forward_backward_matches_torch_autograd(dtype, B, S, E, H, tol):
    kdt = getattr(kf, dtype)
    blk = Block(E, H, dtype=kdt, device=0, seed=3)
    rng = np.random.default_rng(5)
    x_np = rng.uniform(-1, 1, (B, S, E)).astype(np.float32)
    x = kf.from_numpy(x_np, 0).to(kdt)
    x.set_requires_grad(True)
    loss = blk.step(x)
    kf.synchronize()
    tp = {n: torch.tensor(host(p), dtype=torch.float64, requires_grad=True) for n, p in blk.params.items()}
    tx = torch.tensor(host(x), dtype=torch.float64, requires_grad=True)
    tl = torch_block(tp, tx, H)
    tl.backward()
    got_loss = float(host(loss).reshape(-1)[0])
    assert abs(got_loss - tl.item()) <= tol * max(1.0, abs(tl.item()))
    for n, p in blk.params.items():
        g = p.grad()
        assert g.defined(), n
        got, exp = host(g), tp[n].grad.numpy()
        scale = np.abs(exp).max()
        assert np.abs(got - exp).max() <= tol * scale, (n, np.abs(got - exp).max(), scale)
    got, exp = host(x.grad()), tx.grad.numpy()
    assert np.abs(got - exp).max() <= tol * np.abs(exp).max()


def test_leaf_grad_hook_fires_once_per_parameter_with_the_final_gradient():
    """kf.set_leaf_grad_hook (the data-parallel overlap entry point): called inside backward(), once per leaf, with the
    gradient tensor that p.grad() returns afterwards (same storage)."""
    blk = Block(128, 2, dtype=kf.float, device=0, seed=1)
    x = kf.from_numpy(np.random.default_rng(2).uniform(-1, 1, (2, 64, 128)).astype(np.float32), 0)
    seen = []
    kf.set_leaf_grad_hook(lambda leaf, grad: seen.append((leaf.data_ptr(), grad.data_ptr(), tuple(grad.sizes()))))
    try:
        blk.step(x)
    finally:
        kf.set_leaf_grad_hook(None)
    kf.synchronize()
    by_leaf = {l: (g, s) for l, g, s in seen}
    assert len(seen) == len(blk.params) == len(by_leaf)
    for n, p in blk.params.items():
        g, s = by_leaf[p.data_ptr()]
        assert g == p.grad().data_ptr() and s == tuple(p.sizes()), n
    seen.clear()
    blk.step(x)  # hook removed: nothing recorded
    assert not seen
