"""Transformer block (BASELINE.json configs[4]) written ONLY with the kfunca operator API
(gemm, causal_attention, + - * /, mean, permute/view/split/contiguous) so it is expressible through the
reference's register.cpp names; forward + backward through the library's autograd.

  (the two layer norms default to the fused kf.layer_norm kernel pair; fused_norm=False composes them from mean / - / * / rsqrt)
  x:[B,S,E] -> LN1 -> qkv = gemm(xn, Wqkv[E,3E]) -> split -> [B,H,S,D] -> causal_attention -> [B,S,E]
    -> x1 = x + gemm(o, Wo) -> LN2 -> h = gemm(xn2, W1) * gemm(xn2, W3) -> y = x1 + gemm(h, W2[4E,E]); loss = mean(y)
"""
from __future__ import annotations

import numpy as np

import kfunca_b200 as kf


def _norm(x, gain, eps=1e-5):
    """layer norm composed from the reference's own operator set (mean, - , *, rsqrt): 8 launches forward"""
    mu = x.mean(-1)
    xc = x - mu
    var = (xc * xc).mean(-1)
    return xc * kf.rsqrt(var + eps) * gain


def _norm_fused(x, gain, eps=1e-5):
    """the same function as ONE kernel forward and one backward (kf.layer_norm)"""
    return kf.layer_norm(x, gain, eps)


class Block:
    def __init__(self, E: int, H: int, dtype=None, device: int = 0, seed: int = 0, ffn_mult: int = 4, fused_norm: bool = True,
                 fused: bool = True):
        """fused=True (default) uses the fused linear ops of SURVEY §8f rank 4 — `gemm_residual` (x + gemm in the GEMM epilogue),
        `gemm_glu` (both GLU products and their multiply in one dual-B kernel) and `qkv_attention` (attention reading q, k, v in
        place from the packed qkv GEMM output and writing [B, S, E] directly: no head transposes in either direction);
        fused=False is the same block written only with the reference's own operator names (register.cpp)."""
        assert E % H == 0
        self.fused = fused
        self.norm = _norm_fused if fused_norm else _norm
        self.E, self.H, self.D, self.F = E, H, E // H, ffn_mult * E
        self.dtype = dtype or kf.bfloat16
        rng = np.random.default_rng(seed)

        def w(shape, scale):
            t = kf.from_numpy((rng.uniform(-1, 1, shape) * scale).astype(np.float32), device).to(self.dtype)
            t.set_requires_grad(True)
            return t

        s = 1.0 / np.sqrt(E)
        self.params = {
            "g1": w((1, 1, E), 0.0), "g2": w((1, 1, E), 0.0),
            "wqkv": w((E, 3 * E), s), "wo": w((E, E), s), "w1": w((E, self.F), s), "w3": w((E, self.F), s),
            "w2": w((self.F, E), 1.0 / np.sqrt(self.F)),
        }
        for g in ("g1", "g2"):  # gains start at one
            self.params[g] += 1.0

    def forward(self, x):
        p = self.params
        B, S, E = x.sizes()
        H, D = self.H, self.D
        xn = self.norm(x, p["g1"])
        qkv = kf.qkv_linear(xn, p["wqkv"], None) if self.fused else kf.gemm(xn, p["wqkv"], 1.0, 0.0)  # README.md:32 (no bias in this block)
        if self.fused and hasattr(kf, "qkv_attention"):
            o = kf.qkv_attention(qkv, H)  # [B, S, 3E] -> [B, S, E]
        else:
            q, k, v = qkv.split([E, E, E], -1)
            heads = lambda t: t.contiguous().view(B, S, H, D).permute(0, 2, 1, 3).contiguous()
            o = kf.causal_attention(heads(q), heads(k), heads(v))
            o = o.permute(0, 2, 1, 3).contiguous().view(B, S, E)
        if self.fused:
            x1 = kf.gemm_residual(o, p["wo"], x, 1.0)
            xn2 = self.norm(x1, p["g2"])
            h = kf.gemm_glu(xn2, p["w1"], p["w3"])
            return kf.gemm_residual(h, p["w2"], x1, 1.0)
        x1 = x + kf.gemm(o, p["wo"], 1.0, 0.0)
        xn2 = self.norm(x1, p["g2"])
        h = kf.gemm(xn2, p["w1"], 1.0, 0.0) * kf.gemm(xn2, p["w3"], 1.0, 0.0)
        return x1 + kf.gemm(h, p["w2"], 1.0, 0.0)

    def loss(self, x):
        y = self.forward(x)
        return y.contiguous().view(-1).mean(0)  # full-tensor mean: the cross-shard reduce of §8e

    def step(self, x):
        """forward + backward; returns the scalar loss tensor ([1]); grads accumulate in params[...].grad()"""
        for t in self.params.values():
            t.zero_grad()
        loss = self.loss(x)
        one = kf.empty([1], loss.dtype(), loss.device())
        one.fill_(1.0)
        loss.backward(one)
        return loss

    def flops_per_sample(self, S: int) -> float:
        E, F = self.E, self.F
        gemm = 2.0 * S * (3 * E * E + E * E + 2 * E * F + F * E)
        attn = 4.0 * S * S * E / 2
        return 3.0 * (gemm + attn)  # fwd + bwd (2x)
