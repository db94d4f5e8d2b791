"""TEST INFRASTRUCTURE ONLY — run on a B200 box (`gpurun -- python oracle/make_golden_from_ref.py`).

Imports the UNMODIFIED reference build from oracle/_ref (built by oracle/Makefile from /root/reference sources),
feeds it the seeded cases of oracle/golden_cases.py through the reference's own Python API and stores its
outputs under gpurun_out/golden/ (copied into tests/golden/ and committed).  The reference resets the device on
import (launcher_cuda.h:289), so this must be the only CUDA user in its process."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "_ref"))
sys.path.insert(0, os.path.dirname(HERE))

import kfunca  # noqa: E402  (the reference)

from oracle.golden_cases import cases, cases_r2  # noqa: E402


def run(kind, inp, prm):
    g = {k: kfunca.from_numpy(v, 0) for k, v in inp.items()}
    if kind == "binary":
        r = eval(f"g['a'] {prm['op']} g['b']")
        return {"out": r.numpy()}
    if kind == "reduce":
        return {"out": getattr(g["x"], prm["op"])(prm["dim"]).numpy()}
    if kind == "permute":
        return {"out": g["x"].permute(*prm["dims"]).contiguous().numpy()}
    if kind == "sort":
        v, i = g["x"].sort(prm["dim"], prm["descending"])
        return {"values": v.numpy(), "indices": i.numpy()}
    if kind == "topk":
        v, i = g["x"].topk(prm["k"], prm["dim"], prm["largest"])
        return {"values": v.numpy(), "indices": i.numpy()}
    if kind == "gemm":
        return {"out": kfunca.gemm(g["a"], g["b"], 1.0, 0.0).numpy()}
    if kind == "attention":
        return {"out": kfunca.causal_attention(g["q"], g["k"], g["v"]).numpy()}
    if kind == "index_put":
        idx = [g[k] for k in ("i0", "i1", "i2") if k in g]
        g["x"].index_put_(idx, g["values"])
        return {"out": g["x"].numpy()}
    if kind == "norm_stat":
        # fresh pool memory: hand the pool a zeroed block first so the kernel's semaphores start at 0
        z = kfunca.zeros([1024], kfunca.int, 0)
        del z
        m, i = g["x"].norm_stat(0)
        return {"mean": m.numpy(), "invstd": i.numpy()}
    if kind == "mean_var":
        m, v = g["x"].mean_var(prm["dim"], prm["take_sqrt"])
        return {"mean": m.numpy(), "var": v.numpy()}
    raise ValueError(kind)


def main():
    out_dir = os.path.join(os.path.dirname(HERE), "gpurun_out", "golden")
    os.makedirs(out_dir, exist_ok=True)
    which = sys.argv[1] if len(sys.argv) > 1 else "r1"
    gen, fname = (cases_r2, "ref_outputs_r2.npz") if which == "r2" else (cases, "ref_outputs.npz")
    blob = {}
    for name, kind, inp, prm in gen():
        try:
            res = run(kind, inp, prm)
        except Exception as e:  # keep going: record what the reference cannot do
            print("REF-FAIL", name, repr(e))
            continue
        for k, v in res.items():
            blob[f"{name}.{k}"] = v
        print("ok", name, {k: v.shape for k, v in res.items()})
    np.savez_compressed(os.path.join(out_dir, fname), **blob)
    print("wrote", len(blob), "arrays")


if __name__ == "__main__":
    main()
