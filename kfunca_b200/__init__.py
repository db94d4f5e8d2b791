"""kfunca_b200 — B200-native (sm_100a) implementation of xytpai/kfunca's tensor-operator API.

`import kfunca_b200 as kfunca` is the drop-in: every name the reference's pybind module exports
(/root/reference/src/register.cpp:59-225) is re-exported here from the compiled extension `_kfunca`,
which itself is written only against the C ABI in include/kfunca_b200.h.

There is no Python or CPU fallback: if the extension is missing the import fails loudly.
"""
from __future__ import annotations

import importlib
import os

_HERE = os.path.dirname(os.path.abspath(__file__))

try:
    _C = importlib.import_module("kfunca_b200._kfunca")
except ImportError as e:  # pragma: no cover - exercised only on a broken checkout
    raise ImportError(
        "kfunca_b200: the compiled extension kfunca_b200/_kfunca*.so (and libkfunca_b200.so) is missing or "
        "failed to load. Build it in-tree with `python -m kfunca_b200.build` (needs nvcc; no GPU required). "
        "There is deliberately no CPU fallback. Original error: %s" % (e,)
    ) from e

from ._kfunca import *  # noqa: F401,F403  (dtype values are exported at module level like the reference)
from ._kfunca import (  # noqa: F401
    tensor, dtype, empty, zeros, empty_like, from_numpy, to_numpy, causal_attention, gemm, cat,
    device_info, memstat,
)

LIB_PATH = os.path.join(_HERE, "libkfunca_b200.so")
__version__ = "0.1.0"
