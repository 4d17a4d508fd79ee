"""Freeze the host coder's output (libnncabac.so through nerfq_b200.deepcabac) on small seeded inputs:

    python tests/golden/make_coder_golden.py        ->  tests/golden/coder_stream.npz

This fixture is SELF-pinned: it is produced by this repository's coder, not by upstream deepCABAC (absent here, DESIGN.md
section 5).  It does not establish compatibility with upstream streams; it makes any change of the written format -- context
selection, binarisation, the trellis search, the arithmetic coder's termination -- show up as a test failure instead of
going unnoticed, and it gives the decoder a stream that this build did not write itself.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import nerfq_b200  # noqa: E402,F401
from nerfq_b200 import deepcabac as dc  # noqa: E402

dc.DEVICE = "host"
rng = np.random.default_rng(20261018)
w_small = (rng.standard_normal((48, 37)) * 0.07).astype(np.float32)          # a weight matrix at qp -20
w_fine = (rng.standard_normal(300) * 0.9).astype(np.float32)                 # coded at qp -38: levels in the hundreds
bias = (rng.standard_normal(48) * 0.3).astype(np.float32)                    # a bias at qp -75: levels ~1e5
out = {"w_small": w_small, "w_fine": w_fine, "bias": bias}

enc = dc.Encoder()
layers = []
for name, w, dq, qp in (("bias", bias, 1, -75), ("w_small", w_small, 1, -20), ("w_fine", w_fine, 1, -38),
                         ("w_small", w_small, 0, -20), ("w_fine", w_fine, 0, -38)):
    lv = np.zeros(w.shape, dtype=np.int32)
    enc.initCtxModels(10, 0)
    used = enc.quantLayer(w, lv, dq, 2, qp, 0.0, 10, 0)
    enc.iae_v(8, used + 128)
    enc.encodeLayer(lv, dq, 0)
    key = f"levels_{name}_dq{dq}"
    out[key] = lv
    layers.append((key, dq, used))
out["stream"] = np.frombuffer(enc.finish().tobytes(), dtype=np.uint8)
out["layer_keys"] = np.array([k for k, _, _ in layers])
out["layer_dq"] = np.array([d for _, d, _ in layers], dtype=np.int32)
out["layer_qp"] = np.array([q for _, _, q in layers], dtype=np.int32)
np.savez_compressed(os.path.join(HERE, "coder_stream.npz"), **out)
print({k: (v.shape, str(v.dtype)) for k, v in out.items()}, "stream bytes:", out["stream"].size)
