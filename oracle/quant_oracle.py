"""ctypes loader + numpy restatement of oracle/quant_oracle.c (TEST INFRASTRUCTURE).

PARITY UNPINNED for the rounding rule -- see the header of quant_oracle.c.  The step-size
function is pinned against nnc_core/common.py:28-46 by tests/golden/quant_stepsize.npz.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libquant_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "quant_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", _SO,
                               src, "-lm"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.nncq_stepsize.restype = ctypes.c_float
        _lib.nncq_stepsize.argtypes = [ctypes.c_int, ctypes.c_int]
        _lib.nncq_quant_urq.restype = ctypes.c_int
        _lib.nncq_quant_urq.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                        ctypes.c_int, ctypes.c_int]
        _lib.nncq_dequant.restype = None
        _lib.nncq_dequant.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                      ctypes.c_int, ctypes.c_int]
    return _lib


def stepsize(qp: int, qp_density: int) -> float:
    return float(lib().nncq_stepsize(int(qp), int(qp_density)))


def stepsize_py(qp: int, qp_density: int) -> float:
    """Pure-Python restatement of nnc_core/common.py:28-46."""
    k = 1 << qp_density
    return float((k + (qp & (k - 1))) * 2.0 ** ((qp >> qp_density) - qp_density))


def quant_urq(w: np.ndarray, qp: int, qp_density: int):
    w = np.ascontiguousarray(w, dtype=np.float32)
    out = np.zeros(w.shape, dtype=np.int32)
    used = lib().nncq_quant_urq(w.ctypes.data, out.ctypes.data, w.size, int(qp), int(qp_density))
    return out, int(used)


def quant_urq_np(w: np.ndarray, qp: int, qp_density: int) -> np.ndarray:
    """numpy restatement (no qp clip) used to cross-check the C code."""
    d = np.float32(stepsize_py(qp, qp_density))
    a = np.abs(w.astype(np.float32))
    m = ((a / d).astype(np.float32) + np.float32(0.5)).astype(np.float32).astype(np.int32)
    return np.where(w < 0, -m, m).astype(np.int32)


def dequant(lvl: np.ndarray, qp: int, qp_density: int) -> np.ndarray:
    lvl = np.ascontiguousarray(lvl, dtype=np.int32)
    out = np.zeros(lvl.shape, dtype=np.float32)
    lib().nncq_dequant(lvl.ctypes.data, out.ctypes.data, lvl.size, int(qp), int(qp_density))
    return out
