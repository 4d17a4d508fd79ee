"""ctypes binding of the C-ABI library (include/nerfq.h).  Fails loudly when the library is missing:
there is no CPU or PyTorch fallback for any entry point."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# NERFQ_LIB: developer aid for A/B timing of two builds of the same sources on one box (profiles/ab_build.py)
LIB_PATH = os.environ.get("NERFQ_LIB") or os.path.join(_HERE, "libnerfq.so")

c_void_p, c_int, c_ll, c_float = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_float
c_ull = ctypes.c_ulonglong

_PROTOS = {
    "nerfq_packed_net_bytes": (c_ull, []),
    "nerfq_num_channels": (c_int, []),
    "nerfq_pack_net": (c_int, [c_void_p, ctypes.POINTER(c_void_p), ctypes.POINTER(c_float), c_int, c_void_p]),
    "nerfq_pack_status": (c_int, [c_void_p, c_void_p, c_void_p]),
    "nerfq_set_scale_bias": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "nerfq_mlp_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_ll, c_int, c_void_p, c_void_p, c_int, c_void_p]),
    "nerfq_mlp_save_bytes": (c_ull, [c_ll]),
}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                "The nerfq kernels have no fallback path.")
        _lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(_lib, name)
            fn.restype = res
            fn.argtypes = args
    return _lib


class NerfqError(RuntimeError):
    pass


def check(code: int, what: str):
    if code != 0:
        raise NerfqError(f"{what} failed with code {code}")
