"""The `deepCABAC` extension module of the reference, re-provided: same class and method names, argument orders and
in-place numpy conventions, so `sys.modules['deepCABAC'] = nerfq_b200.deepcabac` lets the UNMODIFIED reference run
`nnc.compress_model()` / `nnc.decompress_model()` (tests/test_cpu_codec.py does exactly that).

Call sites in the reference:
    nnc_core/approximator/baseline.py:24-57   Encoder(); initCtxModels(n, param_opt); qp = quantLayer(w, out, dq_flag, qp_density,
                                              qp, lambda_scale, cabac_unary_length_minus1, scan_order)
    nnc_core/approximator/baseline.py:89-98   Decoder(); dequantLayer(out, levels, qp_density, qp, scan_order)
    nnc_core/coder/baseline.py:5-57           iae_v, initCtxModels, encodeLayer / decodeLayer / decodeLayerAndCreateEPs
    nnc_core/coder/__init__.py:118-140        Encoder.finish() -> np.uint8[]; Decoder.setStream(bytearray)
    nnc_core/coder/__init__.py:439-483        Decoder.setEntryPoints; Decoder.finish() -> bytes read

Where the work runs:
  * quantLayer(dq_flag=0) and dequantLayer -- the data-parallel part -- are the CUDA kernels nerfq_quantize_urq /
    nerfq_dequantize (include/nerfq.h).  There is no silent CPU fallback: without a CUDA device they raise, unless the
    caller has EXPLICITLY set `deepcabac.DEVICE = "host"` (a codec-only machine without a GPU; the host library computes
    the same values bit for bit -- tests/test_gpu_codec.py compares the two for every tensor and qp).
  * quantLayer(dq_flag=1) (8-state trellis search), the arithmetic coder and its context models are sequential by
    construction and run in the host library libnncabac.so (include/nncabac.h, csrc_host/nncabac.cpp), as BASELINE's
    north_star prescribes ("DeepCABAC entropy coding stays on the reference's sequential coder").
Both quantisers reconstruct as level * delta(qp), which is what dequantLayer applies and what the fused MLP kernels
consume as integer operands.

Parity: the uniform quantiser is bit-identical between the CUDA kernel, the host library and the C restatement the
tests check against.  UNPINNED against the real deepCABAC (absent here): the bitstream this coder
writes is decodable by this coder; byte-compatibility with upstream's streams cannot be established in this environment.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
HOST_LIB_PATH = os.path.join(_HERE, "libnncabac.so")
_c = ctypes
_host = None
DEVICE = "cuda"          # "cuda": elementwise (de)quantisation on the GPU kernels; "host": explicitly on libnncabac.so


def host_lib():
    """ctypes binding of libnncabac.so (include/nncabac.h); fails loudly when it has not been built."""
    global _host
    if _host is None:
        if not os.path.exists(HOST_LIB_PATH):
            raise RuntimeError(f"{HOST_LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`")
        L = _c.CDLL(HOST_LIB_PATH)
        vp, i32, i64, f32 = _c.c_void_p, _c.c_int, _c.c_int64, _c.c_float
        protos = {
            "nncabac_encoder_new": (vp, []), "nncabac_encoder_free": (None, [vp]),
            "nncabac_encoder_init_ctx": (i32, [vp, i32, i32]), "nncabac_encoder_iae_v": (i32, [vp, i32, i32]),
            "nncabac_quant_layer": (i32, [vp, vp, vp, i64, i32, i32, i32, f32, i32, i32, _c.POINTER(i32)]),
            "nncabac_encoder_encode_layer": (i32, [vp, vp, i64, i32, i32]),
            "nncabac_encoder_finish": (i32, [vp, _c.POINTER(vp), _c.POINTER(i64)]),
            "nncabac_decoder_new": (vp, []), "nncabac_decoder_free": (None, [vp]),
            "nncabac_decoder_set_stream": (i32, [vp, vp, i64]), "nncabac_decoder_init_ctx": (i32, [vp, i32]),
            "nncabac_decoder_iae_v": (i32, [vp, i32, _c.POINTER(i32)]),
            "nncabac_decoder_decode_layer": (i32, [vp, vp, i64, i32, i32]),
            "nncabac_decoder_finish": (i32, [vp, _c.POINTER(i64)]),
            "nncabac_dequant_layer": (i32, [vp, vp, i64, i32, i32]),
            "nncabac_stepsize": (f32, [i32, i32]),
        }
        for name, (res, args) in protos.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _host = L
    return _host


class CoderError(RuntimeError):
    pass


def _ok(code, what):
    if code == -2:
        raise NotImplementedError(f"{what}: block scan orders (scan_order > 0) are not implemented; the reference's NeRF path "
                                  "uses scan_order=0 (nnc/compression.py:82)")
    if code != 0:
        raise CoderError(f"{what} failed with code {code}" + (" (corrupt or exhausted stream)" if code == -3 else ""))


def _check(arr, dtype, what):
    if not isinstance(arr, np.ndarray) or arr.dtype != dtype:
        raise TypeError(f"{what} must be a numpy array of {np.dtype(dtype).name}")
    if not arr.flags["C_CONTIGUOUS"]:
        # the reference works around exactly this (transposed views) by allocating with np.zeros, baseline.py:34-39
        raise ValueError(f"{what} must be C-contiguous")


def _cuda_device():
    import torch
    if DEVICE not in ("cuda", "host"):
        raise ValueError("deepcabac.DEVICE must be 'cuda' or 'host'")
    if not torch.cuda.is_available():
        raise RuntimeError("quantLayer(dq_flag=0) / dequantLayer run on the GPU (nerfq_quantize_urq / nerfq_dequantize); "
                           "no CUDA device is available and there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


class Encoder:
    def __init__(self):
        self._L = host_lib()
        self._h = self._L.nncabac_encoder_new()
        if not self._h:
            raise MemoryError("nncabac_encoder_new")

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._L.nncabac_encoder_free(h)

    def initCtxModels(self, cabac_unary_length_minus1, param_opt_flag):
        _ok(self._L.nncabac_encoder_init_ctx(self._h, int(cabac_unary_length_minus1), int(bool(param_opt_flag))), "initCtxModels")

    def quantLayer(self, weights, quantized, dq_flag, qp_density, qp, lambda_scale, cabac_unary_length_minus1, scan_order):
        """Quantise `weights` into `quantized` (in place, same shape); returns the qp used (raised when the levels would not
        fit int32, baseline.py:60-62).  dq_flag=0: uniform reconstruction quantisation on the GPU.  dq_flag=1: dependent
        quantisation, trellis search on the host."""
        _check(weights, np.float32, "weights")
        _check(quantized, np.int32, "quantized")
        if weights.shape != quantized.shape:
            raise ValueError("weights and quantized must have the same shape")
        if weights.size == 0:
            return int(qp)
        if int(dq_flag) == 0 and DEVICE != "host":
            import torch
            from . import ops
            dev = _cuda_device()
            lv, qp_used = ops.quantize_urq(torch.from_numpy(weights).to(dev), int(qp), int(qp_density))
            quantized[...] = lv.cpu().numpy().reshape(quantized.shape)
            return int(qp_used)
        used = _c.c_int(0)
        _ok(self._L.nncabac_quant_layer(self._h, weights.ctypes.data, quantized.ctypes.data, weights.size, int(dq_flag), int(qp_density), int(qp),
                                        float(lambda_scale), int(cabac_unary_length_minus1), int(scan_order) if weights.ndim > 1 else 0,
                                        _c.byref(used)), "quantLayer")
        return int(used.value)

    def iae_v(self, n_bits, value):
        _ok(self._L.nncabac_encoder_iae_v(self._h, int(n_bits), int(value)), "iae_v")

    def encodeLayer(self, levels, dq_flag, scan_order):
        _check(levels, np.int32, "levels")
        _ok(self._L.nncabac_encoder_encode_layer(self._h, levels.ctypes.data, levels.size, int(dq_flag), int(scan_order) if levels.ndim > 1 else 0),
            "encodeLayer")

    def finish(self):
        """Terminates the codeword; returns the bytes as np.uint8 (the reference calls .tobytes() on it)."""
        data, size = _c.c_void_p(), _c.c_int64()
        _ok(self._L.nncabac_encoder_finish(self._h, _c.byref(data), _c.byref(size)), "finish")
        if size.value == 0:
            return np.zeros(0, dtype=np.uint8)
        return np.ctypeslib.as_array(_c.cast(data, _c.POINTER(_c.c_uint8)), shape=(size.value,)).copy()


class Decoder:
    def __init__(self):
        self._L = host_lib()
        self._h = self._L.nncabac_decoder_new()
        if not self._h:
            raise MemoryError("nncabac_decoder_new")

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._L.nncabac_decoder_free(h)

    def setStream(self, stream):
        buf = np.frombuffer(bytes(stream), dtype=np.uint8)
        _ok(self._L.nncabac_decoder_set_stream(self._h, buf.ctypes.data if buf.size else None, buf.size), "setStream")

    def initCtxModels(self, cabac_unary_length_minus1):
        _ok(self._L.nncabac_decoder_init_ctx(self._h, int(cabac_unary_length_minus1)), "initCtxModels")

    def iae_v(self, n_bits):
        v = _c.c_int(0)
        _ok(self._L.nncabac_decoder_iae_v(self._h, int(n_bits), _c.byref(v)), "iae_v")
        return int(v.value)

    def decodeLayer(self, out, dq_flag, scan_order):
        _check(out, np.int32, "out")
        _ok(self._L.nncabac_decoder_decode_layer(self._h, out.ctypes.data, out.size, int(dq_flag), int(scan_order) if out.ndim > 1 else 0),
            "decodeLayer")

    def decodeLayerAndCreateEPs(self, out, dq_flag, scan_order):
        """Entry points exist for block scans only (nnc_core/coder/__init__.py:132-135 calls this when scan_order > 0)."""
        self.decodeLayer(out, dq_flag, scan_order)
        return np.zeros(0, dtype=np.uint64)

    def setEntryPoints(self, entry_points):
        if len(entry_points):
            raise NotImplementedError("entry points belong to block scans (scan_order > 0), which are not implemented")

    def dequantLayer(self, out, levels, qp_density, qp, scan_order):
        """out[...] = levels * stepsize(qp, qp_density), in place (codebook.py:346-356 states the same reconstruction)."""
        _check(out, np.float32, "out")
        _check(levels, np.int32, "levels")
        if out.shape != levels.shape:
            raise ValueError("out and levels must have the same shape")
        if levels.size == 0:
            return
        if DEVICE == "host":
            _ok(self._L.nncabac_dequant_layer(out.ctypes.data, levels.ctypes.data, levels.size, int(qp_density), int(qp)), "dequantLayer")
            return
        import torch
        from . import ops
        dev = _cuda_device()
        out[...] = ops.dequantize(torch.from_numpy(levels).to(dev), int(qp), int(qp_density)).cpu().numpy().reshape(out.shape)

    def finish(self):
        n = _c.c_int64(0)
        _ok(self._L.nncabac_decoder_finish(self._h, _c.byref(n)), "finish")
        return int(n.value)


# ---- host-side entry points used where a GPU round trip makes no sense (tests, tools) --------------------------------
def host_quant_layer(weights: np.ndarray, dq_flag: int, qp_density: int, qp: int, lambda_scale: float = 0.0):
    """The host library's quantiser for either dq_flag (uniform: same arithmetic as the CUDA kernel).  -> (levels, qp used)"""
    _check(weights, np.float32, "weights")
    out = np.zeros(weights.shape, dtype=np.int32)
    used = _c.c_int(int(qp))
    if weights.size:
        _ok(host_lib().nncabac_quant_layer(None, weights.ctypes.data, out.ctypes.data, weights.size, int(dq_flag), int(qp_density), int(qp),
                                           float(lambda_scale), 10, 0, _c.byref(used)), "nncabac_quant_layer")
    return out, int(used.value)


def host_dequant_layer(levels: np.ndarray, qp_density: int, qp: int) -> np.ndarray:
    _check(levels, np.int32, "levels")
    out = np.zeros(levels.shape, dtype=np.float32)
    if levels.size:
        _ok(host_lib().nncabac_dequant_layer(out.ctypes.data, levels.ctypes.data, levels.size, int(qp_density), int(qp)), "nncabac_dequant_layer")
    return out
