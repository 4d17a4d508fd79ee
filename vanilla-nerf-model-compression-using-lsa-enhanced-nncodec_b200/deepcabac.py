"""The quantiser half of the `deepCABAC` extension module, on the GPU.

`nnc_core/approximator/baseline.py` is the reference's only caller of the two functions on the hot path:

    encoder = deepCABAC.Encoder(); encoder.initCtxModels(cabac_unary_length_minus1, 0)             # :24, :42
    qp = encoder.quantLayer(w_f32, out_i32, dq_flag, qp_density, qp, lambda_scale,
                            cabac_unary_length_minus1, scan_order)                                  # :48-57
    decoder = deepCABAC.Decoder(); decoder.dequantLayer(out_f32, levels_i32, qp_density, qp, scan_order)   # :89, :98

`Encoder` / `Decoder` below keep those names, argument orders and in-place conventions (caller-allocated numpy
arrays; `quantLayer` returns the qp actually used, larger than the request when the levels would not fit int32,
baseline.py:60-62), so `baseline.approx` / `baseline.rec` run on them unmodified for `dq_flag == 0` with
`sys.modules['deepCABAC'] = nerfq_b200.deepcabac`.  Everything else the real module exports is the sequential entropy
coder, which stays on the host (out of scope this round): those methods raise, loudly.
Dependent (trellis-coded) quantisation, `dq_flag == 1`, is sequential too and is refused the same way.

Parity: bit-identical to the C restatement the tests check against; unpinned against the real deepCABAC, which is
absent here.
"""
import numpy as np
import torch

from . import ops

_ENTROPY = ("the DeepCABAC entropy coder is not part of the GPU path (sequential host code, SURVEY.md 8f); "
            "only quantLayer(dq_flag=0) / dequantLayer are provided")


def _device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("nerfq_b200.deepcabac needs a CUDA device (there is no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def _check(arr, dtype, what):
    if not isinstance(arr, np.ndarray) or arr.dtype != dtype:
        raise TypeError(f"{what} must be a numpy array of {np.dtype(dtype).name}")
    if not arr.flags["C_CONTIGUOUS"]:
        # the reference works around exactly this (transposed views) by allocating with np.zeros, baseline.py:34-39
        raise ValueError(f"{what} must be C-contiguous")


class Encoder:
    def __init__(self):
        self._ctx = None

    def initCtxModels(self, cabac_unary_length_minus1, param_opt_flag):
        self._ctx = (int(cabac_unary_length_minus1), int(param_opt_flag))     # state of the (absent) coder; kept for symmetry

    def quantLayer(self, weights, quantized, dq_flag, qp_density, qp, lambda_scale, cabac_unary_length_minus1, scan_order):
        """Uniform reconstruction quantisation of `weights` into `quantized` (in place); returns the qp used."""
        if int(dq_flag) != 0:
            raise NotImplementedError("dependent quantisation (dq_flag=1) is a sequential trellis search and stays on the host coder; "
                                      "the GPU quantiser implements dq_flag=0")
        _check(weights, np.float32, "weights")
        _check(quantized, np.int32, "quantized")
        if weights.shape != quantized.shape:
            raise ValueError("weights and quantized must have the same shape")
        if weights.size == 0:
            return int(qp)
        dev = _device()
        lv, qp_used = ops.quantize_urq(torch.from_numpy(weights).to(dev), int(qp), int(qp_density))
        quantized[...] = lv.cpu().numpy().reshape(quantized.shape)
        return int(qp_used)

    def iae_v(self, *a, **k):
        raise NotImplementedError(_ENTROPY)

    def encodeLayer(self, *a, **k):
        raise NotImplementedError(_ENTROPY)

    def finish(self, *a, **k):
        raise NotImplementedError(_ENTROPY)


class Decoder:
    def dequantLayer(self, out, levels, qp_density, qp, scan_order):
        """out[...] = levels * stepsize(qp, qp_density), in place (codebook.py:346-356 states the same reconstruction)."""
        _check(out, np.float32, "out")
        _check(levels, np.int32, "levels")
        if out.shape != levels.shape:
            raise ValueError("out and levels must have the same shape")
        if levels.size == 0:
            return
        dev = _device()
        out[...] = ops.dequantize(torch.from_numpy(levels).to(dev), int(qp), int(qp_density)).cpu().numpy().reshape(out.shape)

    def setStream(self, *a, **k):
        raise NotImplementedError(_ENTROPY)

    def initCtxModels(self, *a, **k):
        raise NotImplementedError(_ENTROPY)

    def iae_v(self, *a, **k):
        raise NotImplementedError(_ENTROPY)

    def decodeLayer(self, *a, **k):
        raise NotImplementedError(_ENTROPY)

    def decodeLayerAndCreateEPs(self, *a, **k):
        raise NotImplementedError(_ENTROPY)

    def setEntryPoints(self, *a, **k):
        raise NotImplementedError(_ENTROPY)

    def finish(self, *a, **k):
        raise NotImplementedError(_ENTROPY)
