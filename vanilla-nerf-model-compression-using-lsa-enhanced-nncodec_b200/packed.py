"""Device-resident packed form of one NeRF network (csrc/net_layout.h) and the raw kernel entry
points that consume it.  PyTorch is used for device memory and the current stream only."""
import ctypes
import warnings
from typing import Dict, Optional, Sequence

import torch

from . import _lib

LAYER_NAMES = tuple([f"pts_linears.{i}" for i in range(8)] +
                    ["alpha_linear", "feature_linear", "views_linears.0", "rgb_linear"])
LAYER_OUT = (256,) * 8 + (1, 256, 128, 3)
LAYER_IN = (63, 256, 256, 256, 256, 319, 256, 256, 256, 256, 283, 128)
# flat per-channel order used by the kernels: pts0..7, feature, views, alpha, rgb
CHANNEL_ORDER = tuple(range(8)) + (9, 10, 8, 11)
NUM_CHANNELS = 2436


# Integer levels are the tensor-core operands, exact in fp16 up to EXACT_LEVEL.  Beyond it nerfq_pack_net rounds a level to
# 11 significant bits (relative error <= 2^-12, what any fp16 weight gets; csrc/pack.cu): "warn" reports that once per
# packed network, "raise" refuses, "ignore" stays silent.
EXACT_LEVEL = 2048
LEVEL_RANGE_POLICY = "warn"


class LevelRangeError(_lib.NerfqError):
    pass


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def flatten_channels(per_layer: Sequence[torch.Tensor]) -> torch.Tensor:
    """12 per-layer [out]/[out,1] tensors (layer order) -> flat [2436] in kernel channel order."""
    return torch.cat([per_layer[l].reshape(-1).float() for l in CHANNEL_ORDER]).contiguous()


def split_channels(flat: torch.Tensor):
    """Inverse of flatten_channels: returns a list of 12 tensors in layer order."""
    out = [None] * 12
    pos = 0
    for l in CHANNEL_ORDER:
        out[l] = flat[pos:pos + LAYER_OUT[l]]
        pos += LAYER_OUT[l]
    return out


class PackedNet:
    """One network packed for the fused MLP kernels.

    weights: 12 tensors [out,in] on the GPU, either int32 quantisation levels or float32 values.
    deltas:  12 step sizes (1.0 for unquantised float weights).
    """

    def __init__(self, weights: Sequence[torch.Tensor], deltas: Sequence[float], biases: Sequence[torch.Tensor],
                 scales: Optional[Sequence[torch.Tensor]] = None):
        L = _lib.lib()
        assert len(weights) == 12 and len(deltas) == 12 and len(biases) == 12
        dev = weights[0].device
        assert dev.type == "cuda", "PackedNet needs CUDA tensors"
        is_int = weights[0].dtype == torch.int32
        ws = []
        for l, w in enumerate(weights):
            assert tuple(w.shape) == (LAYER_OUT[l], LAYER_IN[l]), (l, tuple(w.shape))
            assert w.dtype == (torch.int32 if is_int else torch.float32)
            ws.append(w.contiguous())
        self.device = dev
        self.buf = torch.empty(int(L.nerfq_packed_net_bytes()), dtype=torch.uint8, device=dev)
        self.is_int = is_int
        self._max_abs = torch.empty(12, dtype=torch.float32, device=dev)
        self.bias_flat = torch.empty(NUM_CHANNELS, dtype=torch.float32, device=dev)
        self.scale_flat = None
        self._pack(ws, deltas, biases)
        self.set_scales(scales)

    def _pack(self, ws, deltas, biases):
        L = _lib.lib()
        ptrs = (ctypes.c_void_p * 12)(*[w.data_ptr() for w in ws])
        dl = (ctypes.c_float * 12)(*[float(d) for d in deltas])
        _lib.check(L.nerfq_pack_net(self.buf.data_ptr(), ptrs, dl, int(self.is_int), _stream()), "nerfq_pack_net")
        self._keep = ws
        _lib.check(L.nerfq_pack_status(self.buf.data_ptr(), self._max_abs.data_ptr(), _stream()), "nerfq_pack_status")
        self._range_checked = False
        if not torch.cuda.is_current_stream_capturing():
            self.check_range()
        self.bias_flat.copy_(flatten_channels(biases))

    def repack(self, weights: Sequence[torch.Tensor], deltas: Sequence[float], biases: Sequence[torch.Tensor]):
        """Pack new values of the same kind (int32 levels / float32) into the SAME buffers.  Like a fresh PackedNet the
        result needs set_scales() before it is used (nerfq_set_scale_bias must follow every nerfq_pack_net); every caller in
        this package goes through render._refresh, which does that."""
        assert len(weights) == 12 and all(w.dtype == (torch.int32 if self.is_int else torch.float32) and w.is_contiguous() for w in weights)
        self._pack(list(weights), deltas, biases)

    def max_abs_per_layer(self):
        """max |level| (or |weight|) of the 12 layers as packed (synchronises)."""
        return [float(v) for v in self._max_abs.cpu()]

    def check_range(self):
        """Applies LEVEL_RANGE_POLICY to the operand range found by nerfq_pack_net.  Non-finite weights always raise."""
        self._range_checked = True
        mx = self.max_abs_per_layer()
        bad = [LAYER_NAMES[l] for l, v in enumerate(mx) if not (v <= 3.0e38)]
        if bad:
            raise LevelRangeError(f"non-finite weights in {bad}")
        if not self.is_int or LEVEL_RANGE_POLICY == "ignore":
            return
        over = {LAYER_NAMES[l]: int(v) for l, v in enumerate(mx) if v > EXACT_LEVEL}
        if over:
            msg = (f"quantisation levels beyond +-{EXACT_LEVEL} are not exact fp16 tensor-core operands; they are rounded to 11 "
                   f"significant bits (relative error <= 2^-12): max |level| {over}")
            if LEVEL_RANGE_POLICY == "raise":
                raise LevelRangeError(msg)
            warnings.warn(msg, RuntimeWarning, stacklevel=3)

    def set_scales(self, scales: Optional[Sequence[torch.Tensor]] = None, flat: Optional[torch.Tensor] = None):
        """(Re)load the LSA scales: the epilogue constant becomes delta*scale per output channel."""
        if flat is None and scales is not None:
            flat = flatten_channels(scales).to(self.device)
        self.scale_flat = flat
        _lib.check(_lib.lib().nerfq_set_scale_bias(self.buf.data_ptr(), flat.data_ptr() if flat is not None else None,
                                                   self.bias_flat.data_ptr(), _stream()), "nerfq_set_scale_bias")

    @property
    def ptr(self) -> int:
        return self.buf.data_ptr()


def mlp_save_bytes(n_points: int) -> int:
    return int(_lib.lib().nerfq_mlp_save_bytes(n_points))


def mlp_forward(net: PackedNet, rays: torch.Tensor, z: torch.Tensor, save: Optional[torch.Tensor] = None,
                max_ctas: int = 0) -> torch.Tensor:
    """raw[N,S,4] = MLP(gamma(o + d z), gamma(viewdir)) for rays [N,11] and depths z [N,S]."""
    assert rays.is_cuda and rays.dtype == torch.float32 and rays.shape[1] == 11 and rays.is_contiguous()
    assert z.is_cuda and z.dtype == torch.float32 and z.is_contiguous() and z.shape[0] == rays.shape[0]
    n, s = z.shape
    raw = torch.empty((n, s, 4), dtype=torch.float32, device=rays.device)
    _lib.check(_lib.lib().nerfq_mlp_forward(net.ptr, rays.data_ptr(), z.data_ptr(), n, s, raw.data_ptr(),
                                            save.data_ptr() if save is not None else None, max_ctas, _stream()),
               "nerfq_mlp_forward")
    return raw
