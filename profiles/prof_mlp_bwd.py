"""ncu target: LSA forward (save) + backward of the fine network, 4096 rays x 192 samples, v3 kernels."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nerfq_b200  # noqa
from nerfq_b200 import codec, model as nmodel, ops, packed

dev = torch.device("cuda:0")
torch.manual_seed(0)
w = nmodel.LSA(nmodel.NeRFWrapper()).add_lsa_params().to(dev)
codec.quantize_model(w, -20)
pn = w.model_fine.packed_net()
pn.set_scales(w.model_fine.scale_tensors())
n, S = 4096, 192
g = torch.Generator().manual_seed(2)
o = 0.1 * torch.randn(n, 3, generator=g) + torch.tensor([0.0, 0.0, 4.0])
d = torch.randn(n, 3, generator=g)
d = -d / d.norm(dim=-1, keepdim=True)
rays = ops.pack_rays(o.to(dev), d.to(dev), False, 4, 4, 1.0, 2.0, 6.0)
z = torch.sort(2.0 + 4.0 * torch.rand(n, S, device=dev), -1).values.contiguous()
save = torch.empty(packed.mlp_save_bytes(n * S), dtype=torch.uint8, device=dev)
raw = packed.mlp_forward(pn, rays, z, save=save)
d_raw = torch.randn_like(raw) * 1e-5
acc = torch.zeros(2436, device=dev)
for _ in range(3):
    ops.mlp_backward(pn, d_raw, raw, save, acc)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ops.mlp_backward(pn, d_raw, raw, save, acc)
e1.record()
torch.cuda.synchronize()
print(f"bwd {e0.elapsed_time(e1) / 5:.3f} ms")

import ctypes
import numpy as np
from nerfq_b200 import _lib
L = _lib.lib()
L.nerfq_mlp_set_trace_bwd.argtypes = [ctypes.c_void_p]
L.nerfq_mlp_set_trace_bwd.restype = None
buf = torch.zeros(148 * 8 + 148 * 32, dtype=torch.int64, device=dev)
L.nerfq_mlp_set_trace_bwd(buf.data_ptr())
ops.mlp_backward(pn, d_raw, raw, save, acc)
torch.cuda.synchronize()
e0.record()
for _ in range(5):
    ops.mlp_backward(pn, d_raw, raw, save, acc)
e1.record()
torch.cuda.synchronize()
print(f"bwd (tracing variant) {e0.elapsed_time(e1) / 5:.3f} ms")
L.nerfq_mlp_set_trace_bwd(None)
groups = (n * S // 256 + 147) // 148
t = buf.cpu().numpy()[:148 * 8].reshape(148, 8).astype(np.float64).mean(0) / groups
w = buf.cpu().numpy()[148 * 8:].reshape(148, 32).astype(np.float64).mean(0) / groups
print(f"trace per group: issuer total {t[0]:.0f} (ideal MMA {68 * 512}); waits WFull {t[1]:.0f} ActLo {t[2]:.0f} ActHi {t[3]:.0f}")
for tm in (0, 1):
    print(f"  team {tm} warp: prologue {w[8 * tm]:.0f}  views job {w[8 * tm + 1]:.0f}  AccReady waits {w[8 * tm + 2]:.0f}  jobs {w[8 * tm + 3]:.0f}")
