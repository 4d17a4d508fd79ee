"""Cycle breakdown of the CTA-pair points-on-lanes forward kernel (mlp5_fwd.cu tracing instantiation, NERFQ_MLP_FWD=5), fine network, 4096 x 192 points."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nerfq_b200  # noqa
from nerfq_b200 import _lib, codec, model as nmodel, ops, packed

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
L = _lib.lib()
L.nerfq_mlp5_set_trace.argtypes = [ctypes.c_void_p]
L.nerfq_mlp5_set_trace.restype = None
for sv in (None,):
    for _ in range(2):
        packed.mlp_forward(pn, rays, z, save=sv)
    buf = torch.zeros(148 * 8, dtype=torch.int64, device=dev)
    L.nerfq_mlp5_set_trace(buf.data_ptr())
    packed.mlp_forward(pn, rays, z, save=sv)
    torch.cuda.synchronize()
    L.nerfq_mlp5_set_trace(None)
    iters = (n * S // 512 + 73) // 74
    t = buf.cpu().numpy().reshape(148, 8).astype(np.float64)
    lead, ep0, ep1 = t[0::2].mean(0) / iters, t[0::2, 4:7].mean(0) / iters, t[1::2, 4:7].mean(0) / iters
    print(f"{'save' if sv is not None else 'no save'}: per pair-iteration (2 groups; ideal MMA {82 * 512}): issuer total {lead[0]:.0f}  "
          f"waits: own slot {lead[1]:.0f} peer slot {lead[2]:.0f} operand {lead[3]:.0f}")
    print(f"   epilogue warp (CTA0 / CTA1): wait accumulator {ep0[0]:.0f} / {ep1[0]:.0f}  jobs {ep0[1]:.0f} / {ep1[1]:.0f}  hand-over {ep0[2]:.0f} / {ep1[2]:.0f}")
