"""ncu / timing target for the v2 forward kernel (fine pass size: 4096 rays x 192)."""
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
n, S = int(os.environ.get("N_RAYS", "4096")), 192
g = torch.Generator().manual_seed(2)
o = 0.1 * torch.randn(n, 3, generator=g) + torch.tensor([0.0, 0.0, 4.0])
d = torch.randn(n, 3, generator=g)
d = -d / d.norm(dim=-1, keepdim=True)
rays = ops.pack_rays(o.to(dev), d.to(dev), False, 4, 4, 1.0, 2.0, 6.0)
z = torch.sort(2.0 + 4.0 * torch.rand(n, S, device=dev), -1).values.contiguous()
for _ in range(3):
    raw = packed.mlp_forward(pn, rays, z, impl=2)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    raw = packed.mlp_forward(pn, rays, z, impl=2)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"flags={os.environ.get('NERFQ_DEBUG_FLAGS', '0')} v2 fwd: {ms:.3f} ms {n * S * 1.186816e6 / ms / 1e9:.0f} TFLOP/s  mean|raw|={float(raw.abs().mean()):.4f}")
