"""ncu target: a few launches of the v3 fused MLP forward kernel (4096 rays x 192 samples, no save / save)."""
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
for _ in range(3):
    packed.mlp_forward(pn, rays, z)
for _ in range(2):
    packed.mlp_forward(pn, rays, z, save=save)
torch.cuda.synchronize()
print("ok")
