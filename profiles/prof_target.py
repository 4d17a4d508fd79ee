"""Small target for ncu: one fine-network forward (with save) and backward at the cfg2 size (4096 rays x 192)."""
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
pp = os.environ.get("NERFQ_PINGPONG", "1") == "1"
for _ in range(2):
    raw = packed.mlp_forward(pn, rays, z, save=save, pingpong=pp)
    raw2 = packed.mlp_forward(pn, rays, z, pingpong=pp)
    acc = ops.mlp_backward(pn, torch.randn_like(raw) * 1e-5, raw, save)
torch.cuda.synchronize()
print("ok", float(raw.abs().mean()), float(acc.abs().max()))
