"""Dump the v2 forward kernel's internal timeline (debug_flags bit 3) for block 0."""
import os
import sys

import numpy as np
import torch

os.environ["NERFQ_DEBUG_FLAGS"] = str(8 | int(os.environ.get("EXTRA_FLAGS", "0")))
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
buf = torch.zeros(3 * 4096 * 8, dtype=torch.uint8, device=dev)
for _ in range(2):
    buf.zero_()
    packed.mlp_forward(pn, rays, z, impl=2, save=buf)
torch.cuda.synchronize()
t = buf.cpu().numpy().view(np.uint64).reshape(3, 2048, 2)
names = ["mma", "epi_lo(w4)", "epi_hi(w8)"]
t0 = int(t[0, 0, 1])
for w_ in range(3):
    rows = [(int(a), int(b) - t0) for a, b in t[w_] if b]
    print(f"== {names[w_]}: {len(rows)} events")
    prev = None
    out = []
    for tag, clk in rows[:int(os.environ.get("N_EVENTS", "140"))]:
        out.append(f"{tag}@{clk}" + (f"(+{clk - prev})" if prev is not None else ""))
        prev = clk
    print(" ".join(out))
