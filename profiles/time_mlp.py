"""CUDA-event timing of the three MLP kernels (fine network, 4096 rays x 192 samples): forward, forward + save, backward."""
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


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for rep in range(2):
    a = timed(lambda: packed.mlp_forward(pn, rays, z))
    b = timed(lambda: packed.mlp_forward(pn, rays, z, save=save))
    c = timed(lambda: ops.mlp_backward(pn, d_raw, raw, save, acc))
    print(f"fwd {a:.3f} ms   fwd+save {b:.3f} ms   bwd {c:.3f} ms")
