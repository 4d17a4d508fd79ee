"""Per-CTA cycle counters of the v3 fused MLP forward kernel (where the MMA issuer and one epilogue warp wait)."""
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
L = _lib.lib()
L.nerfq_mlp_set_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
L.nerfq_mlp_set_trace.restype = None
buf = torch.zeros(148 * 8 + 148 * 32, dtype=torch.int64, device=dev)
save = torch.empty(packed.mlp_save_bytes(n * S), dtype=torch.uint8, device=dev)
for mode, kw, flags in ([("nosave", {}, 0), ("save", {"save": save}, 0)] +
                        [(f"nosave, job {j} timed", {}, j << 8) for j in (0, 1, 2, 3, 8, 9, 14, 15, 18, 19)] +
                        [(f"nosave, ablation {name} (flags {fl}), job 2 timed", {}, fl | (2 << 8)) for name, fl in
                         (("no TMEM loads", 1), ("no operand stores", 2), ("no math", 4), ("no async-proxy fence", 8), ("no alpha head", 16),
                          ("no encodings", 32), ("no loads/stores/math", 7), ("nothing but the barriers", 63))]):
    for _ in range(2):
        packed.mlp_forward(pn, rays, z, **kw)
    L.nerfq_mlp_set_trace(buf.data_ptr(), flags)
    buf.zero_()
    packed.mlp_forward(pn, rays, z, **kw)
    torch.cuda.synchronize()
    L.nerfq_mlp_set_trace(None, 0)
    t = buf.cpu().numpy()[:148 * 8].reshape(148, 8).astype(np.float64)
    groups = (n * S // 256 + 147) // 148
    w = buf.cpu().numpy()[148 * 8:].reshape(148, 32).astype(np.float64).mean(0) / groups
    m = t.mean(0) / groups
    print(f"{mode}: per group: total {m[0]:.0f} cyc (ideal MMA 38144); issuer waits: WFull {m[1]:.0f}  ActLo {m[2]:.0f}  ActHi {m[3]:.0f}; "
          f"epilogue warp 5: wait AccReady {m[4]:.0f}  StageFree {m[5]:.0f}  jobs {m[6]:.0f} (load+math+store part {m[7]:.0f}); "
          f"timed job: {w[0]:.0f} cycles, of which load+math+store {w[1]:.0f}")
