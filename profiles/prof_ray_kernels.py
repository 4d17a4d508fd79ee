"""ncu target + live timing of the HBM-bound kernels at the cfg3 chunk size (32768 rays, 64 + 128 samples):
coarse depths, compositing forward / backward (coarse S=64 and fine S=192), importance sampling + merge, the batched
quantiser and the dequantiser.  Prints, per kernel, CUDA-event time, algorithmic bytes and the achieved GB/s
against the measured HBM peak.

    python profiles/prof_ray_kernels.py                 # timing table
    ncu --set full --clock-control none -k regex:"composite|sample_fine|quantize|dequantize|coarse_depths" \
        -c 40 -o gpurun_out/r02_ray_kernels python profiles/prof_ray_kernels.py --once
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nerfq_b200  # noqa
from nerfq_b200 import codec, model as nmodel, ops

once = "--once" in sys.argv
dev = torch.device("cuda:0")
torch.manual_seed(0)
n, S, Ni = 32768, 64, 128
g = torch.Generator().manual_seed(2)
o = 0.1 * torch.randn(n, 3, generator=g) + torch.tensor([0.0, 0.0, 4.0])
d = torch.randn(n, 3, generator=g)
d = -d / d.norm(dim=-1, keepdim=True)
rays = ops.pack_rays(o.to(dev), d.to(dev), False, 4, 4, 1.0, 2.0, 6.0)
t_rand = torch.rand(n, S, device=dev)
u = torch.rand(n, Ni, device=dev)
z0 = ops.coarse_depths(rays, S, False, t_rand)
raw0 = torch.randn(n, S, 4, device=dev)
raw1 = torch.randn(n, S + Ni, 4, device=dev)
rgb0, disp0, acc0, w0, _ = ops.composite_fwd(raw0, z0, rays, True, None)
z1, z_std, _ = ops.sample_fine(z0, w0, Ni, u)
d_rgb = torch.randn(n, 3, device=dev) * 1e-4

w = nmodel.LSA(nmodel.NeRFWrapper()).add_lsa_params().to(dev)
tensors = [p.detach().contiguous() for k, p in w.state_dict().items() if not k.endswith("weight_scaling")]
qps = [-20 if t.dim() == 2 else -75 for t in tensors]
n_q = sum(t.numel() for t in tensors)
flat = torch.cat([t.reshape(-1) for t in tensors]).contiguous()
lvl_flat, _ = ops.quantize_urq(flat, -20, 2)

peak = 6551.0
p = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(p):
    peak = float(json.load(open(p))["hbm_gbs"])

# (name, callable, algorithmic bytes per launch)
cases = [
    ("coarse_depths S=64 (perturb)", lambda: ops.coarse_depths(rays, S, False, t_rand), n * (44 + 4 * S + 4 * S)),
    ("composite_fwd S=64 (+weights)", lambda: ops.composite_fwd(raw0, z0, rays, True, None), n * (20 * S + 12 + 4 * S + 24)),
    ("composite_fwd S=192 (no weights)", lambda: ops.composite_fwd(raw1, z1, rays, True, None, want_weights=False), n * (20 * (S + Ni) + 12 + 24)),
    ("composite_bwd S=64", lambda: ops.composite_bwd(raw0, z0, rays, True, d_rgb, None), n * (20 * S + 24 + 16 * S)),
    ("composite_bwd S=192", lambda: ops.composite_bwd(raw1, z1, rays, True, d_rgb, None), n * (20 * (S + Ni) + 24 + 16 * (S + Ni))),
    ("sample_fine 64+128 (random u)", lambda: ops.sample_fine(z0, w0, Ni, u), n * (4 * S + 4 * S + 4 * Ni + 4 * (S + Ni) + 4)),
    ("sample_fine 64+128 (det)", lambda: ops.sample_fine(z0, w0, Ni, None), n * (4 * S + 4 * S + 4 * (S + Ni) + 4)),
    ("quantize_batch (48 tensors, 1.19 M values, absmax + quantise + reconstruct)",
     lambda: ops.quantize_batch(tensors, qps, 2, reconstruct_in_place=False), n_q * (4 + 4 + 4 + 4)),
    ("quantize_urq (1.19 M values, one tensor)", lambda: ops.quantize_urq(flat, -20, 2), n_q * (4 + 4 + 4)),
    ("dequantize (1.19 M values)", lambda: ops.dequantize(lvl_flat, -20, 2), n_q * 8),
    ("to8b 800x800x3", lambda: ops.to8b(img), 800 * 800 * 3 * 5),
]
img = torch.rand(800, 800, 3, device=dev)

if once:
    for name, fn, _ in cases:
        fn()
    torch.cuda.synchronize()
    print("ok")
    sys.exit(0)

flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > L2: every timed launch starts cold


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps


print(f"HBM peak {peak:.0f} GB/s (MEASURED_PEAKS.json); {n} rays; L2 flushed before every launch; times include the torch.empty of the outputs")
for name, fn, nbytes in cases:
    ms = timed(fn)
    gbs = nbytes / (ms * 1e-3) / 1e9
    print(f"{name:78s} {ms * 1e3:8.1f} us  {nbytes / 1e6:8.2f} MB  {gbs:7.0f} GB/s  {gbs / peak:5.2f} of peak")
