"""Where does the end-to-end gap of the LSA step come from?  Times the captured step fed from device tensors, from pinned
host tensors, from pageable host tensors, and with the host copies issued on a second stream one step ahead."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nerfq_b200  # noqa
from nerfq_b200 import codec, lsa, model as nmodel
from bench import synth_batch

dev = torch.device("cuda:0")
torch.manual_seed(0)
w = nmodel.LSA(nmodel.NeRFWrapper()).add_lsa_params().to(dev)
codec.quantize_model(w, -20)
n = 4096
o, d, t = synth_batch(n, 2)
rays_h = torch.stack([o, d], 0).pin_memory()
t_h = t.pin_memory()
rays_d, t_d = rays_h.to(dev), t_h.to(dev)
step = lsa.LSAStep(w, n, lr=1e-4, perturb=1.0, white_bkgd=True)
step.capture()


def timed(fn, steps=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, (time.perf_counter() - t0) * 1e3 / steps


print("device inputs            %.3f ms (wall %.3f)" % timed(lambda: step(rays_d, t_d)))
print("pinned host inputs       %.3f ms (wall %.3f)" % timed(lambda: step(rays_h, t_h)))
print("replay only (no copies)  %.3f ms (wall %.3f)" % timed(lambda: step.graph.replay()))


def only_copies():
    step.rays.copy_(rays_h, non_blocking=True)
    step.target.copy_(t_h, non_blocking=True)
print("copies only              %.3f ms (wall %.3f)" % timed(only_copies))


def copies_and_pack():
    step.rays.copy_(rays_h, non_blocking=True)
    step.target.copy_(t_h, non_blocking=True)
    step._pack(step.rays, step.packed_rays)
print("copies + pack            %.3f ms (wall %.3f)" % timed(copies_and_pack))


def d_copies_and_pack():
    step.rays.copy_(rays_d, non_blocking=True)
    step.target.copy_(t_d, non_blocking=True)
    step._pack(step.rays, step.packed_rays)
print("device copies + pack     %.3f ms (wall %.3f)" % timed(d_copies_and_pack))
print("pinned + loss.cpu()      %.3f ms (wall %.3f)" % timed(lambda: float(step(rays_h, t_h).cpu())))
print("device + loss.cpu()      %.3f ms (wall %.3f)" % timed(lambda: float(step(rays_d, t_d).cpu())))
