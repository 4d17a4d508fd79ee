"""GPU-box debug script: per-tensor LSA gradient errors against the oracle and timings of the LSA step."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nerfq_b200  # noqa
from nerfq_b200 import render as R, ops, packed
from oracle import render_oracle as ro
from tests.gpu_util import golden_wrapper
from tests.util import synth_rays, LAYERS, NETS


def main():
    dev = torch.device("cuda:0")
    w, p = golden_wrapper(dev, True)
    for name, prm in w.named_parameters():
        prm.requires_grad_(name.endswith("weight_scaling"))
    kw, _ = R.create_nerf(w, perturb=0.0, white_bkgd=True)
    for k in ("use_viewdirs", "ndc", "lindisp"):
        kw.pop(k)
    n = int(os.environ.get("N_ERR", "512"))
    batch = synth_rays(n, 2)
    target = torch.rand(n, 3, generator=torch.Generator().manual_seed(3))
    t0 = time.time()
    loss_ref, grads_ref, _ = ro.lsa_scale_grads(p, batch, target, white_bkgd=True)
    print(f"oracle {n} rays: {time.time() - t0:.2f}s loss {loss_ref:.6f}")
    out = R.render_rays(batch.to(dev), **kw)
    loss = R.img2mse(out["rgb_map"], target.to(dev)) + R.img2mse(out["rgb0"], target.to(dev))
    loss.backward()
    print("loss", float(loss.detach()))
    for net in NETS:
        for l in LAYERS:
            k = f"{net}.{l}.weight_scaling"
            g = getattr(w, net).get_submodule(l).weight_scaling.grad.detach().cpu().numpy().reshape(-1)
            r = grads_ref[k].numpy().reshape(-1)
            sc = max(np.abs(r).max(), 1e-30)
            print(f"  {k:45s} max|ref|={sc:.3e} relerr={np.abs(g - r).max() / sc:.5f} cos={np.dot(g, r) / (np.linalg.norm(g) * np.linalg.norm(r) + 1e-30):.6f}")
    # timings at the BASELINE cfg2 size
    n = 4096
    batch = synth_rays(n, 2).to(dev)
    target = torch.rand(n, 3, device=dev)
    params = [q for q in w.parameters() if q.requires_grad]
    opt = torch.optim.Adam(params, lr=1e-4)
    for pp in (0,):
        for it in range(3):
            out = R.render_rays(batch, **kw)
            loss = R.img2mse(out["rgb_map"], target) + R.img2mse(out["rgb0"], target)
            loss.backward(); opt.step(); opt.zero_grad()
        torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        iters = 5
        tf = tb = 0.0
        for it in range(iters):
            e[0].record()
            out = R.render_rays(batch, **kw)
            loss = R.img2mse(out["rgb_map"], target) + R.img2mse(out["rgb0"], target)
            e[1].record()
            loss.backward()
            e[2].record()
            opt.step(); opt.zero_grad()
            e[3].record()
            torch.cuda.synchronize()
            tf += e[0].elapsed_time(e[1]); tb += e[1].elapsed_time(e[2])
        print(f"LSA step 4096 rays: fwd {tf / iters:.3f} ms  bwd {tb / iters:.3f} ms  -> {1000.0 / ((tf + tb) / iters):.1f} steps/s (excl. Adam)")
    # kernel-level timing of the two MLP passes
    pn = w.model_fine.packed_net()
    z = torch.sort(2.0 + 4.0 * torch.rand(n, 192, device=dev), -1).values.contiguous()
    save = torch.empty(packed.mlp_save_bytes(n * 192), dtype=torch.uint8, device=dev)
    for pp in (0,):
        for name, fn in (("fwd nosave", lambda: packed.mlp_forward(pn, batch, z)),
                         ("fwd save", lambda: packed.mlp_forward(pn, batch, z, save=save))):
            fn(); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                fn()
            b.record(); torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 5
            print(f"  {name}: {ms:.3f} ms  {n * 192 * 1.186816e6 / ms / 1e9:.0f} TFLOP/s")
    raw = packed.mlp_forward(pn, batch, z, save=save)
    d_raw = torch.randn_like(raw) * 1e-5
    ops.mlp_backward(pn, d_raw, raw, save); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        ops.mlp_backward(pn, d_raw, raw, save)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print(f"  bwd: {ms:.3f} ms  {n * 192 * 2 * 557696 / ms / 1e9:.0f} TFLOP/s (dgrad flops)")


if __name__ == "__main__":
    main()
