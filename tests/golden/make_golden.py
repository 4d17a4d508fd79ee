"""Generate the golden fixtures in tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Every array written here is an output of a reference function (cited per case) on seeded
synthetic inputs; the inputs are stored next to the outputs so the tests never need the
reference or an RNG that matches this machine's.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _ref_import import import_reference  # noqa: E402

ref = import_reference()
torch.set_num_threads(8)

LAYERS = [f"pts_linears.{i}" for i in range(8)] + ["alpha_linear", "feature_linear",
                                                     "views_linears.0", "rgb_linear"]


def stepsize(qp, d):
    return ref.common.get_stepsize_from_qp(qp, d)


def build_model():
    """NeRFWrapper + LSA (utils.py:84-106, transforms.py:113-168) with weights snapped to integer
    levels at qp=-20 so the fixture is small (int8) and exactly representable."""
    torch.manual_seed(0)
    wrapper = ref.utils.NeRFWrapper()
    model = ref.transforms.LSA(wrapper).add_lsa_params()
    delta = np.float32(stepsize(-20, 2))
    g = torch.Generator().manual_seed(1234)
    store = {}
    with torch.no_grad():
        for net in ("model", "model_fine"):
            for lname in LAYERS:
                mod = model.get_submodule(f"{net}.{lname}")
                lv = torch.round(mod.weight / float(delta)).clamp(-100, 100)
                if lname == "alpha_linear":
                    lv = lv * 8          # sharper densities so compositing is exercised
                    mod.bias += 0.5
                mod.weight.copy_(lv * float(delta))
                mod.weight_scaling.copy_(1.0 + 0.05 * torch.randn(mod.weight_scaling.shape, generator=g))
                key = f"{net}.{lname}"
                store[key + ".levels"] = lv.numpy().astype(np.int8)
                store[key + ".bias"] = mod.bias.numpy().copy()
                store[key + ".weight_scaling"] = mod.weight_scaling.numpy().copy()
    store["delta"] = np.float32(delta)
    np.savez_compressed(os.path.join(HERE, "model_qm20.npz"), **store)
    return model


def build_model_trained():
    """A 'trained-like' NeRFWrapper + LSA at qp=-38 (delta 1.46e-3): what a real checkpoint does to the fused path
    that a random-init net does not -- quantisation levels beyond 2048 (not exact as fp16 operands) and densities in
    the hundreds.  A random MLP with large weights is a chaotic function (the reference then disagrees with ITSELF by
    more than the gate under fp32 re-association), so the large levels are introduced through ReLU homogeneity:
    pts_linears.{2,4,6} get weight and bias x64/x80/x100 (levels up to 2731/3413/4267) and the following layer's LSA
    scale is divided by the same factor -- the function stays the smooth random-init field.  alpha_linear x50 (levels to
    2133) with bias 0.7 gives sigma ~0.9..1.0 (weights spread over many samples); the `dense` variant uses x60 and bias
    150 (sigma ~150: saturated alphas, the fixed-point sigma sum far from zero).  rgb_linear x8 for colour contrast."""
    torch.manual_seed(11)
    wrapper = ref.utils.NeRFWrapper()
    model = ref.transforms.LSA(wrapper).add_lsa_params()
    delta = np.float32(stepsize(-38, 2))
    g = torch.Generator().manual_seed(4321)
    store = {"delta": np.float32(delta)}
    with torch.no_grad():
        for net in ("model", "model_fine"):
            for lname in LAYERS:
                mod = model.get_submodule(f"{net}.{lname}")
                mod.weight_scaling.copy_(1.0 + 0.05 * torch.randn(mod.weight_scaling.shape, generator=g))
            for li, c in ((6, 100.0), (2, 64.0), (4, 80.0)):
                a = model.get_submodule(f"{net}.pts_linears.{li}")
                b = model.get_submodule(f"{net}.pts_linears.{li + 1}")
                a.weight.mul_(c); a.bias.mul_(c); b.weight_scaling.mul_(1.0 / c)
            model.get_submodule(f"{net}.rgb_linear").weight.mul_(8.0)
            al = model.get_submodule(f"{net}.alpha_linear")
            base = al.weight.clone()
            lv_dense = torch.round(base * 60.0 / float(delta))
            store[f"{net}.alpha_linear.levels_dense"] = lv_dense.numpy().astype(np.int16)
            al.weight.copy_(base * 50.0)
            al.bias.add_(0.7)
            for lname in LAYERS:
                mod = model.get_submodule(f"{net}.{lname}")
                lv = torch.round(mod.weight / float(delta))
                assert float(lv.abs().max()) < 32767
                mod.weight.copy_(lv * float(delta))
                key = f"{net}.{lname}"
                store[key + ".levels"] = lv.numpy().astype(np.int16)
                store[key + ".bias"] = mod.bias.numpy().copy()
                store[key + ".weight_scaling"] = mod.weight_scaling.numpy().copy()
    store["dense_alpha_bias_shift"] = np.float32(150.0 - 0.7)
    np.savez_compressed(os.path.join(HERE, "model_trained_qm38.npz"), **store)
    return model, delta, store


def case_render_trained():
    """run_nerf.render (run_nerf.py:81-158) on the trained-like model, both alpha variants, 64+128 samples."""
    model, delta, store = build_model_trained()
    o, d = synth_rays(160, 21)
    out = {"rays_o": np_(o), "rays_d": np_(d)}
    for variant in ("spread", "dense"):
        if variant == "dense":
            with torch.no_grad():
                for net in ("model", "model_fine"):
                    al = model.get_submodule(f"{net}.alpha_linear")
                    al.weight.copy_(torch.from_numpy(store[f"{net}.alpha_linear.levels_dense"].astype(np.float32)) * float(delta))
                    al.bias.add_(float(store["dense_alpha_bias_shift"]))
        with torch.no_grad():
            rgb, disp, acc, ex = ref.run_nerf.render(4, 4, None, chunk=80, rays=(o, d), near=2.0, far=6.0, ndc=False,
                                                     retraw=True, **render_kwargs(model))
        out.update({f"{variant}_rgb": np_(rgb), f"{variant}_disp": np_(disp), f"{variant}_acc": np_(acc),
                    f"{variant}_rgb0": np_(ex["rgb0"]), f"{variant}_acc0": np_(ex["acc0"]), f"{variant}_disp0": np_(ex["disp0"]),
                    f"{variant}_z_std": np_(ex["z_std"]), f"{variant}_sigma_minmax": np.array([float(ex["raw"][..., 3].min()),
                                                                                             float(ex["raw"][..., 3].max())], dtype=np.float32)})
    np.savez_compressed(os.path.join(HERE, "render_trained.npz"), **out)


def synth_rays(n, seed, near=2.0, far=6.0):
    g = torch.Generator().manual_seed(seed)
    o = 0.1 * torch.randn(n, 3, generator=g) + torch.tensor([0.0, 0.0, 4.0])
    d = torch.randn(n, 3, generator=g)
    d = -d / torch.norm(d, dim=-1, keepdim=True) * (0.8 + 0.4 * torch.rand(n, 1, generator=g))
    return o, d


def render_kwargs(model, perturb=0.0, raw_noise_std=0.0, n_importance=128, white_bkgd=True, **extra):
    embed_fn, _ = ref.helpers.get_embedder(10, 0)
    embeddirs_fn, _ = ref.helpers.get_embedder(4, 0)
    query = lambda inputs, viewdirs, fn: ref.run_nerf.run_network(
        inputs, viewdirs, fn, embed_fn=embed_fn, embeddirs_fn=embeddirs_fn, netchunk=65536)
    kw = dict(network_query_fn=query, perturb=perturb, N_importance=n_importance,
              network_fine=model.model_fine, N_samples=64, network_fn=model.model,
              use_viewdirs=True, white_bkgd=white_bkgd, raw_noise_std=raw_noise_std)
    kw.update(extra)
    return kw


def np_(x):
    return x.detach().cpu().numpy()


def case_render_blender(model):
    """run_nerf.render (run_nerf.py:81-158) on a ray batch, ndc=False, white background."""
    o, d = synth_rays(96, 1)
    with torch.no_grad():
        rgb, disp, acc, ex = ref.run_nerf.render(4, 4, None, chunk=64, rays=(o, d), near=2.0, far=6.0,
                                                 ndc=False, retraw=True, **render_kwargs(model))
    np.savez_compressed(os.path.join(HERE, "render_blender.npz"), rays_o=np_(o), rays_d=np_(d),
                        rgb=np_(rgb), disp=np_(disp), acc=np_(acc), rgb0=np_(ex["rgb0"]),
                        disp0=np_(ex["disp0"]), acc0=np_(ex["acc0"]), z_std=np_(ex["z_std"]),
                        raw=np_(ex["raw"]))


def case_render_ndc(model):
    """render(c2w=...) with get_rays + ndc_rays (run_nerf_helpers.py:71-115), LLFF-style."""
    H, W, focal = 6, 8, 7.5
    K = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]], dtype=np.float32)
    c2w = torch.tensor([[0.98, 0.05, -0.19, 0.3], [-0.03, 0.99, 0.10, -0.2], [0.19, -0.09, 0.97, 0.1]])
    with torch.no_grad():
        rgb, disp, acc, ex = ref.run_nerf.render(H, W, K, chunk=32, c2w=c2w, ndc=True, near=0.0, far=1.0,
                                                 **render_kwargs(model, white_bkgd=False))
        ro, rd = ref.helpers.get_rays(H, W, K, c2w)
        no, nd = ref.helpers.ndc_rays(H, W, K[0][0], 1.0, ro, rd)
    np.savez_compressed(os.path.join(HERE, "render_ndc.npz"), H=H, W=W, K=K, c2w=np_(c2w),
                        rays_o=np_(ro), rays_d=np_(rd), ndc_o=np_(no), ndc_d=np_(nd),
                        rgb=np_(rgb), disp=np_(disp), acc=np_(acc), rgb0=np_(ex["rgb0"]),
                        disp0=np_(ex["disp0"]), acc0=np_(ex["acc0"]), z_std=np_(ex["z_std"]))


def case_render_perturb(model):
    """render_rays with perturb=1, raw_noise_std=1 under the reference's pytest=True hooks
    (run_nerf.py:319-322,398-401; run_nerf_helpers.py:137-145): every draw is
    np.random.seed(0); np.random.rand(...), which the test replays explicitly."""
    n = 40
    o, d = synth_rays(n, 5)
    vd = d / torch.norm(d, dim=-1, keepdim=True)
    batch = torch.cat([o, d, 2.0 * torch.ones(n, 1), 6.0 * torch.ones(n, 1), vd], -1)
    kw = render_kwargs(model, perturb=1.0, raw_noise_std=1.0, white_bkgd=False)
    kw.pop("use_viewdirs")
    with torch.no_grad():
        out = ref.run_nerf.render_rays(batch, pytest=True, retraw=True, **kw)
    np.random.seed(0)
    t_rand = np.random.rand(n, 64)
    np.random.seed(0)
    u = np.random.rand(n, 128)
    np.random.seed(0)
    noise0 = np.random.rand(n, 64) * 1.0
    np.random.seed(0)
    noise1 = np.random.rand(n, 192) * 1.0
    np.savez_compressed(os.path.join(HERE, "render_perturb.npz"), ray_batch=np_(batch),
                        t_rand=t_rand.astype(np.float32), u=u.astype(np.float32),
                        noise0=noise0.astype(np.float32), noise1=noise1.astype(np.float32),
                        **{k: np_(v) for k, v in out.items()})


def case_functions():
    """Function-level vectors: raw2outputs (run_nerf.py:285-345), sample_pdf
    (run_nerf_helpers.py:119-163), Embedder (run_nerf_helpers.py:18-67)."""
    g = torch.Generator().manual_seed(7)
    n, s = 50, 64
    raw = torch.randn(n, s, 4, generator=g) * 2.0
    raw[:5, :, 3] = -1.0                      # rays with zero density everywhere -> NaN disparity
    z, _ = torch.sort(2.0 + 4.0 * torch.rand(n, s, generator=g), -1)
    d = torch.randn(n, 3, generator=g)
    out = {}
    for wb in (False, True):
        rgb, disp, acc, w, depth = ref.run_nerf.raw2outputs(raw, z, d, 0, wb)
        out.update({f"c_rgb_{int(wb)}": np_(rgb), f"c_disp_{int(wb)}": np_(disp), f"c_acc_{int(wb)}": np_(acc),
                    f"c_w_{int(wb)}": np_(w), f"c_depth_{int(wb)}": np_(depth)})
    bins = 0.5 * (z[:, 1:] + z[:, :-1])
    wts = torch.rand(n, s - 2, generator=g) ** 4
    wts[7] = 0.0                              # flat pdf
    wts[8, :] = 0.0
    wts[8, 20] = 1.0                          # single spike -> knife-edge denom branch
    det = ref.helpers.sample_pdf(bins, wts, 128, det=True)
    u = torch.rand(n, 128, generator=g)
    orig_rand = torch.rand
    torch.rand = lambda *a, **k: u            # feed identical draws (reference calls torch.rand)
    try:
        rnd = ref.helpers.sample_pdf(bins, wts, 128, det=False)
    finally:
        torch.rand = orig_rand
    embed, _ = ref.helpers.get_embedder(10, 0)
    embed_d, _ = ref.helpers.get_embedder(4, 0)
    x = torch.randn(33, 3, generator=g) * 3.0
    np.savez_compressed(os.path.join(HERE, "functions.npz"), raw=np_(raw), z=np_(z), d=np_(d), bins=np_(bins),
                        wts=np_(wts), pdf_det=np_(det), pdf_u=np_(u), pdf_rnd=np_(rnd), pe_x=np_(x),
                        pe10=np_(embed(x)), pe4=np_(embed_d(x)), **out)


def case_lsa_step(model):
    """One LSA objective evaluation + autograd into the 24 weight_scaling tensors
    (run_nerf.py:739-756 with framework/pytorch_model/__init__.py:1129-1145 freezing the rest)."""
    n = 48
    o, d = synth_rays(n, 2)
    g = torch.Generator().manual_seed(3)
    target = torch.rand(n, 3, generator=g)
    for name, p in model.named_parameters():
        p.requires_grad_(name.endswith("weight_scaling"))
    rgb, disp, acc, ex = ref.run_nerf.render(4, 4, None, chunk=32768, rays=(o, d), near=2.0, far=6.0,
                                             ndc=False, retraw=True, **render_kwargs(model))
    loss = ref.helpers.img2mse(rgb, target) + ref.helpers.img2mse(ex["rgb0"], target)
    loss.backward()
    grads = {name.replace(".", "__"): np_(p.grad) for name, p in model.named_parameters()
             if name.endswith("weight_scaling")}
    np.savez_compressed(os.path.join(HERE, "lsa_step.npz"), rays_o=np_(o), rays_d=np_(d), target=np_(target),
                        loss=np.float32(loss.item()), rgb=np_(rgb), rgb0=np_(ex["rgb0"]), **grads)
    for p in model.parameters():
        p.grad = None


def case_stepsize():
    """nnc_core/common.py:28-46 over the whole plausible qp range."""
    rows = []
    for dens in range(0, 5):
        for qp in range(-120, 41):
            rows.append((qp, dens, stepsize(qp, dens)))
    np.savez_compressed(os.path.join(HERE, "quant_stepsize.npz"), table=np.array(rows, dtype=np.float64))


if __name__ == "__main__":
    m = build_model()
    case_render_blender(m)
    case_render_ndc(m)
    case_render_perturb(m)
    case_functions()
    case_lsa_step(m)
    case_stepsize()
    case_render_trained()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
