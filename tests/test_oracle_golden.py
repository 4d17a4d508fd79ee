"""Pin the oracle (oracle/render_oracle.py, oracle/quant_oracle.*) against the golden vectors
produced by the unmodified reference (tests/golden/make_golden.py)."""
import numpy as np
import torch

from oracle import quant_oracle, render_oracle as ro
from tests.util import golden, golden_model_params

TOL = 2e-5   # fp32 CPU vs fp32 CPU; slack only for BLAS kernel selection on a different host


def close(a, b, tol=TOL, nan_ok=False):
    a = np.asarray(a); b = np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    if nan_ok:
        assert (np.isnan(a) == np.isnan(b)).all()
        m = ~np.isnan(a)
        a, b = a[m], b[m]
    err = np.abs(a - b) / np.maximum(1.0, np.abs(b))
    assert err.max() <= tol, err.max()


def test_positional_encoding():
    g = golden("functions.npz")
    x = torch.from_numpy(g["pe_x"])
    close(ro.positional_encoding(x, 10), g["pe10"], 1e-6)
    close(ro.positional_encoding(x, 4), g["pe4"], 1e-6)


def test_composite():
    g = golden("functions.npz")
    raw, z, d = (torch.from_numpy(g[k]) for k in ("raw", "z", "d"))
    for wb in (0, 1):
        rgb, disp, acc, w, depth = ro.composite(raw, z, d, bool(wb))
        close(rgb, g[f"c_rgb_{wb}"], 1e-6)
        close(acc, g[f"c_acc_{wb}"], 1e-6)
        close(w, g[f"c_w_{wb}"], 1e-6)
        close(depth, g[f"c_depth_{wb}"], 1e-6)
        close(disp, g[f"c_disp_{wb}"], 1e-6, nan_ok=True)
    assert np.isnan(g["c_disp_0"][:5]).all()      # zero-density rays: the reference yields NaN


def test_importance_sample():
    g = golden("functions.npz")
    bins, wts = torch.from_numpy(g["bins"]), torch.from_numpy(g["wts"])
    close(ro.importance_sample(bins, wts, 128), g["pdf_det"], 1e-6)
    close(ro.importance_sample(bins, wts, 128, torch.from_numpy(g["pdf_u"])), g["pdf_rnd"], 1e-6)


def test_render_blender():
    g = golden("render_blender.npz")
    p, _ = golden_model_params()
    rays = (torch.from_numpy(g["rays_o"]), torch.from_numpy(g["rays_d"]))
    with torch.no_grad():
        rgb, disp, acc, ex = ro.render(p, 4, 4, None, chunk=64, rays=rays, ndc=False, near=2.0, far=6.0,
                                       white_bkgd=True, retraw=True)
    close(rgb, g["rgb"]); close(disp, g["disp"], nan_ok=True); close(acc, g["acc"])
    close(ex["rgb0"], g["rgb0"]); close(ex["disp0"], g["disp0"], nan_ok=True); close(ex["acc0"], g["acc0"])
    close(ex["z_std"], g["z_std"]); close(ex["raw"], g["raw"], 1e-4)


def test_render_ndc_and_raygen():
    g = golden("render_ndc.npz")
    p, _ = golden_model_params()
    H, W, K, c2w = int(g["H"]), int(g["W"]), g["K"], torch.from_numpy(g["c2w"])
    ro_, rd_ = ro.camera_rays(H, W, K, c2w)
    close(ro_, g["rays_o"], 1e-6); close(rd_, g["rays_d"], 1e-6)
    no, nd = ro.ndc_rays(H, W, K[0][0], 1.0, ro_, rd_)
    close(no, g["ndc_o"], 1e-6); close(nd, g["ndc_d"], 1e-6)
    with torch.no_grad():
        rgb, disp, acc, ex = ro.render(p, H, W, K, chunk=32, c2w=c2w, ndc=True, near=0.0, far=1.0,
                                       white_bkgd=False)
    close(rgb, g["rgb"]); close(disp, g["disp"], nan_ok=True); close(acc, g["acc"])
    close(ex["rgb0"], g["rgb0"]); close(ex["z_std"], g["z_std"])


def test_render_perturb_noise():
    g = golden("render_perturb.npz")
    p, _ = golden_model_params()
    t = {k: torch.from_numpy(g[k]) for k in ("ray_batch", "t_rand", "u", "noise0", "noise1")}
    with torch.no_grad():
        out = ro.render_rays(p, t["ray_batch"], white_bkgd=False, t_rand=t["t_rand"], u=t["u"],
                             noise0=t["noise0"], noise1=t["noise1"], retraw=True)
    for k in ("rgb_map", "acc_map", "rgb0", "acc0", "z_std"):
        close(out[k], g[k])
    close(out["disp_map"], g["disp_map"], nan_ok=True)
    close(out["raw"], g["raw"], 1e-4)


def test_lsa_gradients():
    g = golden("lsa_step.npz")
    p, _ = golden_model_params()
    o, d = torch.from_numpy(g["rays_o"]), torch.from_numpy(g["rays_d"])
    batch, _ = ro.pack_rays(4, 4, None, rays=(o, d), ndc=False, near=2.0, far=6.0)
    loss, grads, out = ro.lsa_scale_grads(p, batch, torch.from_numpy(g["target"]), white_bkgd=True)
    assert abs(loss - float(g["loss"])) < 1e-6
    for k, v in grads.items():
        ref = g[k.replace(".", "__")]
        scale = max(np.abs(ref).max(), 1e-12)
        assert np.abs(v.numpy() - ref).max() <= 1e-4 * scale + 1e-9, k


def test_stepsize_table():
    tab = golden("quant_stepsize.npz")["table"]
    for qp, dens, want in tab:
        assert quant_oracle.stepsize_py(int(qp), int(dens)) == want
        assert quant_oracle.stepsize(int(qp), int(dens)) == np.float32(want)


def test_quant_c_vs_numpy_and_roundtrip():
    rng = np.random.default_rng(0)
    for qp in (-38, -20, -15, -10, -75):
        w = (rng.standard_normal(4097) * 0.2).astype(np.float32)
        w[:3] = [0.0, -0.0, 1e-9]
        lv, used = quant_oracle.quant_urq(w, qp, 2)
        assert used == qp
        assert (lv == quant_oracle.quant_urq_np(w, qp, 2)).all()
        rec = quant_oracle.dequant(lv, qp, 2)
        d = quant_oracle.stepsize(qp, 2)
        assert np.abs(rec - w).max() <= 0.5 * d * (1 + 1e-5)
    big = np.array([3e4, -1.0], dtype=np.float32)       # |w|/delta(-75) > 2^31 -> qp is raised
    lv, used = quant_oracle.quant_urq(big, -75, 2)
    assert used > -75 and abs(int(lv[0])) < 2 ** 31


def test_render_trained_like():
    """The trained-like fixture (qp=-38, levels to 4267, both alpha variants) through the oracle restatement."""
    from tests.util import golden_trained_params
    g = golden("render_trained.npz")
    rays = (torch.from_numpy(g["rays_o"][:48]), torch.from_numpy(g["rays_d"][:48]))
    for variant in ("spread", "dense"):
        p, levels, _ = golden_trained_params(variant)
        assert max(int(np.abs(v).max()) for v in levels.values()) > 4000
        with torch.no_grad():
            rgb, disp, acc, ex = ro.render(p, 4, 4, None, chunk=48, rays=rays, ndc=False, near=2.0, far=6.0, white_bkgd=True, retraw=True)
        close(rgb, g[variant + "_rgb"][:48]); close(acc, g[variant + "_acc"][:48]); close(disp, g[variant + "_disp"][:48], nan_ok=True)
        close(ex["rgb0"], g[variant + "_rgb0"][:48]); close(ex["z_std"], g[variant + "_z_std"][:48])
        lo, hi = g[variant + "_sigma_minmax"]
        assert (hi > 100.0) == (variant == "dense")
