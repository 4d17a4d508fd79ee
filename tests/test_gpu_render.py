"""GPU parity tests, path level: the fused render path through the reference-shaped API against the golden
vectors produced by the unmodified reference and against the oracle.  Gate (north_star): rgb/disp/acc within
1e-3 absolute; the MLP runs with fp16 operands and fp32 accumulation."""
import numpy as np
import pytest
import torch

from tests.util import golden, synth_rays

pytestmark = pytest.mark.gpu
GATE = 1e-3


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def maxerr(a, b, nan_ok=False):
    a = a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    b = b.detach().cpu().numpy() if torch.is_tensor(b) else np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    if nan_ok:
        assert (np.isnan(a) == np.isnan(b)).all()
        m = ~np.isnan(a)
        a, b = a[m], b[m]
    return float(np.abs(a - b).max()) if a.size else 0.0


def disp_err(a, b):
    """disparity is 1/depth and unbounded (1e10 on empty rays): compare relative to max(1,|ref|)."""
    a = a.detach().cpu().numpy().reshape(-1); b = np.asarray(b).reshape(-1)
    assert (np.isnan(a) == np.isnan(b)).all()
    m = ~np.isnan(a)
    return float((np.abs(a[m] - b[m]) / np.maximum(1.0, np.abs(b[m]))).max())


@pytest.mark.parametrize("with_levels", [True, False])
def test_render_blender_golden(dev, with_levels):
    from nerfq_b200 import render as R
    from tests.gpu_util import golden_wrapper
    g = golden("render_blender.npz")
    w, _ = golden_wrapper(dev, with_levels)
    train_kw, test_kw = R.create_nerf(w, white_bkgd=True, dataset_type="blender")
    rays = (torch.from_numpy(g["rays_o"]).to(dev), torch.from_numpy(g["rays_d"]).to(dev))
    with torch.no_grad():
        rgb, disp, acc, ex = R.render(4, 4, None, chunk=64, rays=rays, near=2.0, far=6.0, retraw=True, **test_kw)
    assert maxerr(rgb, g["rgb"]) < GATE and maxerr(acc, g["acc"]) < GATE and disp_err(disp, g["disp"]) < GATE
    assert maxerr(ex["rgb0"], g["rgb0"]) < GATE and maxerr(ex["acc0"], g["acc0"]) < GATE and disp_err(ex["disp0"], g["disp0"]) < GATE
    assert maxerr(ex["z_std"], g["z_std"]) < GATE
    assert maxerr(ex["raw"], g["raw"]) < 2e-2        # raw logits are not gated; the fixture's alpha head is scaled x8


def test_render_ndc_golden(dev):
    from nerfq_b200 import render as R
    from tests.gpu_util import golden_wrapper
    g = golden("render_ndc.npz")
    w, _ = golden_wrapper(dev, True)
    _, test_kw = R.create_nerf(w, white_bkgd=False, dataset_type="llff")
    H, W, K = int(g["H"]), int(g["W"]), g["K"]
    with torch.no_grad():
        rgb, disp, acc, ex = R.render(H, W, K, chunk=32, c2w=torch.from_numpy(g["c2w"]), ndc=True, near=0.0, far=1.0, **test_kw)
    assert rgb.shape == (H, W, 3)
    assert maxerr(rgb, g["rgb"]) < GATE and maxerr(acc, g["acc"]) < GATE and disp_err(disp, g["disp"]) < GATE
    assert maxerr(ex["rgb0"], g["rgb0"]) < GATE and maxerr(ex["z_std"], g["z_std"]) < GATE


def test_render_rays_perturb_golden(dev):
    """perturb=1, raw_noise_std=1 with the reference's pytest=True RNG hooks."""
    from nerfq_b200 import render as R
    from tests.gpu_util import golden_wrapper
    g = golden("render_perturb.npz")
    w, _ = golden_wrapper(dev, True)
    train_kw, _ = R.create_nerf(w, perturb=1.0, raw_noise_std=1.0, white_bkgd=False)
    train_kw.pop("use_viewdirs"); train_kw.pop("ndc"); train_kw.pop("lindisp")
    with torch.no_grad():
        out = R.render_rays(torch.from_numpy(g["ray_batch"]).to(dev), pytest=True, retraw=True, **train_kw)
    for k in ("rgb_map", "acc_map", "rgb0", "acc0", "z_std"):
        assert maxerr(out[k], g[k]) < GATE, k
    assert disp_err(out["disp_map"], g["disp_map"]) < GATE


def test_render_rays_oracle_cfg1(dev):
    """BASELINE configs[0]: 1024 synthetic rays, 64+128 samples, against the oracle."""
    from nerfq_b200 import render as R
    from oracle import render_oracle as ro
    from tests.gpu_util import golden_wrapper
    w, p = golden_wrapper(dev, True)
    _, test_kw = R.create_nerf(w, white_bkgd=True)
    test_kw.pop("use_viewdirs"); test_kw.pop("ndc"); test_kw.pop("lindisp")
    batch = synth_rays(1024, 1)
    with torch.no_grad():
        ref = ro.render_rays(p, batch, white_bkgd=True)
        out = R.render_rays(batch.to(dev), **test_kw)
    for k in ("rgb_map", "acc_map", "rgb0", "acc0", "z_std"):
        assert maxerr(out[k], ref[k]) < GATE, k
    assert disp_err(out["disp_map"], ref["disp_map"].numpy()) < GATE
    mse = float(((out["rgb_map"].cpu() - ref["rgb_map"]) ** 2).mean())
    assert mse < 1e-8                               # PSNR between the two renders > 80 dB (gate: 0.05 dB on test-view PSNR)


def test_run_network_and_edge_cases(dev):
    from nerfq_b200 import render as R
    from oracle import render_oracle as ro
    from tests.gpu_util import golden_wrapper
    w, p = golden_wrapper(dev, False)
    gen = torch.Generator().manual_seed(3)
    pts = torch.randn(5, 7, 3, generator=gen) * 2
    vd = torch.nn.functional.normalize(torch.randn(5, 3, generator=gen), dim=-1)
    raw = R.run_network(pts.to(dev), vd.to(dev), w.model_fine)
    with torch.no_grad():
        ref = ro.query_network(p, "model_fine", pts, vd)
    assert maxerr(raw, ref) < 5e-3
    _, test_kw = R.create_nerf(w, white_bkgd=True)
    test_kw.pop("use_viewdirs"); test_kw.pop("ndc"); test_kw.pop("lindisp")
    for n in (1, 3, 129):                      # ragged tiles: 1 ray = half a tile, 129 rays = odd tile count
        batch = synth_rays(n, 20 + n)
        with torch.no_grad():
            ref = ro.render_rays(p, batch, white_bkgd=True)
            out = R.render_rays(batch.to(dev), **test_kw)
        assert maxerr(out["rgb_map"], ref["rgb_map"]) < GATE and maxerr(out["acc_map"], ref["acc_map"]) < GATE
    out = R.render_rays(torch.zeros(0, 11, device=dev), **test_kw)      # empty batch
    assert out["rgb_map"].shape == (0, 3)


def test_apply_lsa_decoder_side_matches_lsa_model(dev):
    """codec.apply_lsa (approximator/__init__.py:276-318: w *= ls, scales dropped) yields a plain wrapper whose render
    equals the LSA model's, where the fused MLP applies delta*ls in its epilogue on integer-level operands: the two differ
    only by the fp16 rounding of the folded weights."""
    from nerfq_b200 import codec, model as nmodel, render as R
    torch.manual_seed(4)
    w = nmodel.LSA(nmodel.NeRFWrapper()).add_lsa_params().to(dev)
    with torch.no_grad():
        for n, p in w.named_parameters():
            if n.endswith("weight_scaling"):
                p.copy_(1.0 + 0.05 * torch.randn_like(p))            # scales far enough from 1 to matter
    codec.quantize_model(w, -20)
    plain = codec.apply_lsa(w)
    assert not any(k.endswith("weight_scaling") for k in plain.state_dict())
    assert len(plain.state_dict()) == 48
    sd, ps = w.state_dict(), plain.state_dict()
    k = "model_fine.pts_linears.3"
    assert torch.equal(ps[k + ".weight"], sd[k + ".weight"] * sd[k + ".weight_scaling"])
    r = synth_rays(300, 9).to(dev)
    rays = (r[:, :3].contiguous(), r[:, 3:6].contiguous())
    _, kw_lsa = R.create_nerf(w, white_bkgd=True)
    _, kw_plain = R.create_nerf(plain, white_bkgd=True)
    with torch.no_grad():
        a = R.render(4, 4, None, rays=rays, near=2.0, far=6.0, **kw_lsa)
        b = R.render(4, 4, None, rays=rays, near=2.0, far=6.0, **kw_plain)
    assert maxerr(a[0], b[0]) < GATE and maxerr(a[2], b[2]) < GATE
