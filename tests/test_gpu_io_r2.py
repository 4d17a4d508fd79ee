"""GPU tests of the pieces either side of the render path added in round 2: the one-call forward render entry,
the image output path (device to8b, asynchronous double-buffered copies), device-side training-batch selection, and loading
decoded levels straight into the packed networks."""
import numpy as np
import pytest
import torch

from tests.test_gpu_render import GATE, maxerr
from tests.util import LAYERS, NETS, synth_rays

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _pose(theta=0.3):
    c, s = np.cos(theta), np.sin(theta)
    return np.array([[c, 0, s, 4 * s], [0, 1, 0, 0.2], [-s, 0, c, 4 * c]], dtype=np.float32)


def test_one_call_render_equals_the_staged_path(dev):
    """nerfq_render_rays_fwd (one C-ABI call, what render_rays uses without gradients / RNG) against the six separate
    entry points: same kernels, same order -- identical bits; also N_importance = 0 and the shared-network fine pass."""
    from nerfq_b200 import ops, render as R
    from tests.gpu_util import golden_wrapper
    w, _ = golden_wrapper(dev, True)
    rays = synth_rays(777, 4).to(dev)
    sc0, sc1 = R._scale_flat(w.model), R._scale_flat(w.model_fine)
    for ni, fine in ((128, w.model_fine), (64, None), (0, None)):
        cfg = R._make_cfg(rays, w.model, fine, 64, ni, False, 0., True, 0., False)
        with torch.no_grad():
            outs, _ = R._forward_pipeline(cfg, sc0, sc1 if fine is not None else None, save=False)
            pn0 = R._refresh(w.model, sc0)
            pn1 = R._refresh(fine, sc1) if fine is not None else None
            got = ops.render_rays_fwd(pn0, pn1, rays, 64, ni, False, True)
        if ni > 0:
            for a, b in zip(got, outs[:7]):
                assert torch.equal(a, b)
        else:
            for a, b in zip(got[:3], outs[:3]):
                assert torch.equal(a, b)
            assert got[3] is None
    _, kw = R.create_nerf(w, white_bkgd=True)
    kw.pop("use_viewdirs"); kw.pop("ndc"); kw.pop("lindisp")
    with torch.no_grad():
        a = R.render_rays(rays, **kw)                    # one-call path
        b = R.render_rays(rays, retraw=True, **kw)       # staged path (returns raw)
    for k in ("rgb_map", "disp_map", "acc_map", "rgb0", "z_std"):
        assert torch.equal(a[k], b[k]) or (torch.isnan(a[k]) == torch.isnan(b[k])).all()
    assert R.render_rays(torch.zeros(0, 11, device=dev), **kw)["rgb_map"].shape == (0, 3)


def test_to8b_matches_numpy(dev):
    from nerfq_b200 import ops, render as R
    gen = torch.Generator().manual_seed(0)
    for n in (1, 3, 4, 1001, 378 * 504 * 3):
        x = torch.rand(n, generator=gen) * 1.4 - 0.2
        x[: min(n, 3)] = torch.tensor([0.0, 1.0, 0.5])[: min(n, 3)]
        got = ops.to8b(x.to(dev)).cpu().numpy()
        assert (got == R.to8b(x.numpy())).all()
    y = torch.rand(5, 7, 3, generator=gen).to(dev)[:, 1:, :].contiguous()       # odd sizes, offset storage
    assert (ops.to8b(y).cpu().numpy() == R.to8b(y.cpu().numpy())).all()


def test_select_batch_is_a_permutation_with_matching_rays(dev):
    from nerfq_b200 import ops
    H, W = 50, 37
    K = np.array([[40.0, 0, W / 2], [0, 40.0, H / 2], [0, 0, 1]], dtype=np.float32)
    c2w = _pose()
    img = torch.rand(H, W, 3, generator=torch.Generator().manual_seed(1)).to(dev)
    full = ops.camera_rays(H, W, K, c2w, False, 2.0, 6.0, dev)
    rays, tgt, idx = ops.select_batch(H, W, K, c2w, img, H * W, seed=7, step=0, near=2.0, far=6.0, want_index=True)
    i = idx.cpu().numpy()
    assert sorted(i.tolist()) == list(range(H * W))                 # every pixel exactly once: a permutation
    assert not (i == np.arange(H * W)).all()
    assert torch.equal(rays, full[idx.long()]) and torch.equal(tgt, img.reshape(-1, 3)[idx.long()])
    a = ops.select_batch(H, W, K, c2w, img, 512, seed=7, step=1, near=2.0, far=6.0, want_index=True)[2].cpu().numpy()
    b = ops.select_batch(H, W, K, c2w, img, 512, seed=7, step=2, near=2.0, far=6.0, want_index=True)[2].cpu().numpy()
    a2 = ops.select_batch(H, W, K, c2w, img, 512, seed=7, step=1, near=2.0, far=6.0, want_index=True)[2].cpu().numpy()
    assert len(set(a.tolist())) == 512 and (a == a2).all() and not (a == b).all()
    # roughly uniform over the image: mean index of a 512-pixel draw within 10 % of the centre
    assert abs(a.mean() / (H * W) - 0.5) < 0.1
    # NDC variant (LLFF): rays equal nerfq_camera_rays with ndc
    fulln = ops.camera_rays(H, W, K, c2w, True, 0.0, 1.0, dev)
    r2, _, i2 = ops.select_batch(H, W, K, c2w, None, 100, seed=1, step=5, ndc=True, device=dev, want_index=True)
    assert torch.equal(r2, fulln[i2.long()])


def test_lsa_step_with_device_batch_selection(dev):
    from nerfq_b200 import codec, lsa, model as nmodel
    torch.manual_seed(0)
    w = nmodel.LSA(nmodel.NeRFWrapper()).add_lsa_params().to(dev)
    codec.quantize_model(w, -20)
    H = W = 64
    K = np.array([[55.0, 0, W / 2], [0, 55.0, H / 2], [0, 0, 1]], dtype=np.float32)
    img = torch.rand(H, W, 3, generator=torch.Generator().manual_seed(2)).to(dev)
    step = lsa.LSAStep(w, 256, lr=1e-3, perturb=0.0, white_bkgd=True)
    before = step.flat.clone()
    l0 = float(step.step_selected(img, H, W, K, _pose(), seed=3, step=0))
    step.capture(warmup=0)
    l1 = float(step.step_selected(img, H, W, K, _pose(), seed=3, step=1))
    assert np.isfinite(l0) and np.isfinite(l1) and l0 != l1 and not torch.equal(before, step.flat)


def test_render_path_outputs(dev, tmp_path):
    """render_path (run_nerf.py:161-211): float arrays per view equal render() of that pose; the saved 8-bit frames equal
    to8b of them; render_path_8bit delivers the same frames through its sink, shard by shard."""
    from nerfq_b200 import render as R
    from tests.gpu_util import golden_wrapper
    w, _ = golden_wrapper(dev, True)
    _, kw = R.create_nerf(w, white_bkgd=True)
    H, W, focal = 24, 32, 30.0
    K = np.array([[focal, 0, W / 2], [0, focal, H / 2], [0, 0, 1]], dtype=np.float32)
    poses = [torch.from_numpy(np.vstack([_pose(t), [[0, 0, 0, 1]]]).astype(np.float32)) for t in (0.0, 0.4, 0.8, 1.2, 1.6)]
    kw_r = dict(kw, near=2.0, far=6.0)
    rgbs, disps = R.render_path(poses, (H, W, focal), K, 300, kw_r, savedir=str(tmp_path))
    assert rgbs.shape == (5, H, W, 3) and disps.shape == (5, H, W)
    frames = {}
    n = R.render_path_8bit(poses, (H, W, focal), K, 300, kw_r, sink=lambda i, im: frames.__setitem__(i, im.copy()), first_view=1, view_count=3)
    assert n == 3 and sorted(frames) == [1, 2, 3]
    for i, p in enumerate(poses):
        with torch.no_grad():
            rgb, disp, _, _ = R.render(H, W, K, chunk=300, c2w=p[:3, :4], **kw_r)
        assert np.array_equal(rgbs[i], rgb.cpu().numpy()) and np.array_equal(disps[i], disp.cpu().numpy(), equal_nan=True)
        saved = np.load(str(tmp_path / f"{i:03d}.npy")) if (tmp_path / f"{i:03d}.npy").exists() else None
        if saved is not None:
            assert np.array_equal(saved, R.to8b(rgbs[i]))
        if i in frames:
            assert np.array_equal(frames[i], R.to8b(rgbs[i]))


def test_load_levels_matches_apply_lsa(dev):
    """codec.load_levels (decoded integer levels + qps -> packed networks, no float weights in between) renders like the
    reference's decoder output (rec + apply_lsa: float weights level * delta * ls)."""
    from nerfq_b200 import codec, model as nmodel, ops, render as R
    torch.manual_seed(6)
    src = nmodel.LSA(nmodel.NeRFWrapper()).add_lsa_params().to(dev)
    with torch.no_grad():
        for n, p in src.named_parameters():
            if n.endswith("weight_scaling"):
                p.copy_(1.0 + 0.05 * torch.randn_like(p))
    qp = -24
    levels, qps = {}, {}
    for k, v in src.state_dict().items():
        q = qp if k.endswith(".weight") else -75
        lv, used = ops.quantize_urq(v.detach().float().contiguous(), q, 2)
        levels[k], qps[k] = lv.cpu().numpy(), int(used.item())
    dst = nmodel.LSA(nmodel.NeRFWrapper()).add_lsa_params().to(dev)
    codec.load_levels(dst, levels, qps)
    assert dst.model_fine.quant_levels is not None
    sd = dst.state_dict()
    for k in sd:
        assert torch.equal(sd[k].reshape(-1), ops.dequantize(torch.from_numpy(levels[k]).to(dev), qps[k], 2).reshape(-1)), k
    plain = codec.apply_lsa(dst)                                       # what the reference's decompress_model hands out
    r = synth_rays(300, 12).to(dev)
    rays = (r[:, :3].contiguous(), r[:, 3:6].contiguous())
    _, kw_a = R.create_nerf(dst, white_bkgd=True)
    _, kw_b = R.create_nerf(plain, white_bkgd=True)
    with torch.no_grad():
        a = R.render(4, 4, None, rays=rays, near=2.0, far=6.0, **kw_a)
        b = R.render(4, 4, None, rays=rays, near=2.0, far=6.0, **kw_b)
    assert maxerr(a[0], b[0]) < GATE and maxerr(a[2], b[2]) < GATE


def test_peer_finalize_single_rank_equals_the_plain_finalize(dev):
    """nerfq_mlp_backward_finalize_peers with a world of one rank (its own region is the only peer) must produce exactly what
    the two plain finalize launches produce, leave the accumulators zeroed, and keep doing so across epochs (the staging
    area alternates with the epoch parity).  The multi-rank behaviour -- bit-identical to one GPU on the concatenated
    batch -- is asserted on hardware by bench.py --gpus N (`dp_parity`) and profiles/dp_parity_check.py."""
    from nerfq_b200 import ops
    from tests.gpu_util import golden_wrapper
    w, _ = golden_wrapper(dev, True)
    pn0, pn1 = w.model.packed_net(), w.model_fine.packed_net()
    pn0.set_scales(w.model.scale_tensors())
    pn1.set_scales(w.model_fine.scale_tensors())
    n = ops.grad_fix_elems()
    region = torch.zeros(ops.dp_peer_bytes(), dtype=torch.uint8, device=dev)
    peers = torch.tensor([region.data_ptr()], dtype=torch.int64, device=dev)
    epoch = torch.zeros(1, dtype=torch.int32, device=dev)
    gen = torch.Generator().manual_seed(7)
    for it in range(3):
        fix = (torch.randint(-2**40, 2**40, (2, n), generator=gen, dtype=torch.int64)).to(dev)
        fix[:, 2436:] = 0
        ref = torch.zeros((2, 2436), device=dev)
        f0, f1 = fix[0].clone(), fix[1].clone()
        ops.mlp_backward_finalize(pn0, f0, ref[0])
        ops.mlp_backward_finalize(pn1, f1, ref[1])
        got = torch.zeros((2, 2436), device=dev)
        ops.mlp_backward_finalize_peers(pn0, pn1, fix, int(peers.data_ptr()), 1, 0, epoch, got)
        torch.cuda.synchronize()
        assert torch.equal(got, ref) and float(ref.abs().max()) > 0
        assert int(fix.abs().max()) == 0 and int(epoch.item()) == it + 1
