"""Size-independent properties at BASELINE.json's full sizes (the oracle is too slow there): sharding / chunking
invariance of the render path, range and ordering invariants of its outputs, and quantiser round trips over the whole
wrapper for every qp of the sweep."""
import numpy as np
import pytest
import torch

from tests.util import synth_rays

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def wrapper(dev):
    from nerfq_b200 import codec, model as nmodel
    torch.manual_seed(0)
    w = nmodel.LSA(nmodel.NeRFWrapper()).add_lsa_params().to(dev)
    codec.quantize_model(w, -20)
    return w


def test_cfg2_size_shard_and_chunk_invariance(dev, wrapper):
    """4096 rays x (64+128) samples: rendering the batch at once, in chunks of 1000 rays (ragged last chunk) and as four
    independent shards gives the same pixels, bit for bit -- a point's result must not depend on its position in a
    256-point group, on the CTA that computes it, or on the order in which partial sums meet (the alpha head is reduced
    in fixed point with integer atomics for exactly this reason)."""
    from nerfq_b200 import render as R
    _, kw = R.create_nerf(wrapper, white_bkgd=True)
    r = synth_rays(4096, 2).to(dev)
    rays = (r[:, :3].contiguous(), r[:, 3:6].contiguous())
    with torch.no_grad():
        full = R.render(4, 4, None, chunk=32768, rays=rays, near=2.0, far=6.0, retraw=True, **kw)
        chunked = R.render(4, 4, None, chunk=1000, rays=rays, near=2.0, far=6.0, **kw)
        shards = [R.render(4, 4, None, chunk=32768, rays=(rays[0][i:i + 1024], rays[1][i:i + 1024]), near=2.0, far=6.0, **kw)
                  for i in range(0, 4096, 1024)]
    with torch.no_grad():
        again = R.render(4, 4, None, chunk=32768, rays=rays, near=2.0, far=6.0, retraw=True, **kw)
    assert torch.equal(full[3]["raw"], again[3]["raw"])
    for k in range(3):
        assert torch.equal(full[k], again[k])
        assert torch.equal(full[k], chunked[k])
        assert torch.equal(full[k], torch.cat([s[k] for s in shards], 0))
    # ... nor on how many CTAs share the work
    R.TUNING["max_ctas"] = 37
    try:
        with torch.no_grad():
            few = R.render(4, 4, None, chunk=32768, rays=rays, near=2.0, far=6.0, retraw=True, **kw)
    finally:
        R.TUNING["max_ctas"] = 0
    assert torch.equal(full[0], few[0]) and torch.equal(full[3]["raw"], few[3]["raw"])
    rgb, disp, acc, ex = full
    assert torch.isfinite(rgb).all() and float(rgb.min()) >= 0.0 and float(rgb.max()) <= 1.0 + 1e-5
    assert float(acc.min()) >= 0.0 and float(acc.max()) <= 1.0 + 1e-5
    assert (ex["z_std"] >= 0).all()
    assert ex["raw"].shape == (4096, 192, 4)


def test_cfg3_view_row_sharding(dev, wrapper):
    """800 x 800 view from a camera pose (cfg3 intrinsics): rows [200, 300) rendered as their own shard (what rank r of a
    row-sharded view does) equal the same rows of the full render."""
    from nerfq_b200 import render as R, ops
    _, kw = R.create_nerf(wrapper, white_bkgd=True)
    H = W = 800
    f = 0.5 * W / np.tan(0.5 * 0.6911112070083618)
    K = np.array([[f, 0, 0.5 * W], [0, f, 0.5 * H], [0, 0, 1]], dtype=np.float32)
    c2w = np.array([[1, 0, 0, 0.0], [0, 0.8660254, 0.5, 2.0], [0, -0.5, 0.8660254, 3.4641016]], dtype=np.float32)
    with torch.no_grad():
        rgb, disp, acc, _ = R.render(H, W, K, chunk=32768, c2w=torch.from_numpy(c2w), near=2.0, far=6.0, **kw)
        assert rgb.shape == (H, W, 3)
        rays_o, rays_d = R.get_rays(H, W, torch.from_numpy(K), torch.from_numpy(c2w))
        sl = slice(200, 300)
        part = R.render(H, W, K, chunk=32768, rays=(rays_o[sl].reshape(-1, 3).to(dev), rays_d[sl].reshape(-1, 3).to(dev)),
                        near=2.0, far=6.0, **kw)
    # the rays of the shard come from get_rays + pack_rays instead of the fused camera kernel: same values to 1 ulp
    assert torch.allclose(part[0].reshape(100, W, 3), rgb[sl], atol=1e-4, rtol=0)
    assert torch.allclose(part[2].reshape(100, W), acc[sl], atol=1e-4, rtol=0)


@pytest.mark.parametrize("qp", list(range(-38, -9, 4)))
def test_quantiser_round_trip_whole_wrapper(dev, qp):
    """Every weight tensor of a fresh wrapper at every qp of the sweep: |w - dequant(quant(w))| <= delta/2, and the levels
    are a fixed point (quantising the reconstruction returns the same levels)."""
    from nerfq_b200 import model as nmodel, ops
    torch.manual_seed(1)
    w = nmodel.NeRFWrapper().to(dev)
    delta = ops.stepsize(qp, 2)
    n = 0
    for name, p in w.state_dict().items():
        if not name.endswith(".weight"):
            continue
        lv, used = ops.quantize_urq(p, qp, 2)
        assert int(used.item()) == qp
        rec = ops.dequantize(lv, qp, 2)
        assert float((rec - p).abs().max()) <= 0.5 * delta * (1 + 1e-6)
        lv2, _ = ops.quantize_urq(rec, qp, 2)
        assert torch.equal(lv, lv2)
        n += p.numel()
    assert n == 2 * 593408


def test_cfg4_ndc_view_shard_invariance(dev, wrapper):
    """378 x 504 forward-facing view in NDC (cfg4: near 0, far 1, rays warped by ndc_rays inside the ray kernel): the
    second half of the image rendered as its own shard (first_pixel offset in the camera kernel) equals the same pixels
    of the full render, bit for bit; disparity and accumulation stay finite and in range."""
    from nerfq_b200 import ops, render as R
    _, kw = R.create_nerf(wrapper, white_bkgd=False, dataset_type="llff")
    H, W, focal = 378, 504, 407.5
    K = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]], dtype=np.float32)
    c2w = np.array([[1, 0, 0, 0.1], [0, 1, 0, -0.05], [0, 0, 1, 0.2]], dtype=np.float32)
    kw = dict(kw)
    for k in ("ndc", "near", "far"):
        kw.pop(k, None)
    with torch.no_grad():
        rgb, disp, acc, _ = R.render(H, W, K, chunk=32768, c2w=torch.from_numpy(c2w), ndc=True, near=0.0, far=1.0, **kw)
        first = (H // 2) * W
        rays = ops.camera_rays(H, W, K, c2w, True, 0.0, 1.0, dev, first_pixel=first)
        part = R.batchify_rays(rays, 32768, **{k: v for k, v in kw.items() if k not in ("use_viewdirs", "network_query_fn")})
    assert rgb.shape == (H, W, 3) and torch.isfinite(rgb).all() and torch.isfinite(acc).all()
    assert float(acc.min()) >= 0.0 and float(acc.max()) <= 1.0 + 1e-5
    assert torch.equal(part["rgb_map"], rgb.reshape(-1, 3)[first:])
    assert torch.equal(part["acc_map"], acc.reshape(-1)[first:])
