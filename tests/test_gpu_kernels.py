"""GPU parity tests, function level: every HBM-bound kernel and the quantiser against the golden vectors
of the reference and against the oracle on seeded inputs.  All calls go through the C ABI (ops.py)."""
import numpy as np
import pytest
import torch

from tests.util import golden, synth_rays

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _ops():
    import nerfq_b200  # noqa: F401
    from nerfq_b200 import ops
    return ops


def close(a, b, tol, nan_ok=False):
    a = a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    b = b.detach().cpu().numpy() if torch.is_tensor(b) else np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    if nan_ok:
        assert (np.isnan(a) == np.isnan(b)).all()
        m = ~np.isnan(a)
        a, b = a[m], b[m]
    err = np.abs(a - b) / np.maximum(1.0, np.abs(b))
    assert err.max() <= tol, float(err.max())


def test_composite_fwd_golden(dev):
    ops = _ops()
    g = golden("functions.npz")
    raw, z, d = (torch.from_numpy(g[k]).to(dev) for k in ("raw", "z", "d"))
    rays = torch.zeros(raw.shape[0], 11, device=dev)
    rays[:, 3:6] = d
    for wb in (0, 1):
        rgb, disp, acc, w, depth = ops.composite_fwd(raw, z, rays, bool(wb))
        close(rgb, g[f"c_rgb_{wb}"], 2e-6); close(acc, g[f"c_acc_{wb}"], 2e-6); close(w, g[f"c_w_{wb}"], 2e-6)
        close(depth, g[f"c_depth_{wb}"], 2e-6); close(disp, g[f"c_disp_{wb}"], 2e-6, nan_ok=True)


@pytest.mark.parametrize("S", [64, 192, 33, 1, 256])
def test_composite_fwd_bwd_oracle(dev, S):
    ops = _ops()
    from oracle import render_oracle as ro
    gen = torch.Generator().manual_seed(S)
    n = 257
    raw = (torch.randn(n, S, 4, generator=gen) * 1.5)
    z = torch.sort(2 + 4 * torch.rand(n, S, generator=gen), -1).values
    d = torch.randn(n, 3, generator=gen)
    noise = torch.rand(n, S, generator=gen)
    rays = torch.zeros(n, 11)
    rays[:, 3:6] = d
    for wb in (False, True):
        for nz in (None, noise):
            raw_r = raw.clone().requires_grad_(True)
            rgb_r, disp_r, acc_r, w_r, depth_r = ro.composite(raw_r, z, d, wb, nz)
            gout = torch.randn(n, 3, generator=gen)
            (rgb_r * gout).sum().backward()
            rgb, disp, acc, w, depth = ops.composite_fwd(raw.to(dev), z.to(dev), rays.to(dev), wb, None if nz is None else nz.to(dev))
            close(rgb, rgb_r, 3e-6); close(acc, acc_r, 3e-6); close(w, w_r, 3e-6); close(depth, depth_r, 3e-6)
            close(disp, disp_r, 3e-6, nan_ok=True)
            d_raw = ops.composite_bwd(raw.to(dev), z.to(dev), rays.to(dev), wb, gout.to(dev), None if nz is None else nz.to(dev))
            ref = raw_r.grad
            err = (d_raw.cpu() - ref).abs().max().item()
            assert err <= 2e-5 * max(1.0, ref.abs().max().item()), err


def test_sample_pdf_golden(dev):
    ops = _ops()
    g = golden("functions.npz")
    bins, wts = torch.from_numpy(g["bins"]).to(dev), torch.from_numpy(g["wts"]).to(dev)
    # deterministic u: ATen's vectorised CPU linspace (golden) and its CUDA linspace (what the kernel
    # restates) differ in the last bit of some u, which a flat cdf amplifies -> 2e-4
    close(ops.sample_pdf(bins, wts, 128), g["pdf_det"], 2e-4)
    close(ops.sample_pdf(bins, wts, 128, torch.from_numpy(g["pdf_u"]).to(dev)), g["pdf_rnd"], 2e-5)


@pytest.mark.parametrize("S,Ni", [(64, 128), (64, 64), (17, 5), (128, 128), (24, 40), (30, 33), (64, 100)])      # 1, 2, 4, 8 values per lane in the register sort; sizes that are not multiples of 4
def test_sample_fine_oracle(dev, S, Ni):
    ops = _ops()
    from oracle import render_oracle as ro
    gen = torch.Generator().manual_seed(S * 1000 + Ni)
    n = 131
    z = torch.sort(2 + 4 * torch.rand(n, S, generator=gen), -1).values
    w = torch.rand(n, S, generator=gen) ** 3
    w[3] = 0.0
    for u in (None, torch.rand(n, Ni, generator=gen)):
        mid = 0.5 * (z[:, 1:] + z[:, :-1])
        zs_r = ro.importance_sample(mid, w[:, 1:-1], Ni, u)
        z_all_r = torch.sort(torch.cat([z, zs_r], -1), -1).values
        z_all, z_std, zs = ops.sample_fine(z.to(dev), w.to(dev), Ni, None if u is None else u.to(dev), want_samples=True)
        # Knife edge: a bin whose pdf is within rounding of 1e-5 flips between the lerp
        # and the denom=1 branch with the last bit of the cdf; affected samples move by at most one bin.
        tol = 2e-4 if u is None else 2e-5
        rel = ((zs.cpu() - zs_r).abs() / zs_r.abs().clamp(min=1.0))
        assert (rel > tol).float().mean().item() < 0.01, rel.max().item()
        assert rel.max().item() < 2.5 * (4.0 / (S - 1)) / 2.0
        rel = ((z_all.cpu() - z_all_r).abs() / z_all_r.abs().clamp(min=1.0))
        assert (rel > tol).float().mean().item() < 0.01
        clean = (((zs.cpu() - zs_r).abs() / zs_r.abs().clamp(min=1.0)) <= tol).all(dim=-1)    # rows without a knife-edge flip
        close(z_std.cpu()[clean], torch.std(zs_r, dim=-1, unbiased=False)[clean], 1e-4)
        assert clean.float().mean().item() > 0.8
        assert (z_all[:, 1:] >= z_all[:, :-1]).all()
        assert torch.equal(torch.sort(torch.cat([z.to(dev), zs], -1), -1).values, z_all)   # merge == sort of the union


def test_coarse_depths(dev):
    ops = _ops()
    from oracle import render_oracle as ro
    gen = torch.Generator().manual_seed(5)
    n = 100
    rays = torch.randn(n, 11, generator=gen)
    rays[:, 6] = 0.5 + torch.rand(n, generator=gen)
    rays[:, 7] = 3.0 + torch.rand(n, generator=gen)
    t_rand = torch.rand(n, 64, generator=gen)
    for lindisp in (False, True):
        for tr in (None, t_rand):
            ref = ro.coarse_depths(rays[:, 6:7], rays[:, 7:8], 64, lindisp, tr)
            got = ops.coarse_depths(rays.to(dev), 64, lindisp, None if tr is None else tr.to(dev))
            close(got, ref, 1e-6)


def test_ray_generation_golden(dev):
    ops = _ops()
    g = golden("render_ndc.npz")
    H, W, K, c2w = int(g["H"]), int(g["W"]), g["K"], g["c2w"]
    rays = ops.camera_rays(H, W, K, c2w, False, 2.0, 6.0, dev)
    close(rays[:, 0:3].reshape(H, W, 3), g["rays_o"], 1e-6); close(rays[:, 3:6].reshape(H, W, 3), g["rays_d"], 1e-6)
    vd = g["rays_d"] / np.linalg.norm(g["rays_d"], axis=-1, keepdims=True)
    close(rays[:, 8:11].reshape(H, W, 3), vd, 1e-6)
    assert (rays[:, 6] == 2.0).all() and (rays[:, 7] == 6.0).all()
    rays = ops.camera_rays(H, W, K, c2w, True, 0.0, 1.0, dev)
    close(rays[:, 0:3].reshape(H, W, 3), g["ndc_o"], 2e-6); close(rays[:, 3:6].reshape(H, W, 3), g["ndc_d"], 2e-6)
    close(rays[:, 8:11].reshape(H, W, 3), vd, 1e-6)
    # ragged pixel range
    part = ops.camera_rays(H, W, K, c2w, True, 0.0, 1.0, dev, first_pixel=5, count=17)
    assert (part == rays[5:22]).all()
    packed = ops.pack_rays(torch.from_numpy(g["rays_o"]).to(dev), torch.from_numpy(g["rays_d"]).to(dev), True, H, W, float(K[0][0]), 0.0, 1.0)
    close(packed, rays, 1e-6)


def test_mse_grad(dev):
    ops = _ops()
    gen = torch.Generator().manual_seed(0)
    n = 1000
    rgb, rgb0, t = (torch.rand(n, 3, generator=gen) for _ in range(3))
    loss2, d1, d0 = ops.mse_grad(rgb.to(dev), rgb0.to(dev), t.to(dev))
    close(loss2, torch.stack([((rgb - t) ** 2).mean(), ((rgb0 - t) ** 2).mean()]), 1e-5)
    close(d1, 2 * (rgb - t) / (3 * n), 1e-7); close(d0, 2 * (rgb0 - t) / (3 * n), 1e-7)


def test_quantizer_bit_exact(dev):
    """Levels from the GPU kernel == levels from the C oracle for every qp of the BASELINE sweep, every tensor
    shape of the network, plus dequantisation and the step-size table."""
    ops = _ops()
    from oracle import quant_oracle as qo
    tab = golden("quant_stepsize.npz")["table"]
    for qp, dens, want in tab[::7]:
        assert ops.stepsize(int(qp), int(dens)) == np.float32(want)
    rng = np.random.default_rng(1)
    shapes = [(256, 63), (256, 256), (256, 319), (1, 256), (128, 283), (3, 128), (256,), (1,), (3,), (1000003,)]
    for qp in list(range(-38, -9, 4)) + [-75]:
        for shp in shapes:
            w = (rng.standard_normal(shp) * 0.15).astype(np.float32)
            if w.size > 2:
                w.flat[0], w.flat[1], w.flat[2] = 0.0, -0.0, np.float32(qo.stepsize(qp, 2)) * 1.5
            lv_ref, used_ref = qo.quant_urq(w, qp, 2)
            lv, used = ops.quantize_urq(torch.from_numpy(w).to(dev), qp, 2)
            assert int(used.item()) == used_ref == qp
            assert (lv.cpu().numpy() == lv_ref).all(), (qp, shp)
            rec = ops.dequantize(lv, qp, 2)
            assert (rec.cpu().numpy() == qo.dequant(lv_ref, qp, 2)).all()
    big = np.array([3e4, -1.0, 0.3], dtype=np.float32)        # forces the qp clip
    lv_ref, used_ref = qo.quant_urq(big, -75, 2)
    lv, used = ops.quantize_urq(torch.from_numpy(big).to(dev), -75, 2)
    assert int(used.item()) == used_ref > -75 and (lv.cpu().numpy() == lv_ref).all()
    e = torch.empty(0, dtype=torch.float32, device=dev)        # empty tensor
    lv, _ = ops.quantize_urq(e, -20, 2)
    assert lv.numel() == 0


def test_quantizer_batch_bit_exact(dev):
    """The batched launch (every tensor of a model at once, per-tensor qp, reconstruction in place) gives the same
    levels, clipped qps and reconstructed values as the C oracle tensor by tensor."""
    ops = _ops()
    from oracle import quant_oracle as qo
    rng = np.random.default_rng(7)
    shapes = [(256, 63), (256,), (256, 256), (256,), (256, 319), (1, 256), (1,), (128, 283), (3, 128), (3,), (70001,)]
    for qp in (-38, -20, -10):
        ws = [(rng.standard_normal(s) * 0.2).astype(np.float32) for s in shapes] + [np.array([3e4, -1.0, 0.3], dtype=np.float32)]
        qps = [qp if w.ndim == 2 else -75 for w in ws]
        ts = [torch.from_numpy(w.copy()).to(dev) for w in ws]
        lv, used = ops.quantize_batch(ts, qps, 2, reconstruct_in_place=True)
        used = used.cpu().numpy()
        for i, w in enumerate(ws):
            lv_ref, used_ref = qo.quant_urq(w, qps[i], 2)
            assert int(used[i]) == used_ref, (i, qp)
            assert (lv[i].cpu().numpy() == lv_ref).all(), (i, qp)
            assert (ts[i].cpu().numpy() == qo.dequant(lv_ref, used_ref, 2)).all(), (i, qp)
    assert int(used[-1]) > -75                                  # the last tensor forces the qp clip


def test_deepcabac_front_end_matches_oracle(dev):
    """The calls nnc_core/approximator/baseline.py makes (approx :39-62, rec :89-98), through nerfq_b200.deepcabac with
    numpy arrays allocated the way the reference allocates them, against the C oracle: levels, clipped qp, values."""
    from nerfq_b200 import deepcabac
    from oracle import quant_oracle as qo
    rng = np.random.default_rng(11)
    params = {"w": (rng.standard_normal((256, 319)) * 0.1).astype(np.float32), "b": (rng.standard_normal((256,)) * 0.01).astype(np.float32),
              "ls": (1.0 + 1e-3 * rng.standard_normal((256, 1))).astype(np.float32), "clip": np.array([[3e4, -1.0, 0.3]], dtype=np.float32)}
    qps = {"w": -20, "b": -75, "ls": -75, "clip": -75}
    encoder = deepcabac.Encoder()
    for name, values in params.items():
        quantized = np.zeros(values.shape, dtype=np.int32)
        encoder.initCtxModels(10, 0)
        qp = encoder.quantLayer(values, quantized, 0, 2, qps[name], 0.0, 10, 0)
        lv_ref, used_ref = qo.quant_urq(values, qps[name], 2)
        assert qp == used_ref and (quantized == lv_ref).all(), name
        assert (qp != qps[name]) == (name == "clip")
        rec = np.zeros(values.shape, dtype=np.float32)
        deepcabac.Decoder().dequantLayer(rec, quantized, 2, qp, 0)
        assert (rec == qo.dequant(lv_ref, used_ref, 2)).all(), name
        assert np.abs(rec - values).max() <= 0.5 * qo.stepsize(qp, 2) * (1 + 1e-6)


def test_c_abi_error_codes_and_empty_inputs(dev):
    """include/nerfq.h contract, called through ctypes exactly as a foreign host would: 0 on success and on empty inputs
    (nothing launched, pointers not dereferenced), -1 on bad arguments; no exception crosses the ABI."""
    import ctypes as C
    ops = _ops()
    L = ops.L()
    null, stream = C.c_void_p(0), C.c_void_p(torch.cuda.current_stream().cuda_stream)
    buf = torch.zeros(64, dtype=torch.float32, device=dev)
    ibuf = torch.zeros(64, dtype=torch.int32, device=dev)
    p = lambda t: C.c_void_p(t.data_ptr())
    # empty inputs
    assert L.nerfq_quantize_urq(null, null, C.c_longlong(0), -20, 2, null, null, stream) == 0
    assert L.nerfq_dequantize(null, null, C.c_longlong(0), -20, 2, stream) == 0
    assert L.nerfq_mlp_forward(null, null, null, C.c_longlong(0), 64, null, null, 0, stream) == 0
    assert L.nerfq_mlp_backward(null, null, null, null, C.c_longlong(0), null, 0, stream) == 0
    assert int(L.nerfq_mlp_save_bytes(C.c_longlong(0))) == 0
    # bad arguments
    assert L.nerfq_quantize_urq(null, p(ibuf), C.c_longlong(8), -20, 2, null, p(ibuf), stream) == -1          # no input
    assert L.nerfq_quantize_urq(p(buf), p(ibuf), C.c_longlong(-1), -20, 2, null, p(ibuf), stream) == -1      # negative size
    assert L.nerfq_quantize_urq(p(buf), p(ibuf), C.c_longlong(8), -20, 9, null, p(ibuf), stream) == -1       # qp_density out of range
    assert L.nerfq_dequantize(p(ibuf), null, C.c_longlong(8), -20, 2, stream) == -1
    assert L.nerfq_mlp_forward(null, p(buf), p(buf), C.c_longlong(4), 64, p(buf), null, 0, stream) == -1      # no packed network
    assert L.nerfq_mlp_forward(p(buf), p(buf), p(buf), C.c_longlong(4), 0, p(buf), null, 0, stream) == -1     # samples_per_ray <= 0
    assert L.nerfq_mlp_backward(p(buf), null, p(buf), p(buf), C.c_longlong(4), p(buf), 0, stream) == -1
    out = C.c_float()
    assert L.nerfq_stepsize(-20, 2, C.byref(out)) == 0 and out.value == 2.0 ** -5
    assert L.nerfq_stepsize(-20, 9, C.byref(out)) == -1
    torch.cuda.synchronize()
    # a save buffer sized by the library covers whole 256-point groups, ten 128 KB slots each
    assert int(L.nerfq_mlp_save_bytes(C.c_longlong(1))) == 10 * 131072
    assert int(L.nerfq_mlp_save_bytes(C.c_longlong(257))) == 2 * 10 * 131072


def test_mlp_kernels_stay_inside_their_buffers(dev):
    """Ragged sizes (37 rays x 33 samples = 1221 points = 4.77 groups of 256) through the C ABI with sentinel-filled
    guard regions behind every output buffer: raw, the saved activations and d_scale are written only inside their
    documented extents, and the ragged tail computes the same values as a padded run."""
    import ctypes as C
    from nerfq_b200 import codec, model as nmodel, packed
    ops = _ops()
    L = ops.L()
    torch.manual_seed(5)
    w = nmodel.LSA(nmodel.NeRFWrapper()).add_lsa_params().to(dev)
    codec.quantize_model(w, -20)
    pn = w.model_fine.packed_net()
    pn.set_scales(w.model_fine.scale_tensors())
    n, S = 37, 33
    r = synth_rays(n, 31).to(dev)
    z = torch.sort(2.0 + 4.0 * torch.rand(n, S, device=dev), -1).values.contiguous()
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    guard = 4096
    raw = torch.full((n * S * 4 + guard,), 777.0, device=dev)
    save_bytes = packed.mlp_save_bytes(n * S)
    assert save_bytes == 5 * 10 * 131072
    save = torch.full((save_bytes + guard,), 0x5A, dtype=torch.uint8, device=dev)
    p = lambda t: C.c_void_p(t.data_ptr())
    assert L.nerfq_mlp_forward(C.c_void_p(pn.ptr), p(r), p(z), C.c_longlong(n), S, p(raw), p(save), 0, stream) == 0
    torch.cuda.synchronize()
    assert (raw[n * S * 4:] == 777.0).all() and torch.isfinite(raw[:n * S * 4]).all()
    assert (save[save_bytes:] == 0x5A).all()
    ref = packed.mlp_forward(pn, r, z)                       # same call without save through the Python wrapper
    assert torch.equal(ref.reshape(-1), raw[:n * S * 4])
    d_raw = torch.randn(n * S * 4, device=dev) * 1e-4
    d_scale = torch.full((2436 + guard,), 0.0, device=dev)
    d_scale[2436:] = 777.0
    assert L.nerfq_mlp_backward(C.c_void_p(pn.ptr), p(d_raw), p(raw), p(save), C.c_longlong(n * S), p(d_scale), 0, stream) == 0
    torch.cuda.synchronize()
    assert (d_scale[2436:] == 777.0).all() and torch.isfinite(d_scale[:2436]).all() and float(d_scale[:2436].abs().max()) > 0
    assert (save[save_bytes:] == 0x5A).all() and (raw[n * S * 4:] == 777.0).all()


@pytest.mark.parametrize("n,S,Ni", [(37, 33, 21), (13, 64, 128), (5, 17, 5), (9, 200, 56)])
def test_ray_kernels_stay_inside_their_buffers(dev, n, S, Ni):
    """Ragged sizes through the C ABI with sentinel-filled guard regions behind every output of the ray-side kernels (depths,
    compositing forward / backward, sampling + sort with its vector stores): nothing is written outside the documented
    extents, and every documented element is written."""
    import ctypes as C
    ops = _ops()
    L = ops.L()
    gen = torch.Generator().manual_seed(n * 1000 + S)
    rays = synth_rays(n, 17).to(dev)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
    G, SENT = 1024, 777.0

    def guarded(count):
        return torch.full((count + G,), SENT, device=dev)

    def check(buf, count, name):
        torch.cuda.synchronize()
        assert (buf[count:] == SENT).all(), f"{name}: wrote past its end"
        assert not (buf[:count] == SENT).any(), f"{name}: left elements unwritten"
    t_rand = torch.rand(n, S, generator=gen).to(dev)
    z = guarded(n * S)
    assert L.nerfq_coarse_depths(p(rays), p(t_rand), C.c_longlong(n), S, 0, p(z), stream) == 0
    check(z, n * S, "coarse_depths")
    z0 = z[:n * S].reshape(n, S).contiguous()
    raw = torch.randn(n, S, 4, generator=gen).to(dev)
    rgb, disp, acc, depth, wts = guarded(3 * n), guarded(n), guarded(n), guarded(n), guarded(n * S)
    assert L.nerfq_composite_fwd(p(raw), p(z0), p(rays), None, 1, C.c_longlong(n), S, p(rgb), p(disp), p(acc), p(depth), p(wts), stream) == 0
    for buf, cnt, name in ((rgb, 3 * n, "rgb"), (acc, n, "acc"), (depth, n, "depth"), (wts, n * S, "weights")):
        check(buf, cnt, "composite_fwd " + name)
    torch.cuda.synchronize()
    assert (disp[n:] == SENT).all()                      # disp itself may hold NaN (0/0) by the reference's semantics
    d_rgb = torch.randn(n, 3, generator=gen).to(dev)
    d_raw = guarded(n * S * 4)
    assert L.nerfq_composite_bwd(p(raw), p(z0), p(rays), None, 1, p(d_rgb), C.c_longlong(n), S, p(d_raw), stream) == 0
    check(d_raw, n * S * 4, "composite_bwd")
    w = wts[:n * S].reshape(n, S).contiguous()
    for u in (None, torch.rand(n, Ni, generator=gen).to(dev)):
        z_all, z_std, zs = guarded(n * (S + Ni)), guarded(n), guarded(n * Ni)
        assert L.nerfq_sample_fine(p(z0), None, p(w), p(u), C.c_longlong(n), S, Ni, p(z_all), p(z_std), p(zs), stream) == 0
        check(z_all, n * (S + Ni), "sample_fine z_all")
        check(z_std, n, "sample_fine z_std")
        check(zs, n * Ni, "sample_fine z_samples")
        za = z_all[:n * (S + Ni)].reshape(n, S + Ni)
        assert (za[:, 1:] >= za[:, :-1]).all()
        assert torch.equal(torch.sort(torch.cat([z0, zs[:n * Ni].reshape(n, Ni)], -1), -1).values, za)
