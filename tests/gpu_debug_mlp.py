"""GPU-box debug script for the fused MLP forward kernel: per-layer errors from the saved
MN-major operand images, raw output error, timing."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nerfq_b200  # noqa
from nerfq_b200 import packed
from oracle import render_oracle as ro
from tests.util import golden_model_params, golden_model_levels, synth_rays, LAYERS


def oracle_intermediates(p, net, pts, vd):
    enc_p = ro.positional_encoding(pts, 10)
    enc_d = ro.positional_encoding(vd, 4)
    hs = []
    h = enc_p
    for i in range(8):
        h = torch.relu(ro._affine(p, f"{net}.pts_linears.{i}", h))
        hs.append(h)
        if i == 4:
            h = torch.cat([enc_p, h], -1)
    sigma = ro._affine(p, f"{net}.alpha_linear", h)
    feat = ro._affine(p, f"{net}.feature_linear", h)
    hv = torch.relu(ro._affine(p, f"{net}.views_linears.0", torch.cat([feat, enc_d], -1)))
    rgb = ro._affine(p, f"{net}.rgb_linear", hv)
    return hs, feat, hv, torch.cat([rgb, sigma], -1)


def build_net(net, dev):
    levels, delta = golden_model_levels()
    p, _ = golden_model_params()
    ws = [torch.from_numpy(levels[f"{net}.{l}"]).to(dev) for l in LAYERS]
    bs = [p[f"{net}.{l}.bias"].to(dev) for l in LAYERS]
    ss = [p[f"{net}.{l}.weight_scaling"].to(dev) for l in LAYERS]
    return packed.PackedNet(ws, [delta] * 12, bs, ss), p

ACT = 131072
GROUP_BYTES = 10 * ACT


def decode_image(buf_u8, group, slot, width):
    """saved-activation slot ([16 point chunks][256 channels][16 points] fp16) -> [256 points, width channels] float32."""
    k = np.arange(width)[None, :]
    n = np.arange(256)[:, None]
    off = ((n >> 4) * 256 + k) * 32 + (n & 15) * 2
    base = group * GROUP_BYTES + slot * ACT
    u16 = buf_u8[base: base + ACT].view(np.uint16)
    return u16[off // 2].view(np.float16).astype(np.float32)


def main():
    dev = torch.device("cuda:0")
    pn, p = build_net("model", dev)
    n_rays, S = 301, 64
    rays = synth_rays(n_rays, 11)
    z = ro.coarse_depths(rays[:, 6:7], rays[:, 7:8], S).contiguous()
    M = n_rays * S
    save = torch.zeros(packed.mlp_save_bytes(M), dtype=torch.uint8, device=dev)
    raw = packed.mlp_forward(pn, rays.to(dev), z.to(dev), save=save)
    torch.cuda.synchronize()
    raw = raw.cpu().reshape(-1, 4)
    pts = (rays[:, None, 0:3] + rays[:, None, 3:6] * z[:, :, None]).reshape(-1, 3)
    vd = rays[:, None, 8:11].expand(n_rays, S, 3).reshape(-1, 3)
    with torch.no_grad():
        hs, feat, hv, raw_ref = oracle_intermediates(p, "model", pts, vd)
    buf = save.cpu().numpy()
    groups = (M + 255) // 256
    print(f"--- forward: points={M} groups={groups}")
    for slot, (nm, ref) in enumerate(zip(["h1", "h2", "h3", "h4", "h5", "h6", "h7", "h8", "feat"], hs + [feat])):
        got = np.concatenate([decode_image(buf, g, slot, 256) for g in range(groups)], 0)[:M]
        err = np.abs(got - ref.numpy())
        print(f"  {nm}: max|ref|={np.abs(ref.numpy()).max():.4f} maxerr={err.max():.5f} meanerr={err.mean():.6f}")
    got = np.concatenate([decode_image(buf, g, 9, 128) for g in range(groups)], 0)[:M]
    print(f"  hv: maxerr={np.abs(got - hv.numpy()).max():.5f}")
    err = (raw - raw_ref).abs()
    print(f"  raw: max|ref|={raw_ref.abs().max():.4f} maxerr={err.max():.5f} per-channel {err.max(0).values.tolist()}")
    raw2 = packed.mlp_forward(pn, rays.to(dev), z.to(dev)).cpu().reshape(-1, 4)
    print("  save vs nosave identical:", bool((raw2 == raw).all()), " finite:", bool(torch.isfinite(raw).all()))
    n_rays = 16384
    rays = synth_rays(n_rays, 12).to(dev)
    for S in (64, 192):
        z = torch.sort(2.0 + 4.0 * torch.rand(n_rays, S, device=dev), -1).values.contiguous()
        save = torch.empty(packed.mlp_save_bytes(n_rays * S), dtype=torch.uint8, device=dev)
        for name, kw in (("nosave", dict()), ("save", dict(save=save))):
            for _ in range(2):
                packed.mlp_forward(pn, rays, z, **kw)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                packed.mlp_forward(pn, rays, z, **kw)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            pts = n_rays * S
            print(f"timing S={S} {name}: {ms:.3f} ms  {pts / ms / 1e6:.3f} Gpts/s  {pts * 1.186816e6 / ms / 1e9:.1f} TFLOP/s")
        del save


if __name__ == "__main__":
    main()
