"""GPU-box debug script (not a pytest): run the fused MLP forward on the golden model, decode the
saved operand tiles and print per-layer errors against the oracle."""
import sys
import os
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nerfq_b200  # noqa
from nerfq_b200 import packed
from oracle import render_oracle as ro
from tests.util import golden_model_params, golden_model_levels, synth_rays, LAYERS

SAVE_TILE = 9 * 65536 + 32768


def decode_block_index():
    r = np.arange(128)[:, None]
    k = np.arange(64)[None, :]
    return (r * 128 + (((k // 8) ^ (r & 7)) * 16) + (k % 8) * 2) // 2      # uint16 index within a 16 KB block


def decode_tile(buf_u16, tile, slot, width):
    idx = decode_block_index()
    base = (tile * SAVE_TILE + slot * 65536) // 2
    cols = []
    for b in range(width // 64):
        blk = buf_u16[base + b * 8192: base + (b + 1) * 8192]
        cols.append(blk[idx])
    return np.concatenate(cols, axis=1).view(np.float16).astype(np.float32)


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


def main():
    dev = torch.device("cuda:0")
    net_name = "model"
    pn, p = build_net(net_name, dev)
    n_rays, S = 300, 64
    rays = synth_rays(n_rays, 11)
    z = ro.coarse_depths(rays[:, 6:7], rays[:, 7:8], S).contiguous()
    for pingpong in (False, True):
        save = torch.zeros(packed.mlp_save_bytes(n_rays * S), dtype=torch.uint8, device=dev)
        raw = packed.mlp_forward(pn, rays.to(dev), z.to(dev), save=save, pingpong=pingpong)
        torch.cuda.synchronize()
        raw = raw.cpu().reshape(-1, 4)
        pts = (rays[:, None, 0:3] + rays[:, None, 3:6] * z[:, :, None]).reshape(-1, 3)
        vd = rays[:, None, 8:11].expand(n_rays, S, 3).reshape(-1, 3)
        with torch.no_grad():
            hs, feat, hv, raw_ref = oracle_intermediates(p, net_name, pts, vd)
        buf = save.cpu().numpy().view(np.uint16)
        M = n_rays * S
        ntiles = (M + 127) // 128
        print(f"--- pingpong={pingpong}  points={M} tiles={ntiles}")
        names = ["h1", "h2", "h3", "h4", "h5", "h6", "h7", "h8", "feat"]
        refs = hs + [feat]
        for slot, (nm, ref) in enumerate(zip(names, refs)):
            got = np.concatenate([decode_tile(buf, t, slot, 256) for t in range(ntiles)], 0)[:M]
            err = np.abs(got - ref.numpy())
            print(f"  {nm}: max|ref|={np.abs(ref.numpy()).max():.4f} maxerr={err.max():.5f} meanerr={err.mean():.6f}")
        got = np.concatenate([decode_tile(buf, t, 9, 128) for t in range(ntiles)], 0)[:M]
        err = np.abs(got - hv.numpy())
        print(f"  hv: max|ref|={np.abs(hv.numpy()).max():.4f} maxerr={err.max():.5f}")
        err = (raw - raw_ref).abs()
        print(f"  raw: max|ref|={raw_ref.abs().max():.4f} maxerr={err.max():.5f} per-channel {err.max(0).values.tolist()}")
        print("  raw sample", raw[:2].tolist(), raw_ref[:2].tolist())
        # no-save run must give identical raw
        raw2 = packed.mlp_forward(pn, rays.to(dev), z.to(dev), pingpong=pingpong).cpu().reshape(-1, 4)
        print("  save vs nosave identical:", bool((raw2 == raw).all()), " finite:", bool(torch.isfinite(raw).all()))
    # timing at a larger size
    n_rays = 16384
    rays = synth_rays(n_rays, 12).to(dev)
    for S in (64, 192):
        z = torch.sort(2.0 + 4.0 * torch.rand(n_rays, S, device=dev), -1).values.contiguous()
        for pingpong in (False, True):
            for _ in range(2):
                packed.mlp_forward(pn, rays, z, pingpong=pingpong)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                packed.mlp_forward(pn, rays, z, pingpong=pingpong)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            pts = n_rays * S
            print(f"timing S={S} pingpong={pingpong}: {ms:.3f} ms  {pts / ms / 1e6:.3f} Gpts/s  {pts * 1.186816e6 / ms / 1e9:.1f} TFLOP/s")


if __name__ == "__main__":
    main()
