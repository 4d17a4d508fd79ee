"""CPU oracle for the NeRF ray-rendering hot path (TEST INFRASTRUCTURE, not product code).

This file is a plain fp32 torch-on-CPU restatement of the reference's algorithm for the path
SURVEY.md section 8 scopes.  Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import it; the product package never does.

Every function cites the reference lines it restates (paths relative to /root/reference):

  positional_encoding    framework/nerf_model/run_nerf_helpers.py:18-67
  mlp_forward            utils.py:57-80 with framework/applications/utils/transforms.py:104-111
  query_network          framework/nerf_model/run_nerf.py:46-63
  composite              framework/nerf_model/run_nerf.py:285-345
  importance_sample      framework/nerf_model/run_nerf_helpers.py:119-163
  render_rays            framework/nerf_model/run_nerf.py:348-457
  camera_rays / ndc_rays framework/nerf_model/run_nerf_helpers.py:71-115
  render                 framework/nerf_model/run_nerf.py:81-158
  lsa_loss               framework/nerf_model/run_nerf.py:741-752

Parity pinning: tests/golden/*.npz were produced by running the UNMODIFIED reference functions
(tests/golden/make_golden.py); tests/test_oracle_golden.py checks this restatement against them.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

PTS_FREQS = 10     # multires      (run_nerf_helpers.py:52-67 called with multires=10)
DIR_FREQS = 4      # multires_views
NET_DEPTH = 8
NET_WIDTH = 256
SKIP_AFTER = 4     # utils.py:18 skips=[4]: concat happens after layer index 4


# --------------------------------------------------------------------------------------------
# positional encoding
# --------------------------------------------------------------------------------------------
def positional_encoding(x: torch.Tensor, n_freqs: int) -> torch.Tensor:
    """gamma(x) = [x, sin(2^0 x), cos(2^0 x), ..., sin(2^(L-1) x), cos(2^(L-1) x)]
    (run_nerf_helpers.py:23-49; log-sampled bands are exact powers of two)."""
    pieces = [x]
    for level in range(n_freqs):
        f = float(2 ** level)
        pieces.append(torch.sin(x * f))
        pieces.append(torch.cos(x * f))
    return torch.cat(pieces, dim=-1)


# --------------------------------------------------------------------------------------------
# the MLP (one NeRF network) with optional LSA scales
# --------------------------------------------------------------------------------------------
def _affine(params: Dict[str, torch.Tensor], prefix: str, x: torch.Tensor) -> torch.Tensor:
    """y = (s * W) x + b   (transforms.py:104-111); plain Linear when no scale is present."""
    w = params[prefix + ".weight"]
    b = params[prefix + ".bias"]
    s = params.get(prefix + ".weight_scaling")
    if s is not None:
        w = s.reshape(-1, 1) * w
    return torch.nn.functional.linear(x, w, b)


def mlp_forward(params: Dict[str, torch.Tensor], net: str, enc_pts: torch.Tensor,
                enc_dirs: torch.Tensor) -> torch.Tensor:
    """utils.py:57-80.  `net` is 'model' (coarse) or 'model_fine'.  Returns [M,4] = rgb(3), sigma."""
    h = enc_pts
    for i in range(NET_DEPTH):
        h = torch.relu(_affine(params, f"{net}.pts_linears.{i}", h))
        if i == SKIP_AFTER:
            h = torch.cat([enc_pts, h], dim=-1)
    sigma = _affine(params, f"{net}.alpha_linear", h)
    feat = _affine(params, f"{net}.feature_linear", h)
    hv = torch.relu(_affine(params, f"{net}.views_linears.0", torch.cat([feat, enc_dirs], dim=-1)))
    rgb = _affine(params, f"{net}.rgb_linear", hv)
    return torch.cat([rgb, sigma], dim=-1)


def query_network(params, net: str, pts: torch.Tensor, viewdirs: torch.Tensor) -> torch.Tensor:
    """run_nerf.py:46-63: encode points, broadcast+encode the per-ray view direction, run the MLP."""
    n_rays, n_samp, _ = pts.shape
    enc_p = positional_encoding(pts.reshape(-1, 3), PTS_FREQS)
    dirs = viewdirs[:, None, :].expand(n_rays, n_samp, 3).reshape(-1, 3)
    enc_d = positional_encoding(dirs, DIR_FREQS)
    return mlp_forward(params, net, enc_p, enc_d).reshape(n_rays, n_samp, 4)


# --------------------------------------------------------------------------------------------
# volume rendering
# --------------------------------------------------------------------------------------------
def composite(raw: torch.Tensor, z: torch.Tensor, rays_d: torch.Tensor, white_bkgd: bool = False,
              noise: Optional[torch.Tensor] = None):
    """run_nerf.py:301-343.  Returns rgb[N,3], disp[N], acc[N], weights[N,S], depth[N]."""
    far_gap = torch.full_like(z[:, :1], 1e10)
    delta = torch.cat([z[:, 1:] - z[:, :-1], far_gap], dim=-1)
    delta = delta * torch.norm(rays_d[:, None, :], dim=-1)
    colour = torch.sigmoid(raw[..., :3])
    sigma = raw[..., 3] if noise is None else raw[..., 3] + noise
    alpha = 1.0 - torch.exp(-torch.relu(sigma) * delta)
    ones = torch.ones((alpha.shape[0], 1), dtype=alpha.dtype)
    trans = torch.cumprod(torch.cat([ones, 1.0 - alpha + 1e-10], dim=-1), dim=-1)[:, :-1]
    weights = alpha * trans
    rgb = torch.sum(weights[..., None] * colour, dim=-2)
    depth = torch.sum(weights * z, dim=-1)
    acc = torch.sum(weights, dim=-1)
    disp = 1.0 / torch.max(1e-10 * torch.ones_like(depth), depth / acc)
    if white_bkgd:
        rgb = rgb + (1.0 - acc[..., None])
    return rgb, disp, acc, weights, depth


def importance_sample(bins: torch.Tensor, weights: torch.Tensor, n_samples: int,
                      u: Optional[torch.Tensor] = None) -> torch.Tensor:
    """run_nerf_helpers.py:119-163.  `u=None` is the deterministic (perturb==0) mode."""
    w = weights + 1e-5
    pdf = w / torch.sum(w, dim=-1, keepdim=True)
    cdf = torch.cumsum(pdf, dim=-1)
    cdf = torch.cat([torch.zeros_like(cdf[:, :1]), cdf], dim=-1)
    if u is None:
        u = torch.linspace(0.0, 1.0, steps=n_samples).expand(cdf.shape[0], n_samples)
    u = u.contiguous()
    idx = torch.searchsorted(cdf, u, right=True)
    lo = torch.clamp(idx - 1, min=0)
    hi = torch.clamp(idx, max=cdf.shape[-1] - 1)
    cdf_lo, cdf_hi = torch.gather(cdf, 1, lo), torch.gather(cdf, 1, hi)
    bin_lo, bin_hi = torch.gather(bins, 1, lo), torch.gather(bins, 1, hi)
    span = cdf_hi - cdf_lo
    span = torch.where(span < 1e-5, torch.ones_like(span), span)
    t = (u - cdf_lo) / span
    return bin_lo + t * (bin_hi - bin_lo)


def coarse_depths(near: torch.Tensor, far: torch.Tensor, n_samples: int, lindisp: bool = False,
                  t_rand: Optional[torch.Tensor] = None) -> torch.Tensor:
    """run_nerf.py:379-403.  near/far are [N,1]; t_rand [N,S] in [0,1) enables stratified jitter."""
    t = torch.linspace(0.0, 1.0, steps=n_samples)
    if lindisp:
        z = 1.0 / (1.0 / near * (1.0 - t) + 1.0 / far * t)
    else:
        z = near * (1.0 - t) + far * t
    z = z.expand(near.shape[0], n_samples)
    if t_rand is not None:
        mid = 0.5 * (z[:, 1:] + z[:, :-1])
        upper = torch.cat([mid, z[:, -1:]], dim=-1)
        lower = torch.cat([z[:, :1], mid], dim=-1)
        z = lower + (upper - lower) * t_rand
    return z


def render_rays(params, ray_batch: torch.Tensor, n_samples: int = 64, n_importance: int = 128,
                white_bkgd: bool = False, lindisp: bool = False,
                t_rand: Optional[torch.Tensor] = None, u: Optional[torch.Tensor] = None,
                noise0: Optional[torch.Tensor] = None, noise1: Optional[torch.Tensor] = None,
                retraw: bool = False):
    """run_nerf.py:348-457 with the RNG draws (t_rand, u, noise) lifted to arguments.

    ray_batch rows are [o(3), d(3), near, far, viewdirs(3)].
    """
    o, d = ray_batch[:, 0:3], ray_batch[:, 3:6]
    near, far = ray_batch[:, 6:7], ray_batch[:, 7:8]
    viewdirs = ray_batch[:, 8:11]
    z0 = coarse_depths(near, far, n_samples, lindisp, t_rand)
    pts0 = o[:, None, :] + d[:, None, :] * z0[:, :, None]
    raw0 = query_network(params, "model", pts0, viewdirs)
    rgb0, disp0, acc0, w0, _ = composite(raw0, z0, d, white_bkgd, noise0)
    out = {}
    if n_importance > 0:
        mid = 0.5 * (z0[:, 1:] + z0[:, :-1])
        z_new = importance_sample(mid, w0[:, 1:-1], n_importance, u).detach()
        z1, _ = torch.sort(torch.cat([z0, z_new], dim=-1), dim=-1)
        pts1 = o[:, None, :] + d[:, None, :] * z1[:, :, None]
        raw1 = query_network(params, "model_fine", pts1, viewdirs)
        rgb1, disp1, acc1, w1, _ = composite(raw1, z1, d, white_bkgd, noise1)
        out.update(rgb_map=rgb1, disp_map=disp1, acc_map=acc1, rgb0=rgb0, disp0=disp0, acc0=acc0,
                   z_std=torch.std(z_new, dim=-1, unbiased=False))
        if retraw:
            out["raw"] = raw1
        out["z_vals"] = z1
        out["weights"] = w1
    else:
        out.update(rgb_map=rgb0, disp_map=disp0, acc_map=acc0)
        if retraw:
            out["raw"] = raw0
        out["z_vals"] = z0
        out["weights"] = w0
    out["raw0"] = raw0
    out["z_vals0"] = z0
    out["weights0"] = w0
    return out


# --------------------------------------------------------------------------------------------
# ray generation
# --------------------------------------------------------------------------------------------
def camera_rays(H: int, W: int, K, c2w: torch.Tensor):
    """run_nerf_helpers.py:71-85: pinhole rays, image row-major [H,W,3]."""
    jj, ii = torch.meshgrid(torch.linspace(0, H - 1, H), torch.linspace(0, W - 1, W), indexing="ij")
    dirs = torch.stack([(ii - K[0][2]) / K[0][0], -(jj - K[1][2]) / K[1][1], -torch.ones_like(ii)], -1)
    rays_d = torch.sum(dirs[..., None, :] * c2w[:3, :3], dim=-1)
    rays_o = c2w[:3, -1].expand(rays_d.shape)
    return rays_o, rays_d


def ndc_rays(H: int, W: int, focal: float, near: float, rays_o: torch.Tensor, rays_d: torch.Tensor):
    """run_nerf_helpers.py:98-115."""
    t = -(near + rays_o[..., 2]) / rays_d[..., 2]
    rays_o = rays_o + t[..., None] * rays_d
    ax = -1.0 / (W / (2.0 * focal))
    ay = -1.0 / (H / (2.0 * focal))
    o0 = ax * rays_o[..., 0] / rays_o[..., 2]
    o1 = ay * rays_o[..., 1] / rays_o[..., 2]
    o2 = 1.0 + 2.0 * near / rays_o[..., 2]
    d0 = ax * (rays_d[..., 0] / rays_d[..., 2] - rays_o[..., 0] / rays_o[..., 2])
    d1 = ay * (rays_d[..., 1] / rays_d[..., 2] - rays_o[..., 1] / rays_o[..., 2])
    d2 = -2.0 * near / rays_o[..., 2]
    return torch.stack([o0, o1, o2], -1), torch.stack([d0, d1, d2], -1)


def pack_rays(H, W, K, rays=None, c2w=None, ndc=True, near=0.0, far=1.0):
    """run_nerf.py:108-142: the [N,11] ray rows render_rays consumes (use_viewdirs=True)."""
    if c2w is not None:
        rays_o, rays_d = camera_rays(H, W, K, c2w)
    else:
        rays_o, rays_d = rays
    shape = rays_d.shape
    viewdirs = rays_d / torch.norm(rays_d, dim=-1, keepdim=True)
    viewdirs = viewdirs.reshape(-1, 3).float()
    if ndc:
        rays_o, rays_d = ndc_rays(H, W, K[0][0], 1.0, rays_o, rays_d)
    rays_o = rays_o.reshape(-1, 3).float()
    rays_d = rays_d.reshape(-1, 3).float()
    nf = torch.ones_like(rays_d[:, :1])
    return torch.cat([rays_o, rays_d, near * nf, far * nf, viewdirs], dim=-1), shape


def render(params, H, W, K, chunk=32768, rays=None, c2w=None, ndc=True, near=0.0, far=1.0, **kw):
    """run_nerf.py:81-158 (use_viewdirs=True).  Returns rgb, disp, acc, extras."""
    packed, shape = pack_rays(H, W, K, rays, c2w, ndc, near, far)
    parts = []
    for i in range(0, packed.shape[0], chunk):
        sub = {k: (v[i:i + chunk] if v is not None else None)
               for k, v in kw.items() if k in ("t_rand", "u", "noise0", "noise1")}
        rest = {k: v for k, v in kw.items() if k not in sub}
        parts.append(render_rays(params, packed[i:i + chunk], **rest, **sub))
    merged = {k: torch.cat([p[k] for p in parts], 0) for k in parts[0]}
    for k in merged:
        merged[k] = merged[k].reshape(list(shape[:-1]) + list(merged[k].shape[1:]))
    return merged["rgb_map"], merged["disp_map"], merged["acc_map"], merged


# --------------------------------------------------------------------------------------------
# LSA objective
# --------------------------------------------------------------------------------------------
def lsa_loss(out: dict, target: torch.Tensor) -> torch.Tensor:
    """run_nerf.py:741-752: mse(rgb, t) + mse(rgb0, t), means over N*3."""
    loss = torch.mean((out["rgb_map"] - target) ** 2)
    if "rgb0" in out:
        loss = loss + torch.mean((out["rgb0"] - target) ** 2)
    return loss


def psnr_from_mse(mse: float) -> float:
    """run_nerf_helpers.py:13."""
    return -10.0 * math.log(mse) / math.log(10.0)


LAYER_NAMES = tuple([f"pts_linears.{i}" for i in range(NET_DEPTH)] +
                    ["alpha_linear", "feature_linear", "views_linears.0", "rgb_linear"])


def lsa_scale_grads(params, ray_batch, target, **kw):
    """Gradient of lsa_loss w.r.t. every weight_scaling tensor (autograd over this restatement,
    which is what the reference does at run_nerf.py:756 with only the scales trainable)."""
    p = {k: v.clone() for k, v in params.items()}
    scales = [k for k in p if k.endswith("weight_scaling")]
    for k in scales:
        p[k].requires_grad_(True)
    out = render_rays(p, ray_batch, **kw)
    loss = lsa_loss(out, target)
    grads = torch.autograd.grad(loss, [p[k] for k in scales])
    return float(loss.detach()), {k: g.detach() for k, g in zip(scales, grads)}, \
        {k: v.detach() for k, v in out.items()}
