"""The reference's rendering call surface on top of the fused CUDA path.

Mirrors framework/nerf_model/run_nerf.py of the reference (same names, arguments, return values):

  render            run_nerf.py:81-158      render_path   run_nerf.py:161-211
  batchify_rays     run_nerf.py:66-78       render_rays   run_nerf.py:348-457
  raw2outputs       run_nerf.py:285-345     create_nerf   run_nerf.py:214-281
  run_network       run_nerf.py:46-63
and of run_nerf_helpers.py: sample_pdf :119-163, get_rays :71-85, ndc_rays :98-115, img2mse/mse2psnr/to8b :12-14.

Everything numerical runs in the hand-written kernels (ops.py); torch supplies device memory, the RNG
draws the reference makes with torch.rand/torch.randn, and autograd plumbing: `render_rays` is a
torch.autograd.Function whose backward runs the compositing-backward and MLP-backward kernels and hands
the LSA-scale gradients to the `weight_scaling` parameters, so the reference's
`loss.backward(); optimizer.step()` (run_nerf.py:756-757) works unchanged.
"""
import os
from typing import Optional

import numpy as np
import torch

from . import ops, packed
from .model import NeRF

# ---- small helpers the reference exposes ---------------------------------------------------------
img2mse = lambda x, y: torch.mean((x - y) ** 2)
mse2psnr = lambda x: -10. * torch.log(x) / torch.log(torch.tensor([10.], device=x.device))
to8b = lambda x: (255 * np.clip(x, 0, 1)).astype(np.uint8)

TUNING = {"max_ctas": 0}      # kernel scheduling knob (bench/profiling)
# data-parallel LSA tuning: all-reduce the scale gradients over torch.distributed (group None = the default group)
DATA_PARALLEL = {"enabled": False, "group": None, "peer": None, "collective": None}


class _FusedQuery:
    """Placeholder returned by create_nerf as `network_query_fn`: tells render_rays to use the fused
    PE+MLP kernel.  Calling it evaluates explicit points like the reference's run_network."""

    def __call__(self, inputs, viewdirs, network_fn):
        return run_network(inputs, viewdirs, network_fn)


def _refresh(net: NeRF, scale_flat: Optional[torch.Tensor]) -> packed.PackedNet:
    pn = net.packed_net()
    pn.set_scales(flat=scale_flat)
    return pn


def _scale_flat(net: NeRF) -> Optional[torch.Tensor]:
    sc = net.scale_tensors()
    if sc[0] is None:
        return None
    return packed.flatten_channels(sc)


def run_network(inputs, viewdirs, fn, embed_fn=None, embeddirs_fn=None, netchunk=1024 * 64):
    """run_nerf.py:46-63 for explicit sample points [N,S,3]: each point becomes a zero-length ray."""
    n, s, _ = inputs.shape
    pts = inputs.reshape(-1, 3).float()
    vd = viewdirs[:, None, :].expand(n, s, 3).reshape(-1, 3).float()
    rays = torch.cat([pts, torch.zeros_like(pts), torch.zeros_like(pts[:, :2]), vd], -1).contiguous()
    z = torch.zeros((n * s, 1), dtype=torch.float32, device=pts.device)
    with torch.no_grad():
        pn = _refresh(fn, _scale_flat(fn))
        raw = packed.mlp_forward(pn, rays, z, max_ctas=TUNING["max_ctas"])
    return raw.reshape(n, s, 4)


def raw2outputs(raw, z_vals, rays_d, raw_noise_std=0, white_bkgd=False, pytest=False):
    """run_nerf.py:285-345 (forward only): rgb_map, disp_map, acc_map, weights, depth_map."""
    n, s = z_vals.shape
    noise = None
    if raw_noise_std > 0.:
        noise = torch.randn((n, s), device=raw.device) * raw_noise_std
        if pytest:
            np.random.seed(0)
            noise = torch.tensor(np.random.rand(n, s) * raw_noise_std, dtype=torch.float32, device=raw.device)
    rays = torch.zeros((n, 11), dtype=torch.float32, device=raw.device)
    rays[:, 3:6] = rays_d
    return ops.composite_fwd(raw.float().contiguous(), z_vals.float().contiguous(), rays, white_bkgd, noise)


def sample_pdf(bins, weights, N_samples, det=False, pytest=False):
    """run_nerf_helpers.py:119-163: bins [N,B] (midpoints), weights [N,B-1] -> samples [N,N_samples]."""
    n = bins.shape[0]
    u = None
    if not det:
        u = torch.rand((n, N_samples), device=bins.device)
    if pytest:
        np.random.seed(0)
        if det:
            u = torch.tensor(np.broadcast_to(np.linspace(0., 1., N_samples), (n, N_samples)).copy(), dtype=torch.float32, device=bins.device)
        else:
            u = torch.tensor(np.random.rand(n, N_samples), dtype=torch.float32, device=bins.device)
    return ops.sample_pdf(bins.float(), weights.float(), N_samples, u)


def get_rays(H, W, K, c2w):
    """run_nerf_helpers.py:71-85 -> rays_o, rays_d  [H,W,3]."""
    dev = c2w.device if torch.is_tensor(c2w) and c2w.is_cuda else torch.device("cuda")
    c = c2w.detach().cpu().numpy() if torch.is_tensor(c2w) else np.asarray(c2w)
    rays = ops.camera_rays(H, W, K, c, False, 0., 1., dev)
    return rays[:, 0:3].reshape(H, W, 3), rays[:, 3:6].reshape(H, W, 3)


def ndc_rays(H, W, focal, near, rays_o, rays_d):
    """run_nerf_helpers.py:98-115 (near plane at `near`=1 as render() calls it)."""
    assert float(near) == 1.0, "the kernel implements the near=1 NDC warp render() uses (run_nerf.py:133)"
    sh = rays_d.shape
    rays = ops.pack_rays(rays_o, rays_d, True, H, W, float(focal), 0., 1.)
    return rays[:, 0:3].reshape(sh), rays[:, 3:6].reshape(sh)


# ---- the differentiable renderer -------------------------------------------------------------------
class _Cfg:
    __slots__ = ("net0", "net1", "rays", "S", "Ni", "lindisp", "white", "t_rand", "u", "noise0", "noise1", "retraw", "need_grad")


def _forward_pipeline(cfg: _Cfg, sc0, sc1, save: bool):
    mc = TUNING["max_ctas"]
    rays = cfg.rays
    n = rays.shape[0]
    dev = rays.device
    pn0 = _refresh(cfg.net0, sc0)
    z0 = ops.coarse_depths(rays, cfg.S, cfg.lindisp, cfg.t_rand)
    save0 = torch.empty(packed.mlp_save_bytes(n * cfg.S), dtype=torch.uint8, device=dev) if save else None
    raw0 = packed.mlp_forward(pn0, rays, z0, save=save0, max_ctas=mc)
    rgb0, disp0, acc0, w0, _ = ops.composite_fwd(raw0, z0, rays, cfg.white, cfg.noise0)
    st = dict(z0=z0, raw0=raw0, save0=save0, pn0=pn0)
    if cfg.Ni > 0:
        net1 = cfg.net1 if cfg.net1 is not None else cfg.net0
        pn1 = _refresh(net1, sc1 if cfg.net1 is not None else sc0)
        z1, z_std, _ = ops.sample_fine(z0, w0, cfg.Ni, cfg.u)
        save1 = torch.empty(packed.mlp_save_bytes(n * (cfg.S + cfg.Ni)), dtype=torch.uint8, device=dev) if save else None
        raw1 = packed.mlp_forward(pn1, rays, z1, save=save1, max_ctas=mc)
        rgb1, disp1, acc1, _, _ = ops.composite_fwd(raw1, z1, rays, cfg.white, cfg.noise1, want_weights=False)
        st.update(z1=z1, raw1=raw1, save1=save1, pn1=pn1)
        outs = (rgb1, disp1, acc1, rgb0, disp0, acc0, z_std, raw1)
    else:
        outs = (rgb0, disp0, acc0, raw0)
    return outs, st


class _DPState:
    """Per-device scratch of the data-parallel backward: the two networks' fixed-point gradient buffers, contiguous so
    that one all-reduce covers both (kept zeroed by nerfq_mlp_backward_finalize)."""
    _by_device = {}

    def __init__(self, dev):
        self.fix = torch.zeros((2, ops.grad_fix_elems()), dtype=torch.int64, device=dev)

    @classmethod
    def get(cls, dev):
        key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device())
        if key not in cls._by_device:
            cls._by_device[key] = cls(dev)
        return cls._by_device[key]


def _backward_pipeline(cfg: _Cfg, st: dict, d_rgb1, d_rgb0, g_out: Optional[torch.Tensor] = None):
    """d loss / d LSA scales of both networks from d loss / d rgb_map (fine, d_rgb1) and d loss / d rgb0 (coarse, d_rgb0):
    compositing backward + MLP backward per network, fine first.  Returns (g0, g1): flat [2436] gradients in kernel channel
    order for network_fn and network_fine (g1 None when the fine pass used network_fn).  g_out: optional zeroed [2, 2436]
    buffer to accumulate into.

    Data parallel (DATA_PARALLEL['enabled'], one process per GPU): every rank's kernels leave s*ds as 64-bit FIXED-POINT
    sums; those are all-reduced as integers (NCCL int64 sum, 19.5 KB per network) before the conversion to float, so the
    result does not depend on how the batch is split over ranks.  Both networks' buffers go through one all-reduce after the
    last backward kernel."""
    mc = TUNING["max_ctas"]
    rays = cfg.rays
    dev = rays.device
    dp = DATA_PARALLEL["enabled"]
    own_fine = cfg.Ni > 0 and cfg.net1 is not None
    if g_out is None:
        g_out = torch.zeros((2, 2436), dtype=torch.float32, device=dev)
    g0, g1 = g_out[0], (g_out[1] if own_fine else None)
    passes = []                                   # (packed net, raw, z, save, noise, d_rgb, slot) in execution order
    if cfg.Ni > 0 and d_rgb1 is not None:
        passes.append((st["pn1"], st["raw1"], st["z1"], st["save1"], cfg.noise1, d_rgb1, 1 if own_fine else 0))
    if d_rgb0 is not None:
        passes.append((st["pn0"], st["raw0"], st["z0"], st["save0"], cfg.noise0, d_rgb0, 0))
    if not dp:
        for pn, raw, z, save, noise, d_rgb, slot in passes:
            d_raw = ops.composite_bwd(raw, z, rays, cfg.white, d_rgb.contiguous(), noise)
            ops.mlp_backward(pn, d_raw, raw, save, g_out[slot], max_ctas=mc)
        return g0, g1
    import torch.distributed as dist
    group = DATA_PARALLEL.get("group")
    dps = _DPState.get(dev)
    slots = []
    for pn, raw, z, save, noise, d_rgb, slot in passes:
        d_raw = ops.composite_bwd(raw, z, rays, cfg.white, d_rgb.contiguous(), noise)
        ops.mlp_backward_partial(pn, d_raw, raw, save, dps.fix[slot], max_ctas=mc)
        if slot not in slots:
            slots.append(slot)
    peer = DATA_PARALLEL.get("peer")
    if peer is not None and g_out.is_contiguous() and g_out.shape[0] == 2:
        # all-reduce fused into the finalize: one kernel publishes this rank's sums, meets the peers and adds up everybody's
        # sums read straight from their memory (include/nerfq.h: nerfq_mlp_backward_finalize_peers)
        by_slot = {slot: pn for pn, _, _, _, _, _, slot in passes}
        ops.mlp_backward_finalize_peers(by_slot.get(0, by_slot.get(1)), by_slot.get(1), dps.fix, int(peer.ptrs.data_ptr()), peer.world, peer.rank,
                                        peer.epoch, g_out)
        return g0, g1
    # ONE all-reduce for both networks (39 KB of int64).  Overlapping the fine network's all-reduce with the coarse
    # network's backward on a side stream was tried in round 2 and bought nothing: the persistent backward kernel holds
    # every SM, so NCCL's kernel only got an SM when that kernel ended, and the two all-reduces ran back to back at the
    # tail anyway (0.082 ms exposed at N = 8, 2 x 0.030 ms alone).
    dist.all_reduce(dps.fix if len(slots) == 2 else dps.fix[slots[0]], op=dist.ReduceOp.SUM, group=group)
    done = set()
    for pn, _, _, _, _, _, slot in passes:
        if slot not in done:
            ops.mlp_backward_finalize(pn, dps.fix[slot], g_out[slot])
            done.add(slot)
    return g0, g1


class _RenderRaysFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cfg: _Cfg, sc0, sc1):
        outs, st = _forward_pipeline(cfg, sc0, sc1, save=True)
        ctx.cfg, ctx.st = cfg, st
        ctx.has1 = sc1 is not None
        ctx.mark_non_differentiable(*[o for i, o in enumerate(outs) if i not in ((0, 3) if cfg.Ni > 0 else (0,))])
        return outs

    @staticmethod
    def backward(ctx, *grads):
        cfg, st = ctx.cfg, ctx.st
        d_rgb1, d_rgb0 = (grads[0], grads[3]) if cfg.Ni > 0 else (None, grads[0])
        if DATA_PARALLEL["enabled"]:
            # the caller's loss is a mean over THIS rank's rays; the objective is the mean over the global batch
            import torch.distributed as dist
            inv = 1.0 / dist.get_world_size(DATA_PARALLEL.get("group"))
            d_rgb1 = d_rgb1 * inv if d_rgb1 is not None else None
            d_rgb0 = d_rgb0 * inv if d_rgb0 is not None else None
        g0, g1 = _backward_pipeline(cfg, st, d_rgb1, d_rgb0)
        ctx.st = None
        return None, g0, (g1 if ctx.has1 else None)


def _make_cfg(rays, network_fn, network_fine, N_samples, N_importance, lindisp, perturb, white_bkgd, raw_noise_std, pytest,
              retraw=False) -> _Cfg:
    """The per-call configuration of render_rays incl. the random draws the reference makes (run_nerf.py:395-403 t_rand,
    run_nerf_helpers.py:131-145 u, run_nerf.py:316-322 noise), in the reference's order."""
    n, dev = rays.shape[0], rays.device
    cfg = _Cfg()
    cfg.net0, cfg.net1, cfg.rays = network_fn, network_fine, rays
    cfg.S, cfg.Ni, cfg.lindisp, cfg.white, cfg.retraw = int(N_samples), int(N_importance), bool(lindisp), bool(white_bkgd), retraw
    cfg.t_rand = cfg.u = cfg.noise0 = cfg.noise1 = None
    if perturb > 0.:
        cfg.t_rand = torch.rand((n, cfg.S), device=dev)
        if pytest:
            np.random.seed(0)
            cfg.t_rand = torch.tensor(np.random.rand(n, cfg.S), dtype=torch.float32, device=dev)
    if raw_noise_std > 0.:
        cfg.noise0 = torch.randn((n, cfg.S), device=dev) * raw_noise_std
        if pytest:
            np.random.seed(0)
            cfg.noise0 = torch.tensor(np.random.rand(n, cfg.S) * raw_noise_std, dtype=torch.float32, device=dev)
    if cfg.Ni > 0:
        if perturb != 0.:
            cfg.u = torch.rand((n, cfg.Ni), device=dev)
        if pytest:
            np.random.seed(0)
            cfg.u = None if perturb == 0. else torch.tensor(np.random.rand(n, cfg.Ni), dtype=torch.float32, device=dev)
        if raw_noise_std > 0.:
            cfg.noise1 = torch.randn((n, cfg.S + cfg.Ni), device=dev) * raw_noise_std
            if pytest:
                np.random.seed(0)
                cfg.noise1 = torch.tensor(np.random.rand(n, cfg.S + cfg.Ni) * raw_noise_std, dtype=torch.float32, device=dev)
    return cfg


def render_rays(ray_batch, network_fn, network_query_fn=None, N_samples=64, retraw=False, lindisp=False, perturb=0.,
                N_importance=0, network_fine=None, white_bkgd=False, raw_noise_std=0., pytest=False, verbose=False):
    """run_nerf.py:348-457.  ray_batch rows: [o, d, near, far, viewdirs]; returns the same dict."""
    assert ray_batch.is_cuda, "render_rays needs CUDA tensors (there is no CPU path)"
    if ray_batch.shape[-1] <= 8:
        raise NotImplementedError("the fused kernels implement the use_viewdirs=True configuration of the reference")
    rays = ray_batch.float().contiguous()
    cfg = _make_cfg(rays, network_fn, network_fine, N_samples, N_importance, lindisp, perturb, white_bkgd, raw_noise_std, pytest, retraw)

    sc0 = _scale_flat(network_fn)
    sc1 = _scale_flat(network_fine) if (network_fine is not None and cfg.Ni > 0) else None
    need_grad = torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in (sc0, sc1))
    if need_grad:
        outs = _RenderRaysFn.apply(cfg, sc0, sc1)
    elif not retraw and cfg.t_rand is None and cfg.u is None and cfg.noise0 is None and cfg.noise1 is None:
        # the test-view path (perturb = 0, raw_noise_std = 0): the whole of render_rays is one C-ABI call
        with torch.no_grad():
            pn0 = _refresh(network_fn, sc0)
            pn1 = None
            if cfg.Ni > 0 and network_fine is not None:
                pn1 = _refresh(network_fine, sc1)
            rgb, disp, acc, rgb0, disp0, acc0, z_std = ops.render_rays_fwd(pn0, pn1, rays, cfg.S, cfg.Ni, cfg.lindisp, cfg.white,
                                                                           max_ctas=TUNING["max_ctas"])
        ret = {"rgb_map": rgb, "disp_map": disp, "acc_map": acc}
        if cfg.Ni > 0:
            ret.update(rgb0=rgb0, disp0=disp0, acc0=acc0, z_std=z_std)
        return ret
    else:
        with torch.no_grad():
            outs, _ = _forward_pipeline(cfg, sc0, sc1, save=False)
    if cfg.Ni > 0:
        rgb, disp, acc, rgb0, disp0, acc0, z_std, raw = outs
        ret = {"rgb_map": rgb, "disp_map": disp, "acc_map": acc}
        if retraw:
            ret["raw"] = raw
        ret.update(rgb0=rgb0, disp0=disp0, acc0=acc0, z_std=z_std)
    else:
        rgb, disp, acc, raw = outs
        ret = {"rgb_map": rgb, "disp_map": disp, "acc_map": acc}
        if retraw:
            ret["raw"] = raw
    return ret


def batchify_rays(rays_flat, chunk=1024 * 32, **kwargs):
    """run_nerf.py:66-78."""
    all_ret = {}
    for i in range(0, rays_flat.shape[0], chunk):
        ret = render_rays(rays_flat[i:i + chunk], **kwargs)
        for k in ret:
            all_ret.setdefault(k, []).append(ret[k])
    return {k: (v[0] if len(v) == 1 else torch.cat(v, 0)) for k, v in all_ret.items()}


def render(H, W, K, chunk=1024 * 32, rays=None, c2w=None, ndc=True, near=0., far=1., use_viewdirs=False,
           c2w_staticcam=None, **kwargs):
    """run_nerf.py:81-158: returns [rgb_map, disp_map, acc_map, extras]."""
    if not use_viewdirs:
        raise NotImplementedError("the fused kernels implement the use_viewdirs=True configuration of the reference")
    if c2w_staticcam is not None:
        raise NotImplementedError("c2w_staticcam (a visualisation aid, run_nerf.py:122-124) is outside the accelerated path")
    if c2w is not None:
        dev = c2w.device if torch.is_tensor(c2w) and c2w.is_cuda else torch.device("cuda")
        c = c2w.detach().cpu().numpy() if torch.is_tensor(c2w) else np.asarray(c2w)
        packed_rays = ops.camera_rays(H, W, K, c, bool(ndc), float(near), float(far), dev)
        sh = (H, W, 3)
    else:
        rays_o, rays_d = rays
        sh = tuple(rays_d.shape)
        focal = float(K[0][0]) if ndc else 1.0
        packed_rays = ops.pack_rays(rays_o.float().cuda(), rays_d.float().cuda(), bool(ndc), int(H), int(W), focal, float(near), float(far))
    kwargs.pop("network_query_fn", None)
    all_ret = batchify_rays(packed_rays, chunk, **kwargs)
    for k in all_ret:
        all_ret[k] = torch.reshape(all_ret[k], list(sh[:-1]) + list(all_ret[k].shape[1:]))
    k_extract = ["rgb_map", "disp_map", "acc_map"]
    return [all_ret[k] for k in k_extract] + [{k: v for k, v in all_ret.items() if k not in k_extract}]


class _FrameRing:
    """Pinned host buffers + a copy stream: frame i's device->host copy runs while frame i+1 renders (the reference's
    render_path does `rgb.cpu().numpy()` per view, a synchronous pageable copy, run_nerf.py:196-199)."""

    def __init__(self, dev, shapes_dtypes, depth=2):
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.slots = [[torch.empty(sh, dtype=dt).pin_memory() for sh, dt in shapes_dtypes] for _ in range(depth)]
        self.events = [None] * depth
        self.dev = dev
        self.i = 0

    def push(self, tensors):
        """Start copying `tensors` (device) into the next slot; returns (slot index, the slot's host tensors)."""
        k = self.i % len(self.slots)
        self.i += 1
        if self.events[k] is not None:
            self.events[k].synchronize()                     # the slot's previous frame has been consumed by the caller by now
        main = torch.cuda.current_stream(self.dev)
        self.copy_stream.wait_stream(main)
        with torch.cuda.stream(self.copy_stream):
            for h, t in zip(self.slots[k], tensors):
                h.copy_(t, non_blocking=True)
                t.record_stream(self.copy_stream)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        self.events[k] = ev
        return k, self.slots[k]

    def wait(self, k):
        self.events[k].synchronize()


def render_path(render_poses, hwf, K, chunk, render_kwargs, gt_imgs=None, savedir=None, render_factor=0):
    """run_nerf.py:161-211: returns (rgbs [V,H,W,3], disps [V,H,W]) as float32 numpy arrays; with `savedir` every view is
    also written as 8-bit (PNG through imageio when it is installed, else .npy).  The 8-bit conversion (to8b) runs on the
    device and all device->host copies are asynchronous, double-buffered through pinned memory, so the copy and the file
    write of view i overlap the rendering of view i+1."""
    H, W, focal = hwf
    if render_factor != 0:
        H, W, focal = H // render_factor, W // render_factor, focal / render_factor
    n_views = len(render_poses)
    rgbs = np.empty((n_views, H, W, 3), dtype=np.float32)
    disps = np.empty((n_views, H, W), dtype=np.float32)
    if n_views == 0:
        return rgbs, disps
    dev = torch.device("cuda", torch.cuda.current_device())
    shapes = [((H, W, 3), torch.float32), ((H, W), torch.float32)] + ([((H, W, 3), torch.uint8)] if savedir is not None else [])
    ring = _FrameRing(dev, shapes)
    writer = None
    if savedir is not None:
        try:
            import imageio
            writer = lambda i, img: imageio.imwrite(os.path.join(savedir, "{:03d}.png".format(i)), img)
        except ImportError:
            writer = lambda i, img: np.save(os.path.join(savedir, "{:03d}.npy".format(i)), img)

    def drain(i, k, host):
        ring.wait(k)
        rgbs[i] = host[0].numpy()
        disps[i] = host[1].numpy()
        if writer is not None:
            writer(i, host[2].numpy())

    pending = None
    for i, c2w in enumerate(render_poses):
        with torch.no_grad():
            rgb, disp, acc, _ = render(H, W, K, chunk=chunk, c2w=c2w[:3, :4], **render_kwargs)
            frame = [rgb, disp] + ([ops.to8b(rgb)] if savedir is not None else [])
        k, host = ring.push(frame)
        if pending is not None:
            drain(*pending)                                   # view i-1 lands while view i renders
        pending = (i, k, host)
    drain(*pending)
    return rgbs, disps


def render_path_8bit(render_poses, hwf, K, chunk, render_kwargs, sink=None, first_view=0, view_count=None, render_factor=0):
    """The image-output leg of run_nerf.py:196-211 alone, for test-set sweeps: every view is rendered, converted to 8 bit on the
    device (to8b, run_nerf_helpers.py:14) and copied out asynchronously (3 bytes per pixel instead of 16); `sink(i, uint8[H,W,3])`
    is called once the frame is on the host while the next view renders.  Views [first_view, first_view + view_count) only --
    the view-major shard of one rank (distributed.shard_range).  Returns the number of views rendered."""
    H, W, focal = hwf
    if render_factor != 0:
        H, W, focal = H // render_factor, W // render_factor, focal / render_factor
    last = len(render_poses) if view_count is None else first_view + view_count
    if last <= first_view:
        return 0
    dev = torch.device("cuda", torch.cuda.current_device())
    ring = _FrameRing(dev, [((H, W, 3), torch.uint8)])
    pending = None
    for i in range(first_view, last):
        with torch.no_grad():
            rgb, _, _, _ = render(H, W, K, chunk=chunk, c2w=render_poses[i][:3, :4], **render_kwargs)
            img = ops.to8b(rgb)
        k, host = ring.push([img])
        if pending is not None:
            ring.wait(pending[1])
            if sink is not None:
                sink(pending[0], pending[2][0].numpy())
        pending = (i, k, host)
    ring.wait(pending[1])
    if sink is not None:
        sink(pending[0], pending[2][0].numpy())
    return last - first_view


def create_nerf(nerf_wrapper, multires=10, i_embed=0, use_viewdirs=True, multires_views=4, netchunk=1024 * 64, basedir=None,
                perturb=1., N_importance=128, N_samples=64, white_bkgd=False, raw_noise_std=0., dataset_type="blender",
                no_ndc=False, lindisp=False):
    """run_nerf.py:214-281: the render kwargs dictionaries for training and testing."""
    assert multires == 10 and multires_views == 4 and i_embed == 0 and use_viewdirs, \
        "the fused kernels implement PE L=10/4 with view directions"
    train = {"network_query_fn": _FusedQuery(), "perturb": perturb, "N_importance": N_importance,
             "network_fine": nerf_wrapper.model_fine, "N_samples": N_samples, "network_fn": nerf_wrapper.model,
             "use_viewdirs": use_viewdirs, "white_bkgd": white_bkgd, "raw_noise_std": raw_noise_std}
    if dataset_type != "llff" or no_ndc:
        train["ndc"] = False
        train["lindisp"] = lindisp
    test = dict(train)
    test["perturb"] = False
    test["raw_noise_std"] = 0.
    return train, test
