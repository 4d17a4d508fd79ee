"""One Python function per C-ABI entry point (include/nerfq.h).  Tensors must be CUDA float32/int32 and
contiguous; every call is enqueued on torch's current stream."""
import ctypes
from typing import Optional, Sequence

import torch

from . import _lib
from .packed import PackedNet, _stream

_c = ctypes


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _f32(t: torch.Tensor, *shape) -> torch.Tensor:
    assert t.is_cuda and t.dtype == torch.float32, "expected a CUDA float32 tensor"
    if shape:
        assert tuple(t.shape) == tuple(shape), (tuple(t.shape), shape)
    return t.contiguous()


_PROTOS = {
    "nerfq_stepsize": (_c.c_int, [_c.c_int, _c.c_int, _c.POINTER(_c.c_float)]),
    "nerfq_quantize_urq": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_longlong, _c.c_int, _c.c_int, _c.c_void_p, _c.c_void_p, _c.c_void_p]),
    "nerfq_quantize_batch": (_c.c_int, [_c.POINTER(_c.c_void_p), _c.POINTER(_c.c_void_p), _c.POINTER(_c.c_void_p), _c.POINTER(_c.c_longlong),
                                        _c.POINTER(_c.c_int), _c.c_int, _c.c_int, _c.c_void_p, _c.c_void_p, _c.c_void_p]),
    "nerfq_dequantize": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_longlong, _c.c_int, _c.c_int, _c.c_void_p]),
    "nerfq_coarse_depths": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_longlong, _c.c_int, _c.c_int, _c.c_void_p, _c.c_void_p]),
    "nerfq_composite_fwd": (_c.c_int, [_c.c_void_p] * 4 + [_c.c_int, _c.c_longlong, _c.c_int] + [_c.c_void_p] * 6),
    "nerfq_composite_bwd": (_c.c_int, [_c.c_void_p] * 4 + [_c.c_int, _c.c_void_p, _c.c_longlong, _c.c_int, _c.c_void_p, _c.c_void_p]),
    "nerfq_sample_fine": (_c.c_int, [_c.c_void_p] * 4 + [_c.c_longlong, _c.c_int, _c.c_int] + [_c.c_void_p] * 4),
    "nerfq_camera_rays": (_c.c_int, [_c.c_int, _c.c_int, _c.POINTER(_c.c_float), _c.POINTER(_c.c_float), _c.c_int, _c.c_float,
                                     _c.c_float, _c.c_longlong, _c.c_longlong, _c.c_void_p, _c.c_void_p]),
    "nerfq_pack_rays": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_longlong, _c.c_int, _c.c_int, _c.c_int, _c.c_float, _c.c_float,
                                   _c.c_float, _c.c_void_p, _c.c_void_p]),
    "nerfq_mse_grad": (_c.c_int, [_c.c_void_p] * 3 + [_c.c_longlong, _c.c_longlong] + [_c.c_void_p] * 5),
    "nerfq_mse_grad_workspace_bytes": (_c.c_ulonglong, []),
    "nerfq_mlp_backward": (_c.c_int, [_c.c_void_p] * 4 + [_c.c_longlong, _c.c_void_p, _c.c_int, _c.c_void_p]),
    "nerfq_mlp_backward_partial": (_c.c_int, [_c.c_void_p] * 4 + [_c.c_longlong, _c.c_void_p, _c.c_int, _c.c_void_p]),
    "nerfq_mlp_backward_finalize": (_c.c_int, [_c.c_void_p] * 4),
    "nerfq_mlp_grad_fix_bytes": (_c.c_ulonglong, []),
    "nerfq_dp_peer_bytes": (_c.c_ulonglong, []),
    "nerfq_mlp_backward_finalize_peers": (_c.c_int, [_c.c_void_p] * 4 + [_c.c_int, _c.c_int, _c.c_void_p, _c.c_void_p, _c.c_void_p]),
    "nerfq_render_rays_workspace_bytes": (_c.c_ulonglong, [_c.c_longlong, _c.c_int, _c.c_int]),
    "nerfq_render_rays_fwd": (_c.c_int, [_c.c_void_p] * 3 + [_c.c_longlong, _c.c_int, _c.c_int, _c.c_int, _c.c_int] + [_c.c_void_p] * 8 +
                              [_c.c_int, _c.c_void_p]),
    "nerfq_to8b": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_longlong, _c.c_void_p]),
    "nerfq_select_batch": (_c.c_int, [_c.c_int, _c.c_int, _c.POINTER(_c.c_float), _c.POINTER(_c.c_float), _c.c_int, _c.c_float, _c.c_float,
                                      _c.c_void_p, _c.c_ulonglong, _c.c_ulonglong, _c.c_longlong, _c.c_void_p, _c.c_void_p, _c.c_void_p,
                                      _c.c_void_p]),
}
_bound = False


def L():
    global _bound
    lib = _lib.lib()
    if not _bound:
        for name, (res, args) in _PROTOS.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _bound = True
    return lib


# ---- quantiser (nnc_core/approximator/baseline.py, nnc_core/common.py:28-46) ---------------------
def stepsize(qp: int, qp_density: int) -> float:
    out = _c.c_float()
    _lib.check(L().nerfq_stepsize(int(qp), int(qp_density), _c.byref(out)), "nerfq_stepsize")
    return float(out.value)


def quantize_urq(w: torch.Tensor, qp: int, qp_density: int):
    """float32 tensor -> (int32 levels, qp actually used as a 1-element CUDA int32 tensor)."""
    w = _f32(w)
    lvl = torch.empty(w.shape, dtype=torch.int32, device=w.device)
    ws = torch.empty(2, dtype=torch.int32, device=w.device)
    _lib.check(L().nerfq_quantize_urq(w.data_ptr(), lvl.data_ptr(), w.numel(), int(qp), int(qp_density), ws[1:].data_ptr(),
                                      ws.data_ptr(), _stream()), "nerfq_quantize_urq")
    return lvl, ws[1:]


def quantize_batch(tensors: Sequence[torch.Tensor], qps: Sequence[int], qp_density: int, reconstruct_in_place: bool = False,
                   levels_out: Optional[Sequence[torch.Tensor]] = None, reconstruct_into: Optional[Sequence[torch.Tensor]] = None):
    """Quantise many float32 tensors with one launch pair: returns (list of int32 level tensors, qp_used int32[T]).
    reconstruct_in_place overwrites each input with level*delta (what `rec(approx(x))` yields in the reference);
    reconstruct_into writes level*delta into other tensors of the same shapes instead (the inputs stay untouched)."""
    t = len(tensors)
    assert t == len(qps) and 0 < t <= 64
    dev = tensors[0].device
    for x in tensors:
        assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()
    lv = [levels_out[i] if (levels_out is not None and levels_out[i] is not None) else torch.empty(x.shape, dtype=torch.int32, device=dev)
          for i, x in enumerate(tensors)]
    ws = torch.empty(2 * t, dtype=torch.int32, device=dev)
    vp = _c.c_void_p * t
    w_arr = vp(*[x.data_ptr() for x in tensors])
    l_arr = vp(*[x.data_ptr() for x in lv])
    r_arr = None
    if reconstruct_into is not None:
        assert len(reconstruct_into) == t
        for x, r in zip(tensors, reconstruct_into):
            assert r.is_cuda and r.dtype == torch.float32 and r.is_contiguous() and r.numel() == x.numel()
        r_arr = vp(*[r.data_ptr() for r in reconstruct_into])
    elif reconstruct_in_place:
        r_arr = vp(*[x.data_ptr() for x in tensors])
    n_arr = (_c.c_longlong * t)(*[x.numel() for x in tensors])
    q_arr = (_c.c_int * t)(*[int(q) for q in qps])
    _lib.check(L().nerfq_quantize_batch(w_arr, l_arr, r_arr, n_arr, q_arr, t, int(qp_density), ws[t:].data_ptr(), ws.data_ptr(), _stream()),
               "nerfq_quantize_batch")
    return lv, ws[t:]


def dequantize(lvl: torch.Tensor, qp: int, qp_density: int) -> torch.Tensor:
    assert lvl.is_cuda and lvl.dtype == torch.int32
    lvl = lvl.contiguous()
    w = torch.empty(lvl.shape, dtype=torch.float32, device=lvl.device)
    _lib.check(L().nerfq_dequantize(lvl.data_ptr(), w.data_ptr(), lvl.numel(), int(qp), int(qp_density), _stream()), "nerfq_dequantize")
    return w


# ---- render path ---------------------------------------------------------------------------------
def coarse_depths(rays: torch.Tensor, n_samples: int, lindisp: bool = False, t_rand: Optional[torch.Tensor] = None) -> torch.Tensor:
    rays = _f32(rays)
    n = rays.shape[0]
    z = torch.empty((n, n_samples), dtype=torch.float32, device=rays.device)
    if t_rand is not None:
        t_rand = _f32(t_rand, n, n_samples)
    _lib.check(L().nerfq_coarse_depths(rays.data_ptr(), _p(t_rand), n, n_samples, int(lindisp), z.data_ptr(), _stream()), "nerfq_coarse_depths")
    return z


def composite_fwd(raw: torch.Tensor, z: torch.Tensor, rays: torch.Tensor, white_bkgd: bool, noise: Optional[torch.Tensor] = None,
                  want_weights: bool = True):
    n, s = z.shape
    raw, z, rays = _f32(raw, n, s, 4), _f32(z), _f32(rays, n, 11)
    dev = z.device
    rgb = torch.empty((n, 3), dtype=torch.float32, device=dev)
    disp = torch.empty((n,), dtype=torch.float32, device=dev)
    acc = torch.empty((n,), dtype=torch.float32, device=dev)
    depth = torch.empty((n,), dtype=torch.float32, device=dev)
    weights = torch.empty((n, s), dtype=torch.float32, device=dev) if want_weights else None
    if noise is not None:
        noise = _f32(noise, n, s)
    _lib.check(L().nerfq_composite_fwd(raw.data_ptr(), z.data_ptr(), rays.data_ptr(), _p(noise), int(white_bkgd), n, s, rgb.data_ptr(),
                                       disp.data_ptr(), acc.data_ptr(), depth.data_ptr(), _p(weights), _stream()), "nerfq_composite_fwd")
    return rgb, disp, acc, weights, depth


def composite_bwd(raw, z, rays, white_bkgd: bool, d_rgb: torch.Tensor, noise: Optional[torch.Tensor] = None) -> torch.Tensor:
    n, s = z.shape
    raw, z, rays, d_rgb = _f32(raw, n, s, 4), _f32(z), _f32(rays, n, 11), _f32(d_rgb, n, 3)
    d_raw = torch.empty((n, s, 4), dtype=torch.float32, device=z.device)
    if noise is not None:
        noise = _f32(noise, n, s)
    _lib.check(L().nerfq_composite_bwd(raw.data_ptr(), z.data_ptr(), rays.data_ptr(), _p(noise), int(white_bkgd), d_rgb.data_ptr(), n, s,
                                       d_raw.data_ptr(), _stream()), "nerfq_composite_bwd")
    return d_raw


def sample_fine(z_coarse: torch.Tensor, weights: torch.Tensor, n_importance: int, u: Optional[torch.Tensor] = None,
                want_samples: bool = False):
    """sample_pdf on the S-2 interior weights + merge with the coarse depths: returns (z_all sorted [N,S+Ni], z_std[N], z_samples|None)."""
    n, s = z_coarse.shape
    z_coarse, weights = _f32(z_coarse), _f32(weights, n, s)
    dev = z_coarse.device
    z_all = torch.empty((n, s + n_importance), dtype=torch.float32, device=dev)
    z_std = torch.empty((n,), dtype=torch.float32, device=dev)
    zs = torch.empty((n, n_importance), dtype=torch.float32, device=dev) if want_samples else None
    if u is not None:
        u = _f32(u, n, n_importance)
    _lib.check(L().nerfq_sample_fine(z_coarse.data_ptr(), None, weights.data_ptr(), _p(u), n, s, n_importance, z_all.data_ptr(),
                                     z_std.data_ptr(), _p(zs), _stream()), "nerfq_sample_fine")
    return z_all, z_std, zs


def sample_pdf(bins: torch.Tensor, weights: torch.Tensor, n_samples: int, u: Optional[torch.Tensor] = None) -> torch.Tensor:
    """run_nerf_helpers.py:119-163 as a function: bins [N,B], weights [N,B-1] -> samples [N,n_samples]."""
    n, nb = bins.shape
    bins, weights = _f32(bins), _f32(weights, n, nb - 1)
    w = torch.zeros((n, nb + 1), dtype=torch.float32, device=bins.device)
    w[:, 1:-1] = weights
    zs = torch.empty((n, n_samples), dtype=torch.float32, device=bins.device)
    if u is not None:
        u = _f32(u, n, n_samples)
    _lib.check(L().nerfq_sample_fine(None, bins.data_ptr(), w.data_ptr(), _p(u), n, nb + 1, n_samples, None, None, zs.data_ptr(),
                                     _stream()), "nerfq_sample_fine")
    return zs


def camera_rays(H: int, W: int, K, c2w, ndc: bool, near: float, far: float, device, first_pixel: int = 0,
                count: Optional[int] = None) -> torch.Tensor:
    """Packed ray rows [count, 11] for pixels [first, first+count) of an HxW pinhole image."""
    count = H * W - first_pixel if count is None else count
    k4 = (_c.c_float * 4)(float(K[0][0]), float(K[1][1]), float(K[0][2]), float(K[1][2]))
    c = [float(c2w[r][col]) for r in range(3) for col in range(4)]
    c12 = (_c.c_float * 12)(*c)
    out = torch.empty((count, 11), dtype=torch.float32, device=device)
    _lib.check(L().nerfq_camera_rays(H, W, k4, c12, int(ndc), float(near), float(far), first_pixel, count, out.data_ptr(), _stream()),
               "nerfq_camera_rays")
    return out


def pack_rays(rays_o: torch.Tensor, rays_d: torch.Tensor, ndc: bool, H: int, W: int, focal: float, near: float, far: float) -> torch.Tensor:
    rays_o, rays_d = _f32(rays_o.reshape(-1, 3)), _f32(rays_d.reshape(-1, 3))
    n = rays_o.shape[0]
    out = torch.empty((n, 11), dtype=torch.float32, device=rays_o.device)
    _lib.check(L().nerfq_pack_rays(rays_o.data_ptr(), rays_d.data_ptr(), n, int(ndc), int(H), int(W), float(focal), float(near), float(far),
                                   out.data_ptr(), _stream()), "nerfq_pack_rays")
    return out


_MSE_WS = {}


def _mse_workspace(dev) -> torch.Tensor:
    """Per-device scratch of nerfq_mse_grad (zeroed once; the kernel re-arms it).  Calls are stream-ordered."""
    key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device())
    if key not in _MSE_WS:
        _MSE_WS[key] = torch.zeros(int(L().nerfq_mse_grad_workspace_bytes()), dtype=torch.uint8, device=dev)
    return _MSE_WS[key]


def mse_grad(rgb: torch.Tensor, rgb0: Optional[torch.Tensor], target: torch.Tensor, n_norm: int = 0,
             loss2: Optional[torch.Tensor] = None):
    """(loss2 = [mse(rgb,t), mse(rgb0,t)], d_rgb, d_rgb0) for loss = mse(rgb,t) + mse(rgb0,t).  n_norm: rays the mean runs
    over (0 = this batch; a data-parallel rank passes the global batch size).  loss2: optional zeroed [2] output."""
    n = rgb.shape[0]
    rgb, target = _f32(rgb, n, 3), _f32(target, n, 3)
    d_rgb = torch.empty_like(rgb)
    d_rgb0 = None
    if rgb0 is not None:
        rgb0 = _f32(rgb0, n, 3)
        d_rgb0 = torch.empty_like(rgb0)
    if loss2 is None:
        loss2 = torch.zeros(2, dtype=torch.float32, device=rgb.device)
    ws = _mse_workspace(rgb.device)
    _lib.check(L().nerfq_mse_grad(rgb.data_ptr(), _p(rgb0), target.data_ptr(), n, int(n_norm), d_rgb.data_ptr(), _p(d_rgb0),
                                  loss2.data_ptr(), ws.data_ptr(), _stream()), "nerfq_mse_grad")
    return loss2, d_rgb, d_rgb0


def mlp_backward(net: PackedNet, d_raw: torch.Tensor, raw: torch.Tensor, save: torch.Tensor, d_scale: Optional[torch.Tensor] = None,
                 max_ctas: int = 0) -> torch.Tensor:
    """Accumulates d loss / d lsa_scale (flat [2436], kernel channel order) for one network."""
    n_points = raw.shape[0] * raw.shape[1]
    d_raw, raw = _f32(d_raw, *raw.shape), _f32(raw)
    if d_scale is None:
        d_scale = torch.zeros(2436, dtype=torch.float32, device=raw.device)
    _lib.check(L().nerfq_mlp_backward(net.ptr, d_raw.data_ptr(), raw.data_ptr(), save.data_ptr(), n_points, d_scale.data_ptr(), max_ctas,
                                      _stream()), "nerfq_mlp_backward")
    return d_scale


def grad_fix_elems() -> int:
    """int64 elements of one network's fixed-point scale-gradient buffer."""
    return int(L().nerfq_mlp_grad_fix_bytes()) // 8


def mlp_backward_partial(net: PackedNet, d_raw: torch.Tensor, raw: torch.Tensor, save: torch.Tensor, grad_fix: torch.Tensor,
                         max_ctas: int = 0) -> torch.Tensor:
    """The backward kernel alone: accumulates s*ds per channel into grad_fix (int64 [grad_fix_elems()], value * 2^48)."""
    n_points = raw.shape[0] * raw.shape[1]
    d_raw, raw = _f32(d_raw, *raw.shape), _f32(raw)
    assert grad_fix.is_cuda and grad_fix.dtype == torch.int64 and grad_fix.is_contiguous() and grad_fix.numel() >= grad_fix_elems()
    _lib.check(L().nerfq_mlp_backward_partial(net.ptr, d_raw.data_ptr(), raw.data_ptr(), save.data_ptr(), n_points, grad_fix.data_ptr(),
                                              max_ctas, _stream()), "nerfq_mlp_backward_partial")
    return grad_fix


def mlp_backward_finalize(net: PackedNet, grad_fix: torch.Tensor, d_scale: torch.Tensor) -> torch.Tensor:
    """d_scale[2436] += grad_fix / 2^48 / scale; grad_fix is left zeroed."""
    assert d_scale.is_cuda and d_scale.dtype == torch.float32 and d_scale.is_contiguous() and d_scale.numel() == 2436
    _lib.check(L().nerfq_mlp_backward_finalize(net.ptr, grad_fix.data_ptr(), d_scale.data_ptr(), _stream()), "nerfq_mlp_backward_finalize")
    return d_scale


def dp_peer_bytes() -> int:
    return int(L().nerfq_dp_peer_bytes())


def mlp_backward_finalize_peers(net0: PackedNet, net1: Optional[PackedNet], grad_fix2: torch.Tensor, peers_dev: int, world: int, rank: int,
                                epoch: torch.Tensor, d_scale2: torch.Tensor) -> torch.Tensor:
    """Data-parallel finalize with the all-reduce fused in (nerfq_mlp_backward_finalize_peers): grad_fix2 int64 [2, 2440] (coarse,
    fine) is published to the peers, all ranks' sums are added up and d_scale2 float32 [2, 2436] accumulates the gradients."""
    assert grad_fix2.dtype == torch.int64 and grad_fix2.is_contiguous() and grad_fix2.numel() == 2 * grad_fix_elems()
    assert d_scale2.dtype == torch.float32 and d_scale2.is_contiguous() and d_scale2.numel() == 2 * 2436
    assert epoch.dtype == torch.int32 and epoch.is_cuda
    _lib.check(L().nerfq_mlp_backward_finalize_peers(net0.ptr, net1.ptr if net1 is not None else None, grad_fix2.data_ptr(), peers_dev, int(world),
                                                     int(rank), epoch.data_ptr(), d_scale2.data_ptr(), _stream()), "nerfq_mlp_backward_finalize_peers")
    return d_scale2


# ---- one-call forward render, image output, batch selection ---------------------------------------------------
_RENDER_WS = {}


def render_rays_fwd(net0: PackedNet, net1: Optional[PackedNet], rays: torch.Tensor, n_samples: int, n_importance: int, lindisp: bool,
                    white_bkgd: bool, max_ctas: int = 0):
    """nerfq_render_rays_fwd: (rgb, disp, acc, rgb0, disp0, acc0, z_std) for rays [N,11]; the last four are None when
    n_importance == 0.  The workspace is kept per (device, size) and reused by later calls on the same stream order."""
    rays = _f32(rays)
    n, dev = rays.shape[0], rays.device
    f = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)
    rgb, disp, acc = f(n, 3), f(n), f(n)
    rgb0 = disp0 = acc0 = z_std = None
    if n_importance > 0:
        rgb0, disp0, acc0, z_std = f(n, 3), f(n), f(n), f(n)
    if n == 0:
        return rgb, disp, acc, rgb0, disp0, acc0, z_std
    need = int(L().nerfq_render_rays_workspace_bytes(n, int(n_samples), int(n_importance)))
    key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device())
    ws = _RENDER_WS.get(key)
    if ws is None or ws.numel() < need:
        ws = _RENDER_WS[key] = torch.empty(need, dtype=torch.uint8, device=dev)
    _lib.check(L().nerfq_render_rays_fwd(net0.ptr, net1.ptr if net1 is not None else None, rays.data_ptr(), n, int(n_samples), int(n_importance),
                                         int(lindisp), int(white_bkgd), ws.data_ptr(), rgb.data_ptr(), disp.data_ptr(), acc.data_ptr(),
                                         _p(rgb0), _p(disp0), _p(acc0), _p(z_std), max_ctas, _stream()), "nerfq_render_rays_fwd")
    return rgb, disp, acc, rgb0, disp0, acc0, z_std


def to8b(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(255 * clip(x, 0, 1)).astype(uint8) on the device (run_nerf_helpers.py:14)."""
    x = _f32(x)
    if out is None:
        out = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    assert out.is_cuda and out.dtype == torch.uint8 and out.is_contiguous() and out.numel() == x.numel()
    _lib.check(L().nerfq_to8b(x.data_ptr(), out.data_ptr(), x.numel(), _stream()), "nerfq_to8b")
    return out


def select_batch(H: int, W: int, K, c2w, image: Optional[torch.Tensor], n: int, seed: int, step: int, ndc: bool = False, near: float = 0.,
                 far: float = 1., device=None, want_index: bool = False, rays_out: Optional[torch.Tensor] = None,
                 target_out: Optional[torch.Tensor] = None):
    """n distinct pixels of an image: (packed rays [n,11], target colours [n,3] | None, pixel indices | None)."""
    dev = image.device if image is not None else device
    k4 = (_c.c_float * 4)(float(K[0][0]), float(K[1][1]), float(K[0][2]), float(K[1][2]))
    c12 = (_c.c_float * 12)(*[float(c2w[r][col]) for r in range(3) for col in range(4)])
    rays = rays_out if rays_out is not None else torch.empty((n, 11), dtype=torch.float32, device=dev)
    target = None
    if image is not None:
        image = _f32(image.reshape(-1, 3), H * W, 3)
        target = target_out if target_out is not None else torch.empty((n, 3), dtype=torch.float32, device=dev)
    idx = torch.empty((n,), dtype=torch.int32, device=dev) if want_index else None
    _lib.check(L().nerfq_select_batch(int(H), int(W), k4, c12, int(ndc), float(near), float(far), _p(image), int(seed) & (2 ** 64 - 1),
                                      int(step) & (2 ** 64 - 1), int(n), rays.data_ptr(), _p(target), _p(idx), _stream()), "nerfq_select_batch")
    return rays, target, idx
