"""Multi-GPU plumbing (one process per GPU, torch.distributed over NCCL): rays are independent units, so
test-view rendering shards the pixel range with no data-path collective, and data-parallel LSA tuning needs
one all-reduce of the scale gradients per step (39 KB of 64-bit fixed-point sums, both networks).  The reference itself is single-GPU
(README.md:76); SURVEY section 8(e) defines this sharding."""
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [first, first+count) slice of n units for `rank`; remainders go to the lowest ranks."""
    base, rem = divmod(n, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def allreduce_fixed(fix: torch.Tensor, group=None) -> torch.Tensor:
    """Sum the ranks' fixed-point scale-gradient buffers (int64, value * 2^48; nerfq_mlp_backward_partial) in place.
    Integer addition is associative, so the sum -- and the float gradient nerfq_mlp_backward_finalize derives from it --
    does not depend on the number of ranks or on NCCL's reduction order: N ranks x B rays == 1 rank x N*B rays, bit for
    bit, when every rank weights its rays by 1 / (global batch)."""
    assert fix.dtype == torch.int64
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(fix, op=dist.ReduceOp.SUM, group=group)
    return fix


def enable_data_parallel(enabled: bool = True, group=None):
    """Switch the renderer's backward (render._backward_pipeline, used by autograd and by lsa.LSAStep) to data parallel."""
    from . import render
    render.DATA_PARALLEL["enabled"] = bool(enabled) and dist.is_initialized() and dist.get_world_size(group) > 1
    render.DATA_PARALLEL["group"] = group
    return render.DATA_PARALLEL["enabled"]


def render_view_sharded(H: int, W: int, K, c2w, render_kwargs: dict, chunk: int = 32768, ndc: bool = False, near: float = 2.0,
                        far: float = 6.0, gather: bool = True, group=None):
    """Render rank's contiguous slice of the H*W pixels of one view; optionally all-gather the image.

    Returns (rgb, disp, acc): full [H,W,*] tensors when gather=True, else the local [count,*] slices."""
    from . import ops
    from .render import batchify_rays
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    first, count = shard_range(H * W, rank, world)
    dev = torch.device("cuda", torch.cuda.current_device())
    c = c2w.detach().cpu().numpy() if torch.is_tensor(c2w) else c2w
    rays = ops.camera_rays(H, W, K, c, ndc, near, far, dev, first_pixel=first, count=count)
    kw = {k: v for k, v in render_kwargs.items() if k not in ("network_query_fn", "use_viewdirs", "ndc", "near", "far")}
    with torch.no_grad():
        ret = batchify_rays(rays, chunk, **kw)
    local = (ret["rgb_map"], ret["disp_map"], ret["acc_map"])
    if not gather or world == 1:
        if gather:
            return local[0].reshape(H, W, 3), local[1].reshape(H, W), local[2].reshape(H, W)
        return local
    out = []
    maxc = shard_range(H * W, 0, world)[1]
    for t in local:
        pad = torch.zeros((maxc,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[:count] = t
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad, group=group)
        full = torch.cat([parts[r][:shard_range(H * W, r, world)[1]] for r in range(world)], 0)
        out.append(full.reshape((H, W) + tuple(t.shape[1:])))
    return tuple(out)
