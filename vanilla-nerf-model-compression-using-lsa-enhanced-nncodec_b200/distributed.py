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


class PeerState:
    """Symmetric-memory region for the fused data-parallel finalize (include/nerfq.h, nerfq_mlp_backward_finalize_peers):
    every rank of the node maps every other rank's staging area + flag pad.  Built once per (device, group) and kept."""
    _cache = {}

    def __init__(self, dev, group):
        import torch.distributed._symmetric_memory as symm_mem
        from . import ops
        nbytes = ops.dp_peer_bytes()
        self.buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=dev)
        self.buf.zero_()
        torch.cuda.synchronize(dev)
        pg = group if group is not None else dist.group.WORLD
        try:
            self.handle = symm_mem.rendezvous(self.buf, pg)
        except TypeError:
            self.handle = symm_mem.rendezvous(self.buf, pg.group_name)
        self.world, self.rank = int(self.handle.world_size), int(self.handle.rank)
        ptrs = [int(p) for p in self.handle.buffer_ptrs]
        assert len(ptrs) == self.world and ptrs[self.rank] == self.buf.data_ptr()
        self.ptrs = torch.tensor(ptrs, dtype=torch.int64, device=dev)           # DEVICE array of the peers' base pointers
        self.epoch = torch.zeros(1, dtype=torch.int32, device=dev)
        torch.cuda.synchronize(dev)
        dist.barrier(group=group)            # nobody raises a flag in a pad that is not zeroed yet

    @classmethod
    def get(cls, dev, group):
        key = (dev.index if dev.index is not None else torch.cuda.current_device(), id(group))
        if key not in cls._cache:
            cls._cache[key] = cls(dev, group)
        return cls._cache[key]


def enable_data_parallel(enabled: bool = True, group=None, peer: Optional[bool] = None):
    """Switch the renderer's backward (render._backward_pipeline, used by autograd and by lsa.LSAStep) to data parallel.

    peer: None (default) -- use the fused peer-memory finalize when the ranks of the group can map each other's memory
    (torch symmetric memory: one node, NVLink / NVSwitch), otherwise one NCCL int64 all-reduce per step; True -- require it;
    False -- NCCL.  Environment NERFQ_DP_PEER=0 forces NCCL.  render.DATA_PARALLEL['collective'] names what is in use."""
    import os
    import sys
    from . import render
    on = bool(enabled) and dist.is_initialized() and dist.get_world_size(group) > 1
    render.DATA_PARALLEL["enabled"] = on
    render.DATA_PARALLEL["group"] = group
    render.DATA_PARALLEL["peer"] = None
    render.DATA_PARALLEL["collective"] = "nccl all_reduce (int64)" if on else None
    if on and peer is not False and os.environ.get("NERFQ_DP_PEER", "1") != "0" and torch.cuda.is_available():
        try:
            dev = torch.device("cuda", torch.cuda.current_device())
            render.DATA_PARALLEL["peer"] = PeerState.get(dev, group)
            render.DATA_PARALLEL["collective"] = "peer-memory finalize kernel (symmetric memory)"
        except Exception as ex:  # noqa: BLE001
            if peer:
                raise
            sys.stderr.write(f"nerfq: symmetric memory is not available ({type(ex).__name__}: {str(ex)[:200]}); data-parallel gradients use NCCL\n")
    return on


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
