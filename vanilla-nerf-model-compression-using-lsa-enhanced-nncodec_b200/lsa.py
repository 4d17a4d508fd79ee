"""The LSA fine-tuning step of the reference, as one replayable unit.

What the reference does per iteration (framework/pytorch_model/__init__.py:1077-1203 `tune_model` selects the
`weight_scaling` parameters and builds Adam; framework/nerf_model/run_nerf.py:730-757 is the iteration body):

    rgb, disp, acc, extras = render(H, W, K, chunk, rays=batch_rays, **render_kwargs_train)      # :739
    loss = img2mse(rgb, target_s) + img2mse(extras['rgb0'], target_s)                            # :741-751
    loss.backward(); optimizer.step()                                                            # :756-757

Here the same body runs through the fused kernels, and -- because a 3 ms step made of ~25 launches is bound by the
host's launch path once the caller reads the loss every iteration -- it can be captured once in a CUDA graph and
replayed: `LSAStep.capture()` records {optional requantise, render, loss, backward, Adam} on static input buffers,
`LSAStep.__call__` copies a batch into them (from pinned host memory or device memory) and replays.
"""
from typing import Callable, Optional

import torch

from . import codec, ops, packed
from . import render as R


def lsa_parameters(wrapper):
    """The parameters `tune_model` tunes with lsa_flag=True, ft_flag=False: param_types == 'weight.ls', i.e. the
    `weight_scaling` tensors (framework/pytorch_model/__init__.py:1131-1145).  Everything else is frozen."""
    out = []
    for name, prm in wrapper.named_parameters():
        is_ls = name.endswith("weight_scaling")
        prm.requires_grad_(is_ls)
        if is_ls:
            out.append(prm)
    return out


class PendingLoss:
    """The loss of an iteration enqueued by LSAStep.step_async: result() waits for its device-to-host copy."""

    def __init__(self, host: torch.Tensor, event: "torch.cuda.Event"):
        self._host, self._event = host, event

    def result(self) -> float:
        self._event.synchronize()
        return float(self._host[0])


class LSAStep:
    """One LSA iteration on a fixed ray-batch size.

    wrapper        NeRFWrapper with LSA parameters (model.LSA(...).add_lsa_params()), on a CUDA device
    n_rays         rays per call (per rank)
    lr             Adam learning rate (tune_model: torch.optim.Adam(tuning_params, lr=self.learning_rate))
    requantize     optional callable run at the start of every step (BASELINE cfg2 times quantise + render + update
                   together); None when the levels are frozen for the whole tuning run, as in tune_model
    H, W, K        image size and intrinsics, only needed when the rays go through the NDC warp (dataset_type='llff' without
                   no_ndc: render() then calls ndc_rays(H, W, K[0][0], 1., ...), run_nerf.py:131-133)
    render_kwargs  extra keyword arguments for render.create_nerf (perturb, white_bkgd, N_samples, ...)

    The 24 `weight_scaling` parameters are re-pointed at slices of ONE flat [2, 2436] tensor in the kernels' channel
    order (their values, names and state_dict entries are unchanged), so a step needs no gather of scales or scatter of
    gradients: the kernels read and write the flat buffers, Adam updates one tensor, and every parameter's `.grad` is a
    view of the flat gradient.  Launches per step: pack_rays, 2 x (set_scale_bias, MLP forward, composite), coarse depths,
    sample+merge, mse_grad, 2 x (composite backward, MLP backward, finalize), fused Adam, 2 memsets, the loss sum and the
    RNG draws of `perturb`.

    Data parallel (distributed.enable_data_parallel): the objective is the mean over the GLOBAL batch (world x n_rays);
    each rank's returned loss is its share of it (the sum over ranks is the global loss)."""

    def __init__(self, wrapper, n_rays: int, lr: float = 1e-4, requantize: Optional[Callable[[], None]] = None,
                 near: float = 2.0, far: float = 6.0, chunk: int = 32768, H: int = 4, W: int = 4, K=None, **render_kwargs):
        self.wrapper = wrapper
        self.n_rays = int(n_rays)
        self.requantize = requantize
        self.near, self.far, self.chunk = float(near), float(far), int(chunk)
        self.params = lsa_parameters(wrapper)
        dev = self.params[0].device
        if dev.type != "cuda":
            raise RuntimeError("LSAStep needs the model on a CUDA device (there is no CPU path)")
        self.device = dev
        self.train_kwargs, _ = R.create_nerf(wrapper, **render_kwargs)
        self.H, self.W, self.K = int(H), int(W), K
        self.ndc = bool(self.train_kwargs.get("ndc", True))
        if self.ndc and K is None:
            raise ValueError("this configuration renders through the NDC warp (create_nerf omits ndc=False for dataset_type='llff'): "
                             "pass H, W and the intrinsics K")
        self.flat = self._flatten_scales()
        self.flat_param = torch.nn.Parameter(self.flat)
        self.grad = torch.zeros_like(self.flat)
        self.flat_param.grad = self.grad
        for net_i, net in enumerate((wrapper.model, wrapper.model_fine)):
            for l, p, lo, n in self._slices(net):
                p.grad = self.grad[net_i, lo:lo + n].view(n, 1)
        # capturable: the step counter lives on the device, so optimizer.step() can sit inside a CUDA graph
        self.optimizer = torch.optim.Adam([self.flat_param], lr=lr, fused=True, capturable=True)
        self.loss2 = torch.zeros(2, device=dev)
        self.rays = torch.zeros(2, self.n_rays, 3, device=dev)         # [rays_o, rays_d] as run_nerf.py:739 passes batch_rays
        self.target = torch.zeros(self.n_rays, 3, device=dev)
        self.packed_rays = torch.zeros(self.n_rays, 11, device=dev)    # the iteration's input: rows [o, d, near, far, viewdir]
        self.graph = None
        self.loss = None
        self._h_loss = None

    @staticmethod
    def _slices(net):
        """(layer index, weight_scaling parameter, first channel, channels) in the kernels' flat channel order."""
        out, lo = [], 0
        scales = net.scale_tensors()
        for l in packed.CHANNEL_ORDER:
            n = packed.LAYER_OUT[l]
            out.append((l, scales[l], lo, n))
            lo += n
        return out

    def _flatten_scales(self) -> torch.Tensor:
        flat = torch.empty((2, packed.NUM_CHANNELS), dtype=torch.float32, device=self.device)
        with torch.no_grad():
            for net_i, net in enumerate((self.wrapper.model, self.wrapper.model_fine)):
                for l, p, lo, n in self._slices(net):
                    if p is None:
                        raise ValueError("LSAStep needs a model with LSA parameters (model.LSA(w).add_lsa_params())")
                    flat[net_i, lo:lo + n].copy_(p.detach().reshape(-1))
                    p.data = flat[net_i, lo:lo + n].view(n, 1)
        return flat

    def world(self) -> int:
        import torch.distributed as dist
        if R.DATA_PARALLEL["enabled"] and dist.is_initialized():
            return dist.get_world_size(R.DATA_PARALLEL.get("group"))
        return 1

    def _pack(self, rays: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """[2, n, 3] (rays_o, rays_d) -> packed rows [n, 11] (viewdirs, NDC warp, near / far), run_nerf.py:108-142."""
        focal = float(self.K[0][0]) if self.ndc else 1.0
        packed_rays = ops.pack_rays(rays[0], rays[1], self.ndc, self.H, self.W, focal, self.near, self.far)
        if out is not None:
            out.copy_(packed_rays)
            return out
        return packed_rays

    # the iteration body on packed rays
    def step_packed(self, packed_rays: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        if self.requantize is not None:
            self.requantize()
        kw = self.train_kwargs
        n_norm = self.n_rays * self.world()
        self.loss2.zero_()
        self.grad.zero_()
        with torch.no_grad():
            for i in range(0, self.n_rays, self.chunk):
                cfg = R._make_cfg(packed_rays[i:i + self.chunk], kw["network_fn"], kw["network_fine"], kw["N_samples"], kw["N_importance"],
                                  kw.get("lindisp", False), kw["perturb"], kw["white_bkgd"], kw["raw_noise_std"], False)
                outs, st = R._forward_pipeline(cfg, self.flat[0], self.flat[1], save=True)
                rgb, rgb0 = (outs[0], outs[3]) if cfg.Ni > 0 else (outs[0], None)
                _, d_rgb, d_rgb0 = ops.mse_grad(rgb, rgb0, target[i:i + self.chunk], n_norm=n_norm, loss2=self.loss2)
                if cfg.Ni > 0:
                    R._backward_pipeline(cfg, st, d_rgb, d_rgb0, g_out=self.grad)
                else:
                    R._backward_pipeline(cfg, st, None, d_rgb, g_out=self.grad)
            self.optimizer.step()
            return self.loss2.sum()

    # the iteration body, eager, from (rays_o, rays_d)
    def step(self, rays: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        return self.step_packed(self._pack(rays), target)

    def _init_optimizer_state(self):
        """Create Adam's state (step, exp_avg, exp_avg_sq) the way torch.optim.Adam does on its first step().  Inside a
        capture that lazy initialisation would be RECORDED, and every replay would start from a fresh optimizer; with the
        state created beforehand the graph only holds the update itself."""
        for group in self.optimizer.param_groups:
            for p in group["params"]:
                st = self.optimizer.state[p]
                if len(st) == 0:
                    st["step"] = torch.zeros((), dtype=torch.float32, device=p.device)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)

    def capture(self, warmup: int = 3):
        """Record the iteration in a CUDA graph (after `warmup` eager iterations on a side stream, which also update
        the parameters; 0 is allowed).  Raises if anything on the path is not capturable; the eager `step` stays usable."""
        self._init_optimizer_state()
        R._DPState.get(self.device)                       # data-parallel scratch exists before the capture
        ops._mse_workspace(self.device)
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.step_packed(self._pack(self.rays, self.packed_rays), self.target)
        torch.cuda.current_stream(self.device).wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self.loss = self.step_packed(self.packed_rays, self.target)
        self.graph = graph
        return self

    def __call__(self, rays: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        """rays [2, n_rays, 3] and target [n_rays, 3], on the host (pinned for an asynchronous copy) or on the device.
        Returns the loss as a device scalar; it is overwritten by the next call."""
        if self.graph is None:
            return self.step(rays.to(self.device, non_blocking=True), target.to(self.device, non_blocking=True))
        self.rays.copy_(rays, non_blocking=True)
        self.target.copy_(target, non_blocking=True)
        self._pack(self.rays, self.packed_rays)
        self.graph.replay()
        return self.loss

    def step_async(self, rays: torch.Tensor, target: torch.Tensor) -> "PendingLoss":
        """Enqueue one iteration (host or device batch, as __call__) plus an asynchronous copy of its loss to pinned host
        memory, and return at once.  `PendingLoss.result()` blocks until THAT iteration's loss has landed.  A loop that logs
        the loss every iteration keeps the GPU busy by reading iteration i's loss after enqueueing iteration i+1:

            pending = None
            for rays, target in batches:
                nxt = step.step_async(rays, target)
                if pending is not None: log(pending.result())
                pending = nxt

        (the reference's loop does loss.item() right after optimizer.step(): the host then idles for the whole iteration and
        the GPU for the host's launch work -- 0.25 ms of a 2.7 ms iteration here).  Two result slots: keep at most two
        iterations outstanding."""
        if self._h_loss is None:
            self._h_loss = [torch.zeros(1).pin_memory() for _ in range(2)]
            self._ev = [torch.cuda.Event() for _ in range(2)]
            self._slot = 0
        loss = self(rays, target)
        k = self._slot
        self._slot ^= 1
        self._h_loss[k].copy_(loss.reshape(1), non_blocking=True)
        self._ev[k].record(torch.cuda.current_stream(self.device))
        return PendingLoss(self._h_loss[k], self._ev[k])

    def step_selected(self, image: torch.Tensor, H: int, W: int, K, c2w, seed: int, step: int) -> torch.Tensor:
        """One iteration on a batch chosen ON THE DEVICE (run_nerf.py:690-735 draws np.random.choice(H*W, N_rand) and gathers
        rays and colours on the host every step): n_rays distinct pixels of `image` [H, W, 3] (device tensor), their rays from
        the pose `c2w` and their colours, by one kernel (ops.select_batch) straight into the step's input buffers, then the
        iteration itself (graph replay when captured).  No host random numbers, no host<->device traffic."""
        ops.select_batch(H, W, K, c2w, image, self.n_rays, seed, step, ndc=self.ndc, near=self.near, far=self.far,
                         rays_out=self.packed_rays, target_out=self.target)
        if self.graph is None:
            return self.step_packed(self.packed_rays, self.target)
        self.graph.replay()
        return self.loss


def make_requantizer(wrapper, master_state: dict, qp: int, qp_density: int = 2, nonweight_qp: int = -75):
    """The 'quantise' leg of BASELINE cfg2 as a graph-capturable callable: restore the unquantised float weights and
    biases from `master_state`, quantise + reconstruct them on the GPU (nnc_core/approximator/__init__.py:655-661)."""
    master = {k: v.detach().float().contiguous() for k, v in master_state.items() if k.endswith(".weight") or k.endswith(".bias")}

    def requantize():
        # quantise FROM the master copy INTO the parameters: no restore-copy of the 4.8 MB of floats per step
        codec.quantize_model(wrapper, qp, qp_density, nonweight_qp, master_state=master)
    return requantize
