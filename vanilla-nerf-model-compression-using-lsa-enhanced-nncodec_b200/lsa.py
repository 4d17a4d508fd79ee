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

from . import codec
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
    """

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
        # capturable: the step counters live on the device, so optimizer.step() can sit inside a CUDA graph
        self.optimizer = torch.optim.Adam(self.params, lr=lr, fused=True, capturable=True)
        self.train_kwargs, _ = R.create_nerf(wrapper, **render_kwargs)
        self.H, self.W, self.K = int(H), int(W), K
        if self.train_kwargs.get("ndc", True) and K is None:
            raise ValueError("this configuration renders through the NDC warp (create_nerf omits ndc=False for dataset_type='llff'): "
                             "pass H, W and the intrinsics K")
        self.rays = torch.zeros(2, self.n_rays, 3, device=dev)         # [rays_o, rays_d] as run_nerf.py:739 passes batch_rays
        self.target = torch.zeros(self.n_rays, 3, device=dev)
        self.graph = None
        self.loss = None

    # the iteration body, eager
    def step(self, rays: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        if self.requantize is not None:
            self.requantize()
        rgb, _, _, extras = R.render(self.H, self.W, self.K, chunk=self.chunk, rays=(rays[0], rays[1]), near=self.near, far=self.far,
                                     retraw=False, **self.train_kwargs)
        loss = R.img2mse(rgb, target) + R.img2mse(extras["rgb0"], target)
        loss.backward()
        self.optimizer.step()
        return loss.detach()

    def _init_optimizer_state(self):
        """Create Adam's per-parameter state (step, exp_avg, exp_avg_sq) the way torch.optim.Adam does on its first
        step().  Inside a capture that lazy initialisation would be RECORDED, and every replay would start from a fresh
        optimizer; with the state created beforehand the graph only holds the update itself."""
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
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.optimizer.zero_grad(set_to_none=True)
                self.step(self.rays, self.target)
        torch.cuda.current_stream(self.device).wait_stream(side)
        self.optimizer.zero_grad(set_to_none=True)        # backward() then allocates .grad inside the graph's pool
        graph = torch.cuda.CUDAGraph()
        # The parameters' AccumulateGrad nodes were created on whatever stream first ran a backward; the capture stream
        # differs from it by construction, which autograd reports once per process.  The capture orders the streams itself.
        quiet = getattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch", None)
        if quiet is not None:
            quiet(False)
        try:
            with torch.cuda.graph(graph):
                self.loss = self.step(self.rays, self.target)
        finally:
            if quiet is not None:
                quiet(True)
        self.graph = graph
        return self

    def __call__(self, rays: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        """rays [2, n_rays, 3] and target [n_rays, 3], on the host (pinned for an asynchronous copy) or on the device.
        Returns the loss as a device scalar; it is overwritten by the next call."""
        if self.graph is None:
            self.optimizer.zero_grad(set_to_none=True)
            return self.step(rays.to(self.device, non_blocking=True), target.to(self.device, non_blocking=True))
        self.rays.copy_(rays, non_blocking=True)
        self.target.copy_(target, non_blocking=True)
        self.graph.replay()
        return self.loss


def make_requantizer(wrapper, master_state: dict, qp: int, qp_density: int = 2, nonweight_qp: int = -75):
    """The 'quantise' leg of BASELINE cfg2 as a graph-capturable callable: restore the unquantised float weights and
    biases from `master_state`, quantise + reconstruct them on the GPU (nnc_core/approximator/__init__.py:655-661)."""
    sd = wrapper.state_dict()
    keys = [k for k in master_state if k.endswith(".weight") or k.endswith(".bias")]
    dst = [sd[k] for k in keys]
    src = [master_state[k] for k in keys]

    def requantize():
        with torch.no_grad():
            torch._foreach_copy_(dst, src)
        codec.quantize_model(wrapper, qp, qp_density, nonweight_qp)
    return requantize
