"""Quantise / reconstruct a NeRFWrapper the way the reference's approximator does around LSA tuning.

  quantize_model    nnc_core/approximator/__init__.py:603-661 (run_ft_and_lsa: approx + rec before tuning),
                    per-tensor qp assignment :771-780 (weights qp, everything else nonweight_qp = -75)
  apply_lsa         nnc_core/approximator/__init__.py:276-318 (decoder side: w *= ls, drop the scales)

The integer levels stay on the GPU and become the tensor-core operands of the fused MLP (model.NeRF.quant_levels);
the float `weight` parameters are set to level*delta so `state_dict()` matches what the reference's `rec` yields.
Only uniform reconstruction quantisation (use_dq=False) runs on the GPU; dependent (trellis) quantisation is the
reference's default and lives in deepCABAC, which is absent here (parity unpinned).
"""
from typing import Dict

import torch

from . import ops
from .model import NeRF


@torch.no_grad()
def quantize_net(net: NeRF, qp: int, qp_density: int = 2, nonweight_qp: int = -75) -> Dict[str, torch.Tensor]:
    levels, steps, out = [], [], {}
    for i, layer in enumerate(net.layers()):
        lv, used = ops.quantize_urq(layer.weight.detach().float(), qp, qp_density)
        levels.append(lv)
        steps.append(ops.stepsize(qp, qp_density))
        layer.weight.copy_(ops.dequantize(lv, qp, qp_density))
        lb, _ = ops.quantize_urq(layer.bias.detach().float(), nonweight_qp, qp_density)
        layer.bias.copy_(ops.dequantize(lb, nonweight_qp, qp_density))
        out[f"{i}.weight"], out[f"{i}.bias"] = lv, lb
    net.quant_levels, net.quant_steps = levels, steps
    return out


@torch.no_grad()
def quantize_model(wrapper, qp: int, qp_density: int = 2, nonweight_qp: int = -75):
    """Quantise + reconstruct both networks of a NeRFWrapper in place; returns {net: {tensor: int32 levels}}."""
    return {"model": quantize_net(wrapper.model, qp, qp_density, nonweight_qp),
            "model_fine": quantize_net(wrapper.model_fine, qp, qp_density, nonweight_qp)}


@torch.no_grad()
def quantize_scales(wrapper, qp_density: int = 2, nonweight_qp: int = -75):
    """Final pass of nnc/compression.py:534-538 for the tuned scales: ls -> dequant(quant(ls, -75))."""
    for net in (wrapper.model, wrapper.model_fine):
        for s in net.scale_tensors():
            if s is not None:
                lv, _ = ops.quantize_urq(s.detach().float(), nonweight_qp, qp_density)
                s.copy_(ops.dequantize(lv, nonweight_qp, qp_density))
