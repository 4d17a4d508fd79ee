"""Quantise / reconstruct a NeRFWrapper the way the reference's approximator does around LSA tuning.

  quantize_model    nnc_core/approximator/__init__.py:603-661 (run_ft_and_lsa: approx + rec before tuning),
                    per-tensor qp assignment :771-780 (weights qp, everything else nonweight_qp = -75)
  apply_lsa         nnc_core/approximator/__init__.py:276-318 (decoder side: w *= ls, drop the scales)

The integer levels stay on the GPU and become the tensor-core operands of the fused MLP (model.NeRF.quant_levels);
the float `weight` parameters are set to level*delta so `state_dict()` matches what the reference's `rec` yields.
Only uniform reconstruction quantisation (use_dq=False) runs on the GPU; dependent (trellis) quantisation is the
reference's default and lives in deepCABAC, which is absent here (parity unpinned).
"""
from typing import Dict, Optional

import torch

from . import ops, packed
from .model import NeRF


def _quantize_nets(nets, qp: int, qp_density: int, nonweight_qp: int, sources=None):
    """One launch pair for every weight and bias of `nets`: levels out, parameters overwritten with level*delta.
    sources: optional {id(parameter tensor): float32 tensor to quantise instead of the parameter's current value} -- the
    unquantised master copy; the parameters then only receive the reconstruction."""
    tensors, qps, reuse = [], [], []
    for net in nets:
        # a requantisation writes into the level tensors the network already holds (no new allocation: the python attributes
        # and every captured graph keep referring to the same memory); the first quantisation allocates them
        prev = net.quant_levels if (net.quant_levels is not None and net._quant_key is not None) else None
        for i, layer in enumerate(net.layers()):
            tensors += [layer.weight.data, layer.bias.data]
            qps += [qp, nonweight_qp]
            ok = prev is not None and prev[i].dtype == torch.int32 and prev[i].shape == layer.weight.shape and prev[i].is_contiguous()
            reuse += [prev[i] if ok else None, None]
    for x in tensors:
        assert x.dtype == torch.float32 and x.is_contiguous()
    srcs = tensors if sources is None else [sources[x.data_ptr()] for x in tensors]
    outs = []
    for i in range(0, len(tensors), 64):
        lv, _ = ops.quantize_batch(srcs[i:i + 64], qps[i:i + 64], qp_density, levels_out=reuse[i:i + 64], reconstruct_into=tensors[i:i + 64])
        outs += lv
    step = ops.stepsize(qp, qp_density)
    res, k = [], 0
    for net in nets:
        out, levels = {}, []
        for i, layer in enumerate(net.layers()):
            out[f"{i}.weight"], out[f"{i}.bias"] = outs[k], outs[k + 1]
            levels.append(outs[k])
            k += 2
        # (the kernel wrote level*delta through raw pointers: torch's version counters did not move, so the key taken
        # here describes exactly the reconstructed weights)
        net.set_quant_levels(levels, [step] * 12)
        res.append(out)
    return res


@torch.no_grad()
def quantize_net(net: NeRF, qp: int, qp_density: int = 2, nonweight_qp: int = -75) -> Dict[str, torch.Tensor]:
    return _quantize_nets([net], qp, qp_density, nonweight_qp)[0]


@torch.no_grad()
def quantize_model(wrapper, qp: int, qp_density: int = 2, nonweight_qp: int = -75, master_state: Optional[dict] = None):
    """Quantise + reconstruct both networks of a NeRFWrapper in place; returns {net: {tensor: int32 levels}}.
    master_state: optional state_dict-shaped {name: float32 tensor} of UNQUANTISED values to quantise from (the wrapper's
    parameters then receive level*delta without being read) -- what a per-step requantisation uses."""
    sources = None
    if master_state is not None:
        sd = wrapper.state_dict()
        sources = {sd[k].data_ptr(): master_state[k].contiguous() for k in sd if k.endswith(".weight") or k.endswith(".bias")}
    a, b = _quantize_nets([wrapper.model, wrapper.model_fine], qp, qp_density, nonweight_qp, sources)
    return {"model": a, "model_fine": b}


@torch.no_grad()
def quantize_scales(wrapper, qp_density: int = 2, nonweight_qp: int = -75):
    """Final pass of nnc/compression.py:534-538 for the tuned scales: ls -> dequant(quant(ls, -75))."""
    for net in (wrapper.model, wrapper.model_fine):
        for s in net.scale_tensors():
            if s is not None:
                lv, _ = ops.quantize_urq(s.detach().float(), nonweight_qp, qp_density)
                s.copy_(ops.dequantize(lv, nonweight_qp, qp_density))


@torch.no_grad()
def apply_lsa(wrapper):
    """Decoder side of LSA (nnc_core/approximator/__init__.py:276-318): fold every `weight_scaling` into its weight,
    `w *= ls.reshape(-1, 1, ...)`, and drop the scales.  Returns a plain NeRFWrapper (nn.Linear layers, no
    `weight_scaling` entries in its state_dict) holding float32 weights `level * delta * ls` -- what
    `decompress_model` hands to the test-view renderer.  The LSA model itself can be rendered without this detour:
    the fused MLP applies `delta * ls` per output channel in its epilogue with the integer levels as operands."""
    from .model import NeRFWrapper
    src = wrapper.state_dict()
    out = NeRFWrapper().to(next(wrapper.parameters()).device)
    dst = out.state_dict()
    for name, value in src.items():
        if name.endswith("weight_scaling"):
            continue
        if name.endswith(".weight"):
            ls = src.get(name[:-len("weight")] + "weight_scaling")
            value = value * ls.reshape([-1] + [1] * (value.dim() - 1)) if ls is not None else value
        dst[name].copy_(value)
    for net in (out.model, out.model_fine):
        net.set_quant_levels(None, None)
    return out


@torch.no_grad()
def load_levels(wrapper, levels: Dict[str, "torch.Tensor"], qps: Dict[str, int], qp_density: int = 2):
    """Decoder side (nnc_core/approximator/__init__.py:276-318 `rec` + `apply_lsa`, nnc/compression.py:636-659) without the
    float detour: the decoded integer levels go straight into the packed networks.

    levels   {state_dict name: int32 tensor or numpy array}: '<layer>.weight', '<layer>.bias' and, for an LSA bitstream,
             '<layer>.weight_scaling' -- what Decoder.decodeLayer produced
    qps      {same names: qp actually used} (the iae_v field in front of each tensor, coder/baseline.py:31-34)

    Weight levels become the MLP kernels' integer operands (model.NeRF.set_quant_levels); biases and scales are
    dequantised (level * delta, one kernel each) into the module's parameters; the float `weight` parameters are set to
    level * delta so state_dict() equals what the reference's `rec` yields.  The wrapper must carry LSA parameters when the
    bitstream has scales; they are applied in the MLP epilogue, i.e. the reconstruction equals apply_lsa's
    `w * ls` up to the fp16 rounding apply_lsa's folded weights would get as operands."""
    import numpy as np
    dev = next(wrapper.parameters()).device
    sd = wrapper.state_dict()

    def as_dev(x):
        t = torch.from_numpy(np.ascontiguousarray(x)) if not torch.is_tensor(x) else x
        return t.to(device=dev, dtype=torch.int32).contiguous()

    for prefix, net in (("model", wrapper.model), ("model_fine", wrapper.model_fine)):
        lv_w, steps = [], []
        for name, layer in zip(packed.LAYER_NAMES, net.layers()):
            key = f"{prefix}.{name}"
            lw = as_dev(levels[key + ".weight"]).reshape(layer.weight.shape)
            qw = int(qps[key + ".weight"])
            lv_w.append(lw)
            steps.append(ops.stepsize(qw, qp_density))
            layer.weight.copy_(ops.dequantize(lw, qw, qp_density))
            layer.bias.copy_(ops.dequantize(as_dev(levels[key + ".bias"]), int(qps[key + ".bias"]), qp_density).reshape(layer.bias.shape))
            ls_key = key + ".weight_scaling"
            if ls_key in levels:
                if ls_key not in sd:
                    raise ValueError(f"the bitstream carries {ls_key} but the wrapper has no LSA parameters (model.LSA(w).add_lsa_params())")
                layer.weight_scaling.copy_(ops.dequantize(as_dev(levels[ls_key]), int(qps[ls_key]), qp_density).reshape(layer.weight_scaling.shape))
        net.set_quant_levels(lv_w, steps)
    return wrapper
