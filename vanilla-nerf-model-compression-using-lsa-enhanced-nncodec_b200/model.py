"""Model types of the reference, kept as plain nn.Modules with identical state_dict keys:

  NeRF          utils.py:18-80          (D=8, W=256, skip after layer 4, view-direction head)
  NeRFWrapper   utils.py:84-106         (coarse `model` + fine `model_fine`, tuning_optimizer, global_step)
  ScaledLinear  transforms.py:84-111    (nn.Linear + weight_scaling [out,1])
  LSA           transforms.py:113-168   (replace every nn.Linear by ScaledLinear)

`NeRF.forward` is never used on the rendering path of this package: the renderer packs the module's
parameters into the fused-kernel format (packed.PackedNet) and runs the tcgen05 kernels.  It is kept,
with the reference's semantics, so the modules remain usable as ordinary torch modules.
"""
import copy
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import packed

_LAYER_ATTRS = tuple([("pts_linears", i) for i in range(8)] +
                     [("alpha_linear", None), ("feature_linear", None), ("views_linears", 0), ("rgb_linear", None)])


class ScaledLinear(nn.Linear):
    """transforms.py:84-111: y = (weight_scaling * weight) x + bias, weight_scaling ~ N(1, 1e-5)."""

    def __init__(self, in_features, out_features, *args, **kwargs):
        super().__init__(in_features, out_features, *args, **kwargs)
        self.weight_scaling = nn.Parameter(torch.ones(out_features, 1))
        self.reset_parameters()

    def reset_parameters(self):
        if hasattr(self, "weight_scaling"):
            nn.init.normal_(self.weight_scaling, 1, 1e-5)
            super().reset_parameters()

    def forward(self, input):
        return F.linear(input, self.weight_scaling * self.weight, self.bias)


class NeRF(nn.Module):
    """utils.py:18-80.  Only the architecture the fused kernels implement is accepted."""

    def __init__(self, D=8, W=256, input_ch=3, input_ch_views=3, output_ch=4, skips=[4], use_viewdirs=False):
        super().__init__()
        self.D, self.W = D, W
        self.input_ch, self.input_ch_views = input_ch, input_ch_views
        self.skips, self.use_viewdirs = skips, use_viewdirs
        self.pts_linears = nn.ModuleList(
            [nn.Linear(input_ch, W)] +
            [nn.Linear(W, W) if i not in self.skips else nn.Linear(W + input_ch, W) for i in range(D - 1)])
        self.views_linears = nn.ModuleList([nn.Linear(input_ch_views + W, W // 2)])
        if use_viewdirs:
            self.feature_linear = nn.Linear(W, W)
            self.alpha_linear = nn.Linear(W, 1)
            self.rgb_linear = nn.Linear(W // 2, 3)
        else:
            self.output_linear = nn.Linear(W, output_ch)
        self._packed: Optional[packed.PackedNet] = None
        self._packed_key = None
        self.quant_levels = None       # optional: 12 int32 level tensors + 12 step sizes, see set_quant_levels
        self.quant_steps = None
        self._quant_key = None         # identity + version of the float weights the levels belong to

    def forward(self, x):
        input_pts, input_views = torch.split(x, [self.input_ch, self.input_ch_views], dim=-1)
        h = input_pts
        for i, layer in enumerate(self.pts_linears):
            h = F.relu(layer(h))
            if i in self.skips:
                h = torch.cat([input_pts, h], -1)
        if not self.use_viewdirs:
            return self.output_linear(h)
        alpha = self.alpha_linear(h)
        h = torch.cat([self.feature_linear(h), input_views], -1)
        for layer in self.views_linears:
            h = F.relu(layer(h))
        return torch.cat([self.rgb_linear(h), alpha], -1)

    # ---- bridge to the fused kernels ----------------------------------------------------------
    def layers(self):
        out = []
        for attr, idx in _LAYER_ATTRS:
            m = getattr(self, attr)
            out.append(m if idx is None else m[idx])
        return out

    def supported(self) -> bool:
        return (self.D == 8 and self.W == 256 and self.input_ch == 63 and self.input_ch_views == 27 and
                list(self.skips) == [4] and self.use_viewdirs)

    def scale_tensors(self):
        """The 12 LSA scale parameters in layer order (None entries when the model has no LSA)."""
        return [getattr(l, "weight_scaling", None) for l in self.layers()]

    def _weight_key(self):
        return tuple((l.weight.data_ptr(), l.weight._version) for l in self.layers())

    def set_quant_levels(self, levels, steps):
        """Attach the integer levels (12 int32 tensors [out,in]) and step sizes whose product the float `weight`
        parameters hold right now.  They are the MLP kernels' operands for as long as the float weights stay untouched:
        any later load_state_dict / in-place update of a weight (a version-counter bump) drops them, and the renderer
        packs the float weights again -- so state_dict(), NeRF.forward and the fused path never disagree."""
        # the same level tensors as before (a requantisation wrote new values into them, codec._quantize_nets): repack into the
        # existing buffers instead of allocating new ones, so every holder of the packed network -- python attributes, and
        # each CUDA graph that recorded a requantisation -- keeps seeing the one live copy
        same = (levels is not None and self.quant_levels is not None and self._packed is not None and self._packed.is_int and
                len(levels) == len(self.quant_levels) and all(a.data_ptr() == b.data_ptr() for a, b in zip(levels, self.quant_levels)))
        self.quant_levels, self.quant_steps = (list(levels), list(steps)) if levels is not None else (None, None)
        self._quant_key = self._weight_key() if levels is not None else None
        if same:
            self._packed.repack(self.quant_levels, self.quant_steps, [l.bias.detach() for l in self.layers()])
            self._packed_key = self._pack_key()
        else:
            self._packed = None

    def _pack_key(self):
        return tuple((l.weight.data_ptr(), l.weight._version, l.bias.data_ptr(), l.bias._version) for l in self.layers()) + \
            (id(self.quant_levels),)

    def packed_net(self) -> packed.PackedNet:
        """Pack (or reuse) the frozen weights; biases and scales are refreshed by the caller."""
        if not self.supported():
            raise NotImplementedError("the fused kernels implement the vanilla NeRF architecture only "
                                      "(D=8, W=256, input_ch=63, input_ch_views=27, skips=[4], use_viewdirs=True)")
        ls = self.layers()
        if self.quant_levels is not None and self._quant_key is not None and self._quant_key != self._weight_key():
            self.quant_levels = self.quant_steps = self._quant_key = None       # the float weights moved on: levels are stale
        key = self._pack_key()
        if self._packed is None or key != self._packed_key:
            biases = [l.bias.detach() for l in ls]
            if self.quant_levels is not None:
                self._packed = packed.PackedNet(self.quant_levels, self.quant_steps, biases)
            else:
                self._packed = packed.PackedNet([l.weight.detach().float().contiguous() for l in ls], [1.0] * 12, biases)
            self._packed_key = key
        return self._packed

    def __deepcopy__(self, memo):
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k in ("_packed", "_packed_key"):
                new.__dict__[k] = None
            else:
                new.__dict__[k] = copy.deepcopy(v, memo)
        if new.__dict__.get("quant_levels") is not None:
            new._quant_key = new._weight_key()          # the copy's weights are new tensors holding the same values
        return new


class NeRFWrapper(nn.Module):
    """utils.py:84-106."""

    def __init__(self, D=8, W=256, input_ch=63, input_ch_views=27, output_ch=4, skips=[4], use_viewdirs=True):
        super().__init__()
        self.model = NeRF(D=D, W=W, input_ch=input_ch, input_ch_views=input_ch_views, output_ch=output_ch,
                          skips=skips, use_viewdirs=use_viewdirs)
        self.model_fine = NeRF(D=D, W=W, input_ch=input_ch, input_ch_views=input_ch_views, output_ch=output_ch,
                               skips=skips, use_viewdirs=use_viewdirs)
        self.tuning_optimizer = None
        self.global_step = 0


class LSA:
    """transforms.py:113-168: deep-copy the model and swap each nn.Linear for a ScaledLinear that
    shares its weight and bias."""

    def __init__(self, original_model):
        self.mdl = copy.deepcopy(original_model)

    @staticmethod
    def _swap(parent: nn.Module):
        for name, child in list(parent.named_children()):
            if isinstance(child, nn.Linear) and not isinstance(child, ScaledLinear) and child.weight.requires_grad:
                new = ScaledLinear(child.in_features, child.out_features)
                new.weight, new.bias = child.weight, child.bias
                setattr(parent, name, new)
            else:
                LSA._swap(child)

    def add_lsa_params(self):
        self._swap(self.mdl)
        return self.mdl
