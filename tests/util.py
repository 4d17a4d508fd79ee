"""Shared helpers for the tests: golden-fixture loading and synthetic workloads."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

LAYERS = tuple([f"pts_linears.{i}" for i in range(8)] +
               ["alpha_linear", "feature_linear", "views_linears.0", "rgb_linear"])
NETS = ("model", "model_fine")


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def golden_model_params():
    """The fixture model of tests/golden/make_golden.py::build_model as a NeRFWrapper-style
    state dict (weights = int8 level * delta, exact in fp32)."""
    z = golden("model_qm20.npz")
    delta = float(z["delta"])
    p = {}
    for net in NETS:
        for l in LAYERS:
            k = f"{net}.{l}"
            p[k + ".weight"] = torch.from_numpy(z[k + ".levels"].astype(np.float32) * np.float32(delta))
            p[k + ".bias"] = torch.from_numpy(z[k + ".bias"])
            p[k + ".weight_scaling"] = torch.from_numpy(z[k + ".weight_scaling"])
    return p, delta


def golden_model_levels():
    z = golden("model_qm20.npz")
    return {f"{net}.{l}": z[f"{net}.{l}.levels"].astype(np.int32) for net in NETS for l in LAYERS}, float(z["delta"])


def synth_rays(n, seed, near=2.0, far=6.0):
    """Same recipe as SURVEY 8(d) cfg1: origins near (0,0,4), unit-ish directions toward -z."""
    g = torch.Generator().manual_seed(seed)
    o = 0.1 * torch.randn(n, 3, generator=g) + torch.tensor([0.0, 0.0, 4.0])
    d = torch.randn(n, 3, generator=g)
    d = -d / torch.norm(d, dim=-1, keepdim=True)
    vd = d / torch.norm(d, dim=-1, keepdim=True)
    return torch.cat([o, d, near * torch.ones(n, 1), far * torch.ones(n, 1), vd], -1)


def golden_trained_params(variant="spread"):
    """The trained-like fixture model of make_golden.py::build_model_trained at qp=-38: returns (state-dict style float
    params, {layer: int32 levels}, delta).  variant 'dense' swaps in the alpha head with densities around 150."""
    z = golden("model_trained_qm38.npz")
    delta = np.float32(z["delta"])
    p, levels = {}, {}
    for net in NETS:
        for l in LAYERS:
            k = f"{net}.{l}"
            lv = z[k + ".levels"].astype(np.int32)
            bias = z[k + ".bias"].copy()
            if variant == "dense" and l == "alpha_linear":
                lv = z[k + ".levels_dense"].astype(np.int32)
                bias = bias + np.float32(z["dense_alpha_bias_shift"])
            levels[k] = lv
            p[k + ".weight"] = torch.from_numpy(lv.astype(np.float32) * delta)
            p[k + ".bias"] = torch.from_numpy(bias)
            p[k + ".weight_scaling"] = torch.from_numpy(z[k + ".weight_scaling"])
    return p, levels, float(delta)
