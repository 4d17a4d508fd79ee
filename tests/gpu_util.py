"""Helpers shared by the GPU tests: build our modules from the golden fixture model."""
import numpy as np
import torch

import nerfq_b200  # noqa: F401
from nerfq_b200 import model as nmodel
from tests.util import LAYERS, NETS, golden_model_levels, golden_model_params, golden_trained_params


def golden_wrapper(dev, with_levels: bool):
    """Our NeRFWrapper+LSA carrying the fixture model; with_levels packs int32 levels + delta
    (the quantised path), otherwise float weights (level*delta) are packed."""
    p, delta = golden_model_params()
    levels, _ = golden_model_levels()
    w = nmodel.LSA(nmodel.NeRFWrapper()).add_lsa_params()
    sd = {k: v.reshape(-1, 1) if k.endswith("weight_scaling") else v for k, v in p.items()}
    w.load_state_dict(sd)
    w = w.to(dev)
    if with_levels:
        for net in NETS:
            m = getattr(w, net)
            m.set_quant_levels([torch.from_numpy(levels[f"{net}.{l}"]).to(dev) for l in LAYERS], [delta] * 12)
    return w, p


def trained_wrapper(dev, variant="spread"):
    """Our NeRFWrapper+LSA carrying the trained-like fixture (qp=-38, levels beyond 2048) with its integer levels attached."""
    p, levels, delta = golden_trained_params(variant)
    w = nmodel.LSA(nmodel.NeRFWrapper()).add_lsa_params()
    w.load_state_dict({k: v.reshape(-1, 1) if k.endswith("weight_scaling") else v for k, v in p.items()})
    w = w.to(dev)
    for net in NETS:
        getattr(w, net).set_quant_levels([torch.from_numpy(levels[f"{net}.{l}"]).to(dev) for l in LAYERS], [delta] * 12)
    return w, p
