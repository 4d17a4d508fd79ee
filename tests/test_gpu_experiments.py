"""The two opt-in CTA-pair forward kernels (csrc/mlp4_fwd.cu, csrc/mlp5_fwd.cu) stay parity-checked:
each is run in a subprocess with NERFQ_MLP_FWD set (the choice is read once per process) on a ragged batch and compared
with the default single-CTA kernel."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_SCRIPT = r"""
import sys, numpy as np, torch
sys.path.insert(0, %r)
import nerfq_b200
from nerfq_b200 import codec, model as nmodel, packed
from tests.util import synth_rays
dev = torch.device("cuda:0")
torch.manual_seed(7)
w = nmodel.LSA(nmodel.NeRFWrapper()).add_lsa_params().to(dev)
codec.quantize_model(w, -20)
pn = w.model_fine.packed_net()
pn.set_scales(w.model_fine.scale_tensors())
n, S = 1000, 37                                    # 37000 points = 144.5 groups: ragged tail, odd number of groups
r = synth_rays(n, 41).to(dev)
z = torch.sort(2.0 + 4.0 * torch.rand(n, S, device=dev, generator=torch.Generator(device=dev).manual_seed(1)), -1).values.contiguous()
raw = packed.mlp_forward(pn, r, z)
raw2 = packed.mlp_forward(pn, r, z)
torch.cuda.synchronize()
assert torch.isfinite(raw).all()
np.save(sys.argv[1], raw.cpu().numpy())
print("repeatable", bool(torch.equal(raw, raw2)))
"""


def _run(tmp_path, variant):
    out = str(tmp_path / f"raw_{variant}.npy")
    env = dict(os.environ)
    env.pop("NERFQ_MLP_FWD", None)
    if variant:
        env["NERFQ_MLP_FWD"] = str(variant)
    res = subprocess.run([sys.executable, "-c", _SCRIPT % ROOT, out], env=env, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    return np.load(out), "repeatable True" in res.stdout


@pytest.mark.parametrize("variant", [4, 5])
def test_pair_schedule_kernels_match_the_default(tmp_path, variant):
    ref, ref_rep = _run(tmp_path, 0)
    got, rep = _run(tmp_path, variant)
    assert ref_rep                                              # the default kernel is bit-reproducible
    assert got.shape == ref.shape
    # rgb logits: same arithmetic per element; sigma: the alpha head is reduced in a different order / rounding
    assert np.abs(got[..., :3] - ref[..., :3]).max() <= 1e-5 * max(1.0, np.abs(ref[..., :3]).max())
    assert np.abs(got[..., 3] - ref[..., 3]).max() <= 1e-4 * max(1.0, np.abs(ref[..., 3]).max())
    if variant == 5:
        assert rep                                              # fixed-point alpha head there too
