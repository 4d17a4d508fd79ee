"""CPU tests of the host side: the C-ABI library exports what include/nerfq.h declares, the model types keep
the reference's state_dict layout, the product never touches oracle/, and the multi-GPU plumbing (gloo, 2 ranks)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "vanilla-nerf-model-compression-using-lsa-enhanced-nncodec_b200")


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    hdr = open(os.path.join(ROOT, "include", "nerfq.h")).read()
    names = sorted(set(re.findall(r"\b(nerfq_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 17
    lib = ctypes.CDLL(os.path.join(PKG, "libnerfq.so"))
    for n in names:
        assert hasattr(lib, n), n
    # host-only entry points can be called without a GPU
    lib.nerfq_packed_net_bytes.restype = ctypes.c_ulonglong
    assert lib.nerfq_packed_net_bytes() > 2_000_000
    assert lib.nerfq_num_channels() == 2436
    out = ctypes.c_float()
    assert lib.nerfq_stepsize(-20, 2, ctypes.byref(out)) == 0 and out.value == 0.03125
    assert lib.nerfq_stepsize(-20, 2, None) == -1


def test_sass_uses_tcgen05_and_bulk_copies():
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(PKG, "libnerfq.so")], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass or "UTCMMA" in sass, "tcgen05.mma missing from SASS"
    assert "LDTM" in sass, "tcgen05.ld missing from SASS"
    assert "UBLKCP" in sass, "bulk async copy missing from SASS"


def test_missing_library_fails_loudly(monkeypatch):
    import nerfq_b200  # noqa: F401
    from nerfq_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libnerfq.so")
    with pytest.raises(RuntimeError, match="no fallback"):
        _lib.lib()


def test_product_never_imports_oracle():
    for fn in os.listdir(PKG):
        if fn.endswith(".py"):
            src = open(os.path.join(PKG, fn)).read()
            assert "oracle" not in src.replace("# noqa", ""), fn
    assert "oracle" not in open(os.path.join(ROOT, "nerfq_b200.py")).read()


def test_model_types_keep_reference_layout():
    import copy
    import nerfq_b200  # noqa: F401
    from nerfq_b200 import model, packed
    torch.manual_seed(0)
    w = model.NeRFWrapper()
    assert len(w.state_dict()) == 48 and w.tuning_optimizer is None and w.global_step == 0
    m = model.LSA(w).add_lsa_params()
    sd = m.state_dict()
    assert len(sd) == 72
    for net in ("model", "model_fine"):
        for l, (o, i) in zip(packed.LAYER_NAMES, zip(packed.LAYER_OUT, packed.LAYER_IN)):
            assert tuple(sd[f"{net}.{l}.weight"].shape) == (o, i)
            assert tuple(sd[f"{net}.{l}.weight_scaling"].shape) == (o, 1)
            assert tuple(sd[f"{net}.{l}.bias"].shape) == (o,)
    assert len(w.state_dict()) == 48                      # LSA works on a deep copy (transforms.py:117)
    x = torch.randn(5, 90)
    ref = m.model(x)
    assert ref.shape == (5, 4)
    c = copy.deepcopy(m)
    assert torch.equal(c.model(x), ref) and c.model._packed is None
    flat = packed.flatten_channels([sd[f"model.{l}.bias"] for l in packed.LAYER_NAMES])
    assert flat.numel() == 2436
    back = packed.split_channels(flat)
    for l, name in enumerate(packed.LAYER_NAMES):
        assert torch.equal(back[l], sd[f"model.{name}.bias"])
    with pytest.raises(NotImplementedError):
        model.NeRF(D=4, W=128, input_ch=63, input_ch_views=27, use_viewdirs=True).packed_net()


def test_shard_range_covers_everything():
    import nerfq_b200  # noqa: F401
    from nerfq_b200.distributed import shard_range
    for n in (0, 1, 7, 640000, 190512):
        for world in (1, 2, 3, 8):
            pos = 0
            for r in range(world):
                f, c = shard_range(n, r, world)
                assert f == pos and c >= 0
                pos += c
            assert pos == n


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
import nerfq_b200
from nerfq_b200.distributed import allreduce_fixed, shard_range
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % os.environ["PORT"], rank=rank, world_size=world)
# the data-parallel gradient exchange: every rank holds 64-bit fixed-point partial sums (value * 2^48) of ITS rays; the
# integer all-reduce must equal the sum a single rank would have formed over all rays, bit for bit, for any split
gen = torch.Generator().manual_seed(7)
terms = (torch.randn(64, 2440, generator=gen, dtype=torch.float64) * 1e-4 * 2.0 ** 48).round().to(torch.int64)     # 64 "rays"
whole = terms.sum(0)
lo, cnt = shard_range(64, rank, world)
mine = terms[lo:lo + cnt].sum(0)
got = allreduce_fixed(mine.clone())
assert torch.equal(got, whole)
assert not torch.equal(mine, whole)
# sharded "render": each rank fills its pixel slice, all_gather with padding reassembles the image
n = 1001
f, c = shard_range(n, rank, world)
local = torch.arange(f, f + c, dtype=torch.float32)
maxc = shard_range(n, 0, world)[1]
pad = torch.zeros(maxc); pad[:c] = local
parts = [torch.empty_like(pad) for _ in range(world)]
dist.all_gather(parts, pad)
full = torch.cat([parts[r][:shard_range(n, r, world)[1]] for r in range(world)])
assert torch.equal(full, torch.arange(n, dtype=torch.float32))
dist.destroy_process_group()
print("ok", rank)
"""


def test_two_rank_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", PORT=port)
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=240)
        assert p.returncode == 0, out


def test_lsa_parameter_selection_matches_tune_model():
    """lsa.lsa_parameters mirrors tune_model(lsa_flag=True, ft_flag=False): only the 24 weight_scaling tensors train
    (framework/pytorch_model/__init__.py:1131-1145); LSAStep refuses a CPU model instead of falling back."""
    import pytest
    from nerfq_b200 import lsa, model as nmodel
    w = nmodel.LSA(nmodel.NeRFWrapper()).add_lsa_params()
    params = lsa.lsa_parameters(w)
    assert len(params) == 24
    trainable = [n for n, p in w.named_parameters() if p.requires_grad]
    assert len(trainable) == 24 and all(n.endswith("weight_scaling") for n in trainable)
    with pytest.raises(RuntimeError):
        lsa.LSAStep(w, 16)


def test_deepcabac_front_end_contract():
    """nerfq_b200.deepcabac keeps the names baseline.py:24-57,89-98 and coder/baseline.py:5-57 bind.  Elementwise
    (de)quantisation belongs to the GPU kernels: without a CUDA device it refuses loudly (no silent CPU fallback) unless the
    caller explicitly selects the host library; argument errors are raised before any device work; empty tensors are
    accepted."""
    import numpy as np
    import pytest
    from nerfq_b200 import deepcabac
    assert deepcabac.DEVICE == "cuda"
    enc, dec = deepcabac.Encoder(), deepcabac.Decoder()
    for name in ("initCtxModels", "quantLayer", "iae_v", "encodeLayer", "finish"):
        assert callable(getattr(enc, name))
    for name in ("setStream", "initCtxModels", "iae_v", "decodeLayer", "decodeLayerAndCreateEPs", "setEntryPoints", "dequantLayer", "finish"):
        assert callable(getattr(dec, name))
    enc.initCtxModels(10, 0)
    w = np.ones((4, 3), dtype=np.float32)
    out = np.zeros((4, 3), dtype=np.int32)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            enc.quantLayer(w, out, 0, 2, -20, 0.0, 10, 0)
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            dec.dequantLayer(w, out, 2, -20, 0)
    assert enc.quantLayer(w, out, 1, 2, -20, 0.0, 10, 0) == -20        # dependent quantisation: host trellis search by design
    assert np.abs(out * 0.03125 - w).max() <= 2 * 0.03125
    with pytest.raises(TypeError):
        enc.quantLayer(w.astype(np.float64), out, 0, 2, -20, 0.0, 10, 0)
    with pytest.raises(ValueError):
        enc.quantLayer(np.ones((3, 4), dtype=np.float32).T, out, 0, 2, -20, 0.0, 10, 0)      # transposed view, baseline.py:34-36
    with pytest.raises(ValueError):
        dec.dequantLayer(np.zeros(5, dtype=np.float32), np.zeros(4, dtype=np.int32), 2, -20, 0)
    e_w, e_l = np.zeros((0, 7), dtype=np.float32), np.zeros((0, 7), dtype=np.int32)
    assert enc.quantLayer(e_w, e_l, 0, 2, -20, 0.0, 10, 0) == -20
    dec.dequantLayer(e_w, e_l, 2, -20, 0)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the arm the driver divides by): runs on the host cores alone, prints ONE JSON line with the
    same metric / unit / config.workload as the CUDA arm, `impl: reference`, a cpu_baseline describing the run and an e2e
    object that repeats the line's value; a non-zero rank under torchrun exits 0 without output."""
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    sys.path.insert(0, root)
    import bench
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == "rays/s" and d["higher_is_better"] is True
    assert d["config"]["workload"] == bench.WORKLOAD and d["config"]["rays_per_step_timed"] in (4096, 1024)
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert abs(d["ms_per_step"] - 1e3 * 4096 / d["value"]) < 1e-6 * d["ms_per_step"] and d["gpu_launches"] == 0
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1"], capture_output=True, text=True,
                         env=dict(env, RANK="1", WORLD_SIZE="2"), timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""
