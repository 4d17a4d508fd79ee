"""The host coder (include/nncabac.h, nerfq_b200.deepcabac) on the CPU:
  * arithmetic-coder round trips with exact byte accounting (what nnc_core/coder/__init__.py:154,162,484 assert),
  * dependent quantisation: trellis output is a valid path, reconstructs within its bound, beats the coarse grid,
  * the UNMODIFIED reference (imported from /root/reference with `sys.modules['deepCABAC']` = ours) runs
    nnc.compress_model(lsa=True) -> nnc.decompress_model on a NeRFWrapper and reproduces level * delta * ls.
Elementwise (de)quantisation normally runs on the GPU kernels; here deepcabac.DEVICE is set to "host" EXPLICITLY (this
container has no GPU).  tests/test_gpu_codec.py compares the two devices bit for bit."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAVE_REF = os.path.isdir("/root/reference/nnc_core")


@pytest.fixture()
def dc():
    import nerfq_b200  # noqa: F401
    from nerfq_b200 import deepcabac
    old = deepcabac.DEVICE
    deepcabac.DEVICE = "host"
    yield deepcabac
    deepcabac.DEVICE = old


def _roundtrip(dc, layers, unary=10, opt=1, tail=b""):
    enc = dc.Encoder()
    for lv, dq, qp in layers:
        enc.iae_v(8, qp)
        enc.initCtxModels(unary, opt)
        enc.encodeLayer(lv, dq, 0)
    bs = enc.finish().tobytes()
    dec = dc.Decoder()
    dec.setStream(bytearray(bs + tail))
    for lv, dq, qp in layers:
        assert dec.iae_v(8) == qp
        dec.initCtxModels(unary)
        out = np.zeros(lv.shape, dtype=np.int32)
        dec.decodeLayer(out, dq, 0)
        assert (out == lv).all()
    assert dec.finish() == len(bs)                    # the decoder knows where the codeword ends, whatever follows it
    return len(bs)


def test_header_symbols_exported():
    import ctypes
    import re
    import __graft_entry__ as ge
    ge.build()
    hdr = open(os.path.join(ROOT, "include", "nncabac.h")).read()
    names = sorted(set(re.findall(r"\b(nncabac_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 16
    from nerfq_b200 import deepcabac
    lib = ctypes.CDLL(deepcabac.HOST_LIB_PATH)
    for n in names:
        assert hasattr(lib, n), n


def test_level_streams_round_trip(dc):
    rng = np.random.default_rng(0)
    for trial in range(120):
        n = int(rng.integers(0, 400))
        lv = np.round(rng.standard_normal(n) * 10.0 ** rng.uniform(-1, 6)).astype(np.int32)
        if n and trial % 7 == 0:
            lv[rng.integers(0, n)] = np.int32(2 ** 31 - 1) if trial % 2 else np.int32(-2 ** 31 + 1)
        tail = bytes(rng.integers(0, 256, 5).astype(np.uint8)) if trial % 3 else b""
        _roundtrip(dc, [(lv, 0, -38)], unary=int(rng.integers(0, 31)), opt=trial & 1, tail=tail)
    # several tensors in one codeword, as a block NDU carries them (weight_scaling, bias, weight): coder/baseline.py:5-20
    w = np.round(rng.standard_normal((64, 31)) * 20).astype(np.int32)
    _roundtrip(dc, [(np.round(rng.standard_normal(64) * 4e5).astype(np.int32), 0, -75), (np.zeros(64, np.int32), 0, -75), (w, 0, -20)])
    # compression actually happens: N(0, 0.05) weights at qp -20 need ~2.7 bits each
    lv, _ = dc.host_quant_layer((rng.standard_normal((256, 256)) * 0.05).astype(np.float32), 0, 2, -20)
    assert _roundtrip(dc, [(lv, 0, -20)]) < 0.1 * lv.size * 4


def test_corrupt_stream_is_reported(dc):
    dec = dc.Decoder()
    dec.setStream(bytearray(b"\x00\x01"))
    dec.initCtxModels(10)
    out = np.zeros(4096, dtype=np.int32)
    with pytest.raises(dc.CoderError):
        dec.decodeLayer(out, 0, 0)
        dec.finish()
    with pytest.raises(NotImplementedError):
        dc.Encoder().encodeLayer(np.zeros((8, 8), np.int32), 0, 1)           # block scans are not implemented: loud


def test_dependent_quantisation(dc):
    rng = np.random.default_rng(3)
    for trial in range(40):
        n = int(rng.integers(1, 3000))
        w = (rng.standard_normal(n) * 10.0 ** rng.uniform(-3, 0)).astype(np.float32)
        qp = int(rng.integers(-40, -8))
        out = np.zeros(n, dtype=np.int32)
        enc = dc.Encoder()
        enc.initCtxModels(10, 0)
        used = enc.quantLayer(w, out, 1, 2, qp, 0.0, 10, 0)
        assert used == qp
        d = np.float32(dc.host_lib().nncabac_stepsize(used, 2))
        err = np.abs(out.astype(np.float32) * d - w)
        assert err.max() <= 2.0 * d * 1.0001                      # sub-quantiser spacing 2 delta
        coarse = np.round(w / (2 * d)) * 2 * d                    # a scalar quantiser with the sub-quantisers' spacing
        if n > 200:
            assert (err ** 2).mean() <= ((coarse - w) ** 2).mean()      # the trellis never loses to it ...
            if w.std() > 4 * d:
                assert (err ** 2).mean() < 0.9 * ((coarse - w) ** 2).mean()      # ... and gains where there is something to code
        urq, _ = dc.host_quant_layer(w, 0, 2, used)
        _roundtrip(dc, [(out, 1, used), (urq, 0, used)], tail=b"\x01\x02\x03")      # a valid trellis path, decodable
    bad = np.array([1, 0, 0], dtype=np.int32)                     # odd value in state 0 (quantiser Q0): not a path
    with pytest.raises(dc.CoderError):
        e = dc.Encoder(); e.initCtxModels(10, 0); e.encodeLayer(bad, 1, 0)
    nf = np.array([0.5, np.nan, np.inf, -0.25], dtype=np.float32)
    out = np.zeros(4, dtype=np.int32)
    assert dc.Encoder().quantLayer(nf, out, 1, 2, -20, 0.0, 10, 0) == -20 and out[1] == 0 and out[2] == 0


def test_host_urq_matches_the_c_restatement(dc):
    from oracle import quant_oracle as qo
    rng = np.random.default_rng(5)
    for qp in range(-38, -9):
        w = (rng.standard_normal(5000) * 0.2).astype(np.float32)
        w[:3] = (0.0, -0.0, np.float32(qo.stepsize(qp, 2)) * 1.5)
        lv, used = dc.host_quant_layer(w, 0, 2, qp)
        ref, used_ref = qo.quant_urq(w, qp, 2)
        assert used == used_ref and (lv == ref).all()
        assert (dc.host_dequant_layer(lv, 2, qp) == qo.dequant(ref, qp, 2)).all()
    big = np.array([3e4, -1.0, 0.3], dtype=np.float32)
    lv, used = dc.host_quant_layer(big, 0, 2, -75)
    ref, used_ref = qo.quant_urq(big, -75, 2)
    assert used == used_ref > -75 and (lv == ref).all()


@pytest.mark.skipif(not HAVE_REF, reason="/root/reference is only present in the build container")
@pytest.mark.parametrize("use_dq", [True, False])
def test_unmodified_reference_compress_decompress(dc, tmp_path, use_dq):
    """nnc.compress_model(lsa=True) -> nnc.decompress_model of the reference itself, with our module as `deepCABAC`.
    The reference's training loop (dataset loaders, run_nerf.train) is replaced by a stub that nudges the LSA scales:
    everything between the model and the .nnc file -- block detection, approx / rec around tuning, set_lsa, the final
    approx, the NNR unit syntax with its size checks, decoding, rec and apply_lsa -- is the reference's own code."""
    sys.modules["deepCABAC"] = dc
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from _ref_import import import_reference
    ref = import_reference()
    import nnc
    import nnc_core
    import framework.pytorch_model as ref_pm
    assert nnc_core.coder.deepCABAC is dc and nnc_core.approximator.baseline.deepCABAC is dc

    torch.manual_seed(0)
    wrapper = ref.utils.NeRFWrapper()
    plain_sd = {k: v.clone() for k, v in wrapper.state_dict().items()}

    class Handler:              # stands in for use_cases['NERF_PYT'] (dataset + run_nerf.train): moves the scales, nothing else
        @staticmethod
        def train(nerf_wrapper, **kw):
            g = torch.Generator().manual_seed(5)
            with torch.no_grad():
                for n, p in nerf_wrapper.named_parameters():
                    if n.endswith("weight_scaling"):
                        p.add_(0.02 * torch.randn(p.shape, generator=g))
            return 0.0, 0.0
    old = ref_pm.use_cases["NERF_PYT"]
    ref_pm.use_cases["NERF_PYT"] = Handler
    qp = -20
    path = str(tmp_path / "model.nnc")
    try:
        bs = nnc.compress_model(wrapper, bitstream_path=path, qp=qp, use_dq=use_dq, lsa=True, model_struct=wrapper, dataset_path=str(tmp_path),
                                task_type="NeRF", dataset_type="blender", epochs=1, N_iters=1, learning_rate_decay=0, verbose=False,
                                return_bitstream=True)
    finally:
        ref_pm.use_cases["NERF_PYT"] = old
    assert os.path.getsize(path) == len(bs) and len(bs) < 0.25 * 4 * sum(v.numel() for v in plain_sd.values())
    out_pt = str(tmp_path / "rec.pt")
    nnc.decompress_model(path, model_path=out_pt, verbose=False)
    rec = torch.load(out_pt)
    assert set(rec.keys()) == set(plain_sd.keys())                    # 48 tensors: apply_lsa folded and dropped the scales
    # expected reconstruction: level * delta (* decoded scale) with the levels OUR quantiser yields for the original tensors
    d_w, d_o = np.float32(dc.host_lib().nncabac_stepsize(qp, 2)), np.float32(dc.host_lib().nncabac_stepsize(-75, 2))
    worst = 0.0
    for k, v in plain_sd.items():
        got = rec[k].numpy() if torch.is_tensor(rec[k]) else np.asarray(rec[k])
        if k.endswith(".bias"):
            lv, _ = dc.host_quant_layer(v.numpy(), int(use_dq), 2, -75)
            assert (got == lv.astype(np.float32) * d_o).all(), k
        else:
            lv, _ = dc.host_quant_layer(np.ascontiguousarray(v.numpy()), int(use_dq), 2, qp)
            base = lv.astype(np.float32) * d_w
            rows = np.abs(base).max(1) > 0
            ratio = got[rows] / np.where(base[rows] == 0, 1, base[rows])
            nz = base[rows] != 0
            per_row = np.array([np.median(r[m]) for r, m in zip(ratio, nz)], dtype=np.float32)
            assert np.abs(per_row - 1).max() < 0.15 and np.abs(per_row - 1).max() > 1e-4        # a real, small scale was applied
            assert np.allclose(got[rows], base[rows] * per_row[:, None], rtol=2e-6, atol=0), k   # same levels, one scale per row
            worst = max(worst, float(np.abs(got - v.numpy()).max()))
    # reconstruction error: the quantiser's bound (delta/2 uniform, 2 delta on the trellis) plus the applied scale's effect
    assert worst < (2.0 if use_dq else 0.5) * d_w + 0.15 * float(max(v.abs().max() for v in plain_sd.values()))


def test_frozen_stream(dc):
    """tests/golden/coder_stream.npz (made by tests/golden/make_coder_golden.py with THIS repository's coder: self-pinned, see the
    script's header): quantising the stored floats reproduces the stored levels (uniform and trellis), encoding them reproduces the
    stored bytes, and decoding the stored bytes yields the stored levels with exact byte accounting.  A change of the written
    format fails here."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "coder_stream.npz"))
    keys, dqs, qps = [str(k) for k in g["layer_keys"]], g["layer_dq"].tolist(), g["layer_qp"].tolist()
    enc = dc.Encoder()
    for key, dq, qp in zip(keys, dqs, qps):
        w = g[key[len("levels_"):key.rindex("_dq")]]
        lv = np.zeros(w.shape, dtype=np.int32)
        enc.initCtxModels(10, 0)
        assert enc.quantLayer(w, lv, dq, 2, qp, 0.0, 10, 0) == qp
        assert (lv == g[key]).all(), key
        enc.iae_v(8, qp + 128)
        enc.encodeLayer(lv, dq, 0)
    stream = enc.finish().tobytes()
    assert stream == g["stream"].tobytes()
    dec = dc.Decoder()
    dec.setStream(bytearray(g["stream"].tobytes() + b"\xff\x00"))
    for key, dq, qp in zip(keys, dqs, qps):
        dec.initCtxModels(10)
        assert dec.iae_v(8) == qp + 128
        out = np.zeros(g[key].shape, dtype=np.int32)
        dec.decodeLayer(out, dq, 0)
        assert (out == g[key]).all(), key
    assert dec.finish() == g["stream"].size


def test_decoder_survives_garbage(dc):
    """Random, truncated and bit-flipped streams: the decoder either returns levels or raises CoderError -- it never reads
    outside the buffer, loops or crashes -- and whatever it returns re-encodes to a stream that decodes to the same levels."""
    rng = np.random.default_rng(8)

    def attempt(stream, n, dq, unary=10):
        dec = dc.Decoder()
        dec.setStream(bytearray(stream))
        out = np.zeros(n, dtype=np.int32)
        try:
            dec.initCtxModels(unary)
            dec.decodeLayer(out, dq, 0)
        except dc.CoderError:
            return None
        try:
            dec.finish()
        except dc.CoderError:
            pass                                               # levels came out, the termination pattern did not match
        return out

    decoded = 0
    for trial in range(400):
        size = int(rng.integers(0, 64)) if trial % 2 else 4096
        out = attempt(bytes(rng.integers(0, 256, size).astype(np.uint8)), int(rng.integers(1, 300)), trial & 1, int(rng.integers(0, 31)))
        if out is not None and not (trial & 1):
            decoded += 1
            _roundtrip(dc, [(out, 0, -20)])
    assert decoded > 0
    lv = np.round(rng.standard_normal(3000) * 50).astype(np.int32)
    enc = dc.Encoder()
    enc.initCtxModels(10, 0)
    enc.encodeLayer(lv, 0, 0)
    good = enc.finish().tobytes()
    for cut in range(0, len(good), max(1, len(good) // 60)):
        attempt(good[:cut], 3000, 0)
    for _ in range(100):
        b = bytearray(good)
        b[int(rng.integers(0, len(b)))] ^= 1 << int(rng.integers(0, 8))
        attempt(bytes(b), 3000, 0)
    assert (attempt(good, 3000, 0) == lv).all()
