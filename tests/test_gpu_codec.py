"""GPU side of the codec boundary: nerfq_b200.deepcabac with DEVICE='cuda' (the default) against DEVICE='host', call for
call as nnc_core/approximator/baseline.py makes them, and a whole NDU (weight_scaling, bias, weight) quantised on the GPU,
entropy-coded and decoded on the host, dequantised on the GPU -- the levels the fused MLP kernels consume are the levels
the bitstream carries."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_quant_dequant_device_equals_host():
    from nerfq_b200 import deepcabac as dc
    rng = np.random.default_rng(2)
    tensors = {"w": (rng.standard_normal((256, 319)) * 0.1).astype(np.float32), "b": (rng.standard_normal((256,)) * 0.01).astype(np.float32),
               "ls": (1.0 + 1e-3 * rng.standard_normal((256,))).astype(np.float32), "clip": np.array([3e4, -1.0, 0.3], dtype=np.float32)}
    for qp in (-38, -29, -20, -10, -75):
        for name, w in tensors.items():
            res = {}
            for device in ("cuda", "host"):
                dc.DEVICE = device
                try:
                    lv = np.zeros(w.shape, dtype=np.int32)
                    enc = dc.Encoder()
                    enc.initCtxModels(10, 0)
                    used = enc.quantLayer(w, lv, 0, 2, qp, 0.0, 10, 0)
                    rec = np.zeros(w.shape, dtype=np.float32)
                    dc.Decoder().dequantLayer(rec, lv, 2, used, 0)
                    res[device] = (used, lv, rec)
                finally:
                    dc.DEVICE = "cuda"
            assert res["cuda"][0] == res["host"][0], (qp, name)
            assert (res["cuda"][1] == res["host"][1]).all() and (res["cuda"][2] == res["host"][2]).all(), (qp, name)


def test_block_ndu_gpu_levels_through_the_host_coder():
    """One Linear layer as the reference codes it (coder/baseline.py:5-20; order weight_scaling, bias, weight in one
    codeword): levels from the GPU quantiser -> host encoder -> host decoder -> GPU dequantiser -> packed network operands."""
    from nerfq_b200 import deepcabac as dc, ops
    dev = torch.device("cuda:0")
    torch.manual_seed(1)
    w = torch.randn(256, 256, device=dev) * 0.08
    b = torch.randn(256, device=dev) * 0.01
    ls = 1.0 + 0.02 * torch.randn(256, device=dev)
    parts = [(ls, -75), (b, -75), (w, -20)]
    lv_gpu, _ = ops.quantize_batch([p[0].contiguous() for p in parts], [p[1] for p in parts], 2)
    enc = dc.Encoder()
    for (t, qp), lv in zip(parts, lv_gpu):
        enc.iae_v(8, qp)
        enc.initCtxModels(10, 1)
        enc.encodeLayer(lv.cpu().numpy(), 0, 0)
    bs = enc.finish().tobytes()
    assert len(bs) < 0.2 * 4 * sum(p[0].numel() for p in parts)
    dec = dc.Decoder()
    dec.setStream(bytearray(bs + b"next unit"))
    for (t, qp), lv in zip(parts, lv_gpu):
        assert dec.iae_v(8) == qp
        dec.initCtxModels(10)
        out = np.zeros(tuple(t.shape), dtype=np.int32)
        dec.decodeLayer(out, 0, 0)
        assert (out == lv.cpu().numpy()).all()
        rec = np.zeros(out.shape, dtype=np.float32)
        dec.dequantLayer(rec, out, 2, qp, 0)
        assert np.abs(rec - t.cpu().numpy()).max() <= 0.5 * ops.stepsize(qp, 2) * (1 + 1e-6)
    assert dec.finish() == len(bs)


def test_gpu_levels_match_the_frozen_stream_fixture():
    """tests/golden/coder_stream.npz: the uniform-quantiser levels stored there (made on the host) come out of the CUDA kernel bit
    for bit, and the stream's decoded levels dequantise on the GPU to level * delta exactly."""
    import os
    from nerfq_b200 import deepcabac as dc, ops
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "coder_stream.npz"))
    dev = torch.device("cuda:0")
    for key, dq, qp in zip([str(k) for k in g["layer_keys"]], g["layer_dq"].tolist(), g["layer_qp"].tolist()):
        w = g[key[len("levels_"):key.rindex("_dq")]]
        if dq == 0:
            lv, used = ops.quantize_urq(torch.from_numpy(w).to(dev).contiguous(), qp, 2)
            assert int(used) == qp and (lv.cpu().numpy() == g[key]).all(), key
        rec = ops.dequantize(torch.from_numpy(g[key]).to(dev).contiguous(), qp, 2).cpu().numpy()
        d = np.float32(dc.host_lib().nncabac_stepsize(qp, 2))
        assert (rec == g[key].astype(np.float32) * d).all(), key
