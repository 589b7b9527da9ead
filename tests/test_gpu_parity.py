"""GPU suite (B200): the CUDA path, called through the C-ABI shims, against the CPU oracle on
the same seeded inputs -- bit-exact for indices / masks / integers / bytes, <= 1e-5 relative
for fp32 features (exact for +-1 tables)."""
import numpy as np
import pytest
import torch

from conftest import R2, R3, R16

pytestmark = pytest.mark.gpu


def T(a, dev, dt=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    return t if dt is None else t.to(dt)


def gpu_fwd(dev, x, tab, offs, res, L, Rb=128, vxl=None, ml=None, bits=False):
    from cnc_b200 import _gridencoder as G

    N, D = x.shape
    F = tab.shape[1]
    out = torch.full((L, N, F), float("nan"), device=dev)
    tx, tt = T(x, dev), T(tab, dev)
    to, tr = T(np.asarray(offs, np.int32), dev), T(np.asarray(res, np.int32), dev)
    tv = None if vxl is None else T(vxl.astype(np.bool_), dev)
    tm = None if ml is None else T(ml.astype(np.int32), dev)
    if bits:
        b = G.sign_pack(tt)
        G.grid_encode_forward_bits(tx, b, to, tr, out, N, D, F, L, Rb, tv, tm)
    else:
        G.grid_encode_forward(tx, tt, to, tr, out, N, D, F, L, 0, Rb, 0.0, None, tv, tm)
    torch.cuda.synchronize()
    return out.cpu().numpy()


def cfg1_points():
    rng = np.random.default_rng(0)
    x = rng.random((4096, 3), dtype=np.float32)
    corners = np.array([[i, j, k] for i in (0, 1) for j in (0, 1) for k in (0, 1)], np.float32)
    cells = (np.arange(64, dtype=np.float32)[:, None] / 64.0).repeat(3, 1)
    oob = np.array([[1.0000001, 0.5, 0.5], [0.5, -1e-7, 0.5]], np.float32)
    return np.concatenate([x, corners, cells, oob]).astype(np.float32)


def test_config1_forward_matches_oracle(cuda, oracle):
    """BASELINE config 1: L=16, F=2, T=2^14, 4096 random points (+corners, cell boundaries, OOB)."""
    x = cfg1_points()
    offs = oracle.grid_layout(3, R16, 14)
    rng = np.random.default_rng(1)
    tab = rng.uniform(-1e-4, 1e-4, (offs[-1], 2)).astype(np.float32)
    ref = oracle.grid_encode_fwd(x, tab, offs, R16, 16)
    got = gpu_fwd(cuda, x, tab, offs, R16, 16)
    # same rounding sequence on both sides -> expected bit-exact; the contract is 1e-5 relative
    np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-12)
    assert (got == ref).mean() > 0.999
    pm = np.where(rng.random(tab.shape) < 0.5, -1.0, 1.0).astype(np.float32)
    ref = oracle.grid_encode_fwd(x, pm, offs, R16, 16)
    np.testing.assert_array_equal(gpu_fwd(cuda, x, pm, offs, R16, 16), ref)
    np.testing.assert_array_equal(gpu_fwd(cuda, x, pm, offs, R16, 16, bits=True), ref)


def test_hash_indices_bit_exact_via_index_table(cuda, oracle):
    """a1 on the GPU: a table whose row r stores r (split over 2 features, exact in fp32) read at
    grid vertices returns the row index the kernel computed; compare with the oracle / reference python."""
    for D, res_list, log2T in ((3, R3, 19), (2, R2, 17)):
        offs = oracle.grid_layout(D, res_list, log2T)
        rows = np.arange(offs[-1], dtype=np.int64)
        tab = np.stack([(rows & 0xFFF).astype(np.float32), (rows >> 12).astype(np.float32)], 1)
        rng = np.random.default_rng(D)
        for l, res in enumerate(res_list):
            v = rng.integers(1, res - 1, (2000, D))
            x = ((v - 0.5) / (res - 2)).astype(np.float32)  # lands exactly on vertex v (frac == 0)
            out, dbg = oracle.grid_encode_fwd(x, tab, offs[l:l + 2], [res], 1, return_rows=True)
            got = gpu_fwd(cuda, x, tab, offs[l:l + 2], [res], 1)
            np.testing.assert_array_equal(got, out)
            exact = (np.abs(x * np.float32(res - 2) + np.float32(0.5) - v) == 0).all(1)
            idx = got[0][:, 0].astype(np.int64) + (got[0][:, 1].astype(np.int64) << 12) - offs[l]
            want = oracle.grid_rows(v.astype(np.uint32), int(offs[l + 1] - offs[l]), res)
            assert exact.sum() > 100
            np.testing.assert_array_equal(idx[exact], want[exact].astype(np.int64))


@pytest.mark.parametrize("D,F", [(3, 8), (3, 1), (3, 4), (3, 16), (3, 32), (2, 8), (2, 2)])
def test_forward_variants(cuda, oracle, D, F):
    rng = np.random.default_rng(10 * D + F)
    res = R3[:7] if D == 3 else R2
    offs = oracle.grid_layout(D, res, 15)
    x = rng.random((3001, D), dtype=np.float32)
    tab = rng.normal(size=(offs[-1], F)).astype(np.float32)
    ref = oracle.grid_encode_fwd(x, tab, offs, res, len(res))
    got = gpu_fwd(cuda, x, tab, offs, res, len(res))
    np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-6)
    pm = np.sign(tab).astype(np.float32)
    pm[pm == 0] = 1
    np.testing.assert_array_equal(gpu_fwd(cuda, x, pm, offs, res, len(res), bits=True),
                                  oracle.grid_encode_fwd(x, pm, offs, res, len(res)))


def ball_occupancy(Rb=128, radius=1.0, half=1.5):
    c = (np.arange(Rb) + 0.5) / Rb * 2 * half - half
    X, Y, Z = np.meshgrid(c, c, c, indexing="ij")
    return (X * X + Y * Y + Z * Z <= radius * radius)


def test_forward_with_occupancy_and_per_point_levels(cuda, oracle):
    rng = np.random.default_rng(5)
    vx = ball_occupancy(32)
    offs = oracle.grid_layout(3, R3, 19)
    tab = np.where(rng.random((offs[-1], 8)) < 0.5, -1.0, 1.0).astype(np.float32)
    x = rng.random((5000, 3), dtype=np.float32)
    ref = oracle.grid_encode_fwd(x, tab, offs[2:6], R3[2:5], 3, Rb=32, binary_vxl=vx)
    np.testing.assert_array_equal(gpu_fwd(cuda, x, tab, offs[2:6], R3[2:5], 3, Rb=32, vxl=vx), ref)
    ml = rng.integers(0, 9, 5000).astype(np.int32)
    ref = oracle.grid_encode_fwd(x, tab, offs, R3, 3, Rb=32, binary_vxl=vx, min_level_id=ml)
    np.testing.assert_array_equal(gpu_fwd(cuda, x, tab, offs, R3, 3, Rb=32, vxl=vx, ml=ml), ref)
    np.testing.assert_array_equal(gpu_fwd(cuda, x, tab, offs, R3, 3, Rb=32, vxl=vx, ml=ml, bits=True), ref)
    # 2D with a 2D occupancy (context_model_2D path, utils_bpp_acc.py:551)
    vx2 = vx.any(2)
    offs2 = oracle.grid_layout(2, R2, 17)
    tab2 = np.where(rng.random((offs2[-1], 8)) < 0.5, -1.0, 1.0).astype(np.float32)
    x2 = rng.random((5000, 2), dtype=np.float32)
    ref = oracle.grid_encode_fwd(x2, tab2, offs2[0:4], R2[0:3], 3, Rb=32, binary_vxl=vx2)
    np.testing.assert_array_equal(gpu_fwd(cuda, x2, tab2, offs2[0:4], R2[0:3], 3, Rb=32, vxl=vx2), ref)


@pytest.mark.parametrize("D,F", [(3, 8), (3, 2), (2, 8), (3, 1)])
def test_backward_matches_oracle(cuda, oracle, D, F):
    from cnc_b200 import _gridencoder as G

    rng = np.random.default_rng(100 + F)
    res = R3[:8] if D == 3 else R2
    offs = oracle.grid_layout(D, res, 15)
    L, N = len(res), 4099
    x = rng.random((N, D), dtype=np.float32)
    g = rng.normal(size=(L, N, F)).astype(np.float32)
    vx = ball_occupancy(16) if D == 3 else ball_occupancy(16).any(2)
    for use_v in (False, True):
        ref = oracle.grid_encode_bwd(g, x, int(offs[-1]), offs, res, L, Rb=16, binary_vxl=vx if use_v else None)
        gt = torch.zeros(int(offs[-1]), F, device=cuda)
        G.grid_encode_backward(T(g, cuda), T(x, cuda), torch.empty(1, F, device=cuda), T(offs, cuda),
                               T(np.asarray(res, np.int32), cuda), gt, N, D, F, L, 0, 16, None, None,
                               T(vx, cuda) if use_v else None, None)
        got = gt.cpu().numpy()
        scale = np.abs(ref).max()
        np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-5 * scale)  # float atomics: order differs
        assert ((got != 0) == (ref != 0)).all()


def test_backward_is_adjoint_of_forward_full_size(cuda, oracle):
    """size-independent property at the product size: <fwd(table), g> == <table, bwd(g)>."""
    from cnc_b200 import _gridencoder as G

    torch.manual_seed(0)
    offs = T(oracle.grid_layout(3, R3, 19), cuda)
    res = T(np.asarray(R3, np.int32), cuda)
    N, L, F = 262144, 12, 8
    x = torch.rand(N, 3, device=cuda)
    tab = torch.randn(int(offs[-1]), F, device=cuda)
    g = torch.randn(L, N, F, device=cuda)
    out = torch.empty(L, N, F, device=cuda)
    G.grid_encode_forward(x, tab, offs, res, out, N, 3, F, L, 0, 128, 0.0)
    gt = torch.zeros_like(tab)
    G.grid_encode_backward(g, x, tab, offs, res, gt, N, 3, F, L, 0, 128)
    lhs = (out.double() * g.double()).sum().item()
    rhs = (gt.double() * tab.double()).sum().item()
    assert abs(lhs - rhs) <= 1e-5 * abs(lhs) + 1e-3
    # linearity: fwd(2*table) == 2*fwd(table) exactly (power-of-two scaling commutes with rounding)
    out2 = torch.empty_like(out)
    G.grid_encode_forward(x, tab * 2, offs, res, out2, N, 3, F, L, 0, 128, 0.0)
    assert torch.equal(out2, out * 2)


def test_ste_and_sign_bits(cuda, oracle, golden):
    from cnc_b200 import _gridencoder as G
    from cnc_b200.gridencoder import STE_binary

    p = T(golden["ste_in"], cuda).requires_grad_(True)
    y = STE_binary.apply(p)
    y.backward(T(golden["ste_gin"], cuda))
    np.testing.assert_array_equal(y.detach().cpu().numpy(), golden["ste_out"])
    np.testing.assert_array_equal(p.grad.cpu().numpy(), golden["ste_gout"])
    rng = np.random.default_rng(3)
    for rows, F in ((1000, 8), (1001 * 8, 2), (333 * 8, 1), (40, 32), (7 * 8, 4)):
        v = rng.normal(size=(rows, F)).astype(np.float32)
        v[0, 0] = 0.0
        v[1, 0] = -0.0
        bits = G.sign_pack(T(v, cuda))
        back = G.sign_unpack(bits, rows, F).cpu().numpy()
        np.testing.assert_array_equal(back, oracle.ste_binary(v))


def test_query_mask_bit_exact(cuda, oracle):
    from cnc_b200 import pack_and_align as P

    rng = np.random.default_rng(8)
    vx = ball_occupancy(128)
    for res in (44, 148, 514):
        pts = rng.integers(0, res, (20000, 3)).astype(np.int16)
        m_ref, o_ref = oracle.query_mask(pts, vx, resolution=res)
        mask = torch.zeros(len(pts), dtype=torch.int16, device=cuda)
        ov = torch.zeros(len(pts), dtype=torch.int32, device=cuda)
        P.query_mask_3D(T(pts, cuda), T(vx, cuda), mask, ov, torch.tensor(res, device=cuda), len(pts))
        np.testing.assert_array_equal(mask.cpu().numpy(), m_ref)
        np.testing.assert_array_equal(ov.cpu().numpy(), o_ref)
        assert 0 < m_ref.mean() < 1
    # per-point resolutions (training path) + 2D
    lv = rng.integers(3, 12, 30000)
    rl = np.asarray(R3)[lv].astype(np.int64)
    pts = (rng.random((30000, 3)) * rl[:, None]).astype(np.int16)
    m_ref, o_ref = oracle.query_mask(pts, vx, resolution_list=rl)
    mask = torch.zeros(len(pts), dtype=torch.int16, device=cuda)
    ov = torch.zeros(len(pts), dtype=torch.int32, device=cuda)
    P.query_mask_3D_qlist(T(pts, cuda), T(vx, cuda), mask, ov, T(rl, cuda), len(pts))
    np.testing.assert_array_equal(mask.cpu().numpy(), m_ref)
    np.testing.assert_array_equal(ov.cpu().numpy(), o_ref)
    vx2 = vx.any(2)
    pts2 = rng.integers(0, 258, (20000, 2)).astype(np.int16)
    m_ref, o_ref = oracle.query_mask(pts2, vx2, resolution=258)
    mask = torch.zeros(len(pts2), dtype=torch.int16, device=cuda)
    ov = torch.zeros(len(pts2), dtype=torch.int32, device=cuda)
    P.query_mask_3D(T(pts2, cuda), T(vx2, cuda), mask, ov, 258, len(pts2))
    np.testing.assert_array_equal(mask.cpu().numpy(), m_ref)
    np.testing.assert_array_equal(ov.cpu().numpy(), o_ref)


def test_align_pack_and_segment_sum(cuda, oracle):
    from cnc_b200 import pack_and_align as P

    rng = np.random.default_rng(9)
    cnt = rng.integers(0, 40, 3000).astype(np.int64)
    cnt[17] = 57
    feat = rng.normal(size=(int(cnt.sum()), 8)).astype(np.float32)
    cs = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    packed = P.align_and_pack_forward(T(feat, cuda), T(cnt, cuda), T(cs, cuda), len(cnt), int(cnt.max()), 8, 0.0, 3)
    ref = oracle.align_pack_fwd(feat, cnt, 0.0)
    np.testing.assert_array_equal(packed.cpu().numpy(), ref)
    back = P.align_and_pack_backward(packed, T(feat, cuda), T(cnt, cuda), T(cs, cuda), len(cnt), int(cnt.max()), 8,
                                     int(cs[-1]), 3)
    np.testing.assert_array_equal(back.cpu().numpy(), feat)
    w = rng.random(len(feat)).astype(np.float32)
    got = P.segment_wsum(T(feat, cuda), T(cs, cuda), T(w, cuda)).cpu().numpy()
    want = np.zeros((len(cnt), 8), np.float32)
    for i in range(len(cnt)):
        acc = np.zeros(8, np.float32)
        for j in range(cs[i], cs[i + 1]):
            acc = (acc + feat[j] * w[j]).astype(np.float32)
        want[i] = acc
    np.testing.assert_array_equal(got, want)  # fixed left-to-right order -> bit exact
    # empty input
    e = P.align_and_pack_forward(torch.zeros(0, 8, device=cuda), torch.zeros(0, dtype=torch.int64, device=cuda),
                                 torch.zeros(1, dtype=torch.int64, device=cuda), 0, 0, 8, 0.0, 3)
    assert e.shape == (0, 0, 8)


def test_vote_planes(cuda, oracle):
    from cnc_b200 import _gridencoder as G

    rng = np.random.default_rng(11)
    res, T_, F = 130, 2 ** 15, 8
    tab = np.where(rng.random((T_, F)) < 0.4, -1.0, 1.0).astype(np.float32)
    pts = rng.integers(0, res, (200000, 3)).astype(np.int16)
    for axis in range(3):
        ref = oracle.vote_planes_fwd(pts, tab, res, T_, axis)
        out = torch.zeros(res - 2, res - 2, F, 2, device=cuda)
        G.cnt_np_embed(T(pts, cuda), T(tab, cuda), out, len(pts), res, F, T_, axis)
        np.testing.assert_array_equal(out.cpu().numpy(), ref)
        s = ref.sum(-1, keepdims=True) + 1e-6
        g = rng.normal(size=ref.shape).astype(np.float32)
        gref = oracle.vote_planes_bwd(pts, tab, s, g, res, T_, axis)
        gt = torch.zeros(T_, F, device=cuda)
        G.cnt_np_embed_backward(T(pts, cuda), T(tab, cuda), T(s.astype(np.float32), cuda), T(g, cuda), gt, len(pts),
                                res, F, T_, axis)
        np.testing.assert_allclose(gt.cpu().numpy(), gref, rtol=1e-4, atol=1e-4 * np.abs(gref).max())


# ------------------------------------------------------------------------------------------ coder
def test_cdf_quantiser_bit_exact(cuda, oracle):
    from cnc_b200 import torchac as tac

    rng = np.random.default_rng(12)
    p = np.concatenate([rng.uniform(1e-6, 1 - 1e-6, 100000), [1e-6, 1 - 1e-6, 0.5, 0.25]]).astype(np.float32)
    c1 = tac.cdf_from_p(T(p, cuda)).cpu().numpy().view(np.uint16)
    np.testing.assert_array_equal(c1, oracle.cdf_from_p(p))


def _streams(rng, lens, skew):
    ps, syms = [], []
    for n in lens:
        p = np.clip(rng.beta(skew, skew, n), 1e-6, 1 - 1e-6).astype(np.float32)
        ps.append(p)
        syms.append((rng.random(n) < p).astype(np.uint8))
    return ps, syms


@pytest.mark.parametrize("skew", [0.05, 0.5, 5.0])
def test_coder_bytes_identical_to_oracle(cuda, oracle, skew):
    from cnc_b200 import torchac as tac

    rng = np.random.default_rng(int(skew * 100))
    lens = [0, 1, 2, 31, 32, 33, 1023, 1024, 1025, 4096, 70001, 200000]
    ps, syms = _streams(rng, lens, skew)
    c1s = [oracle.cdf_from_p(p) for p in ps]
    want = [oracle.ac_encode(c, s) for c, s in zip(c1s, syms)]
    got = tac.encode_streams([T(c.view(np.int16), cuda) for c in c1s], [T(s, cuda) for s in syms])
    for k, (a, b) in enumerate(zip(got, want)):
        assert a == b, f"stream {k} (n={lens[k]}) differs: {len(a)} vs {len(b)} bytes"
    dec = tac.decode_streams([T(c.view(np.int16), cuda) for c in c1s], want)
    for d, s in zip(dec, syms):
        np.testing.assert_array_equal(d.cpu().numpy(), s)


def test_coder_extreme_probabilities_and_long_pending_runs(cuda, oracle):
    from cnc_b200 import torchac as tac

    n = 50000
    # improbable symbols at p = 1e-6 (17+ bits each) and near-1/2 runs that stress E3 underflow
    for c1v, symv in ((1, 1), (1, 0), (65535, 0), (65535, 1), (32768, 0), (32768, 1), (32767, 1)):
        c1 = np.full(n, c1v, np.uint16)
        sym = np.full(n, symv, np.uint8)
        want = oracle.ac_encode(c1, sym)
        got = tac.encode_streams([T(c1.view(np.int16), cuda)], [T(sym, cuda)])[0]
        assert got == want
        dec = tac.decode_streams([T(c1.view(np.int16), cuda)], [want])[0]
        np.testing.assert_array_equal(dec.cpu().numpy(), sym)
    rng = np.random.default_rng(0)
    c1 = rng.choice(np.array([32767, 32768, 32769], np.uint16), n)
    sym = (np.arange(n) % 2).astype(np.uint8)
    want = oracle.ac_encode(c1, sym)
    assert tac.encode_streams([T(c1.view(np.int16), cuda)], [T(sym, cuda)])[0] == want


def _underflow_stream(n_straddle, n_tail, seed=0):
    """(c1, sym) that keeps the interval straddling 1/2 while it narrows: every symbol adds ~15 pending (E3) bits,
    so the run gets far beyond the packer's fast path (SLOW_PEND = 1024 in csrc/coder.cu)."""
    low, high, pending, max_pending = 0, 2 ** 32 - 1, 0, 0
    cs, ss = [], []
    for i in range(n_straddle):
        span = high - low + 1
        if i % 2 == 0:   # s = 1: raise low to just below 1/2
            c, s = ((2 ** 31 - 1 - low) << 16) // span, 1
        else:            # s = 0: lower high to just above 1/2
            c, s = -((-((2 ** 31 - low + 1) << 16)) // span), 0
        c = min(max(c, 1), 65535)
        t = (span * c) >> 16
        if s:
            low = low + t
        else:
            high = low + t - 1
        while True:
            if high < 2 ** 31 or low >= 2 ** 31:
                low, high, pending = (low << 1) & 0xFFFFFFFF, ((high << 1) | 1) & 0xFFFFFFFF, 0
            elif low >= 2 ** 30 and high < 3 * 2 ** 30:
                low, high, pending = (low << 1) & 0x7FFFFFFF, ((high << 1) | 0x80000001) & 0xFFFFFFFF, pending + 1
            else:
                break
        max_pending = max(max_pending, pending)
        cs.append(c)
        ss.append(s)
    rng = np.random.default_rng(seed)
    c1 = np.concatenate([np.array(cs, np.uint16), rng.integers(1, 65536, n_tail).astype(np.uint16)])
    sym = np.concatenate([np.array(ss, np.uint8), rng.integers(0, 2, n_tail).astype(np.uint8)])
    return c1, sym, max_pending


@pytest.mark.parametrize("n_straddle", [40, 200, 5000])
def test_coder_pending_runs_beyond_the_fast_path(cuda, oracle, n_straddle):
    from cnc_b200 import torchac as tac

    for n_tail in (0, 3000):
        c1, sym, max_pending = _underflow_stream(n_straddle, n_tail, seed=n_straddle)
        assert max_pending > 12 * n_straddle   # 40 -> below SLOW_PEND, 200 -> above, 5000 -> several 1024-bit chunks
        want = oracle.ac_encode(c1, sym)
        got = tac.encode_streams([T(c1.view(np.int16), cuda)], [T(sym, cuda)])[0]
        assert got == want
        dec = tac.decode_streams([T(c1.view(np.int16), cuda)], [want])[0]
        np.testing.assert_array_equal(dec.cpu().numpy(), sym)


def test_coder_roundtrip_full_size(cuda, oracle):
    """size-independent property at the product size: 33 streams, ~4e7 symbols, decode(encode(x)) == x,
    and the largest stream is byte-identical to the oracle."""
    from cnc_b200 import torchac as tac

    torch.manual_seed(1)
    lens = [5832 * 8, 13824 * 8, 35944 * 8, 85184 * 8, 205384 * 8, 512000 * 8] + [524288 * 8] * 3 + \
           [504198 * 8, 20090 * 8] + [197258 * 8] * 2 + [129772 * 8] + [77216 * 8] * 6 + [60992 * 8] + \
           [16904 * 8, 66568 * 8, 131072 * 8, 131072 * 8] * 3
    assert len(lens) == 33
    ps = [torch.rand(n, device=cuda).clamp_(1e-6, 1 - 1e-6) for n in lens]
    syms = [(torch.rand(n, device=cuda) < p).to(torch.uint8) for n, p in zip(lens, ps)]
    c1s = [tac.cdf_from_p(p) for p in ps]
    data = tac.encode_streams(c1s, syms)
    dec = tac.decode_streams(c1s, data)
    for d, s in zip(dec, syms):
        assert torch.equal(d, s)
    k = 6
    want = oracle.ac_encode(c1s[k].cpu().numpy().view(np.uint16), syms[k].cpu().numpy())
    assert data[k] == want
    total_bits = sum(len(b) for b in data) * 8
    ideal = sum((-(torch.where(s.bool(), p.log2(), (1 - p).log2())).sum().item()) for p, s in zip(ps, syms))
    assert total_bits <= ideal * 1.001 + 33 * 64


def test_torchac_shaped_entry_points(cuda, oracle):
    from cnc_b200 import torchac as tac

    rng = np.random.default_rng(21)
    p = np.clip(rng.random(10000), 1e-6, 1 - 1e-6).astype(np.float32)
    x = np.where(rng.random(10000) < p, 1.0, -1.0).astype(np.float32)
    pt = T(p, cuda)
    cdf = torch.cat([torch.zeros_like(pt)[:, None], (1 - pt)[:, None], torch.ones_like(pt)[:, None]], -1)
    sym = ((T(x, cuda) + 1) // 2).to(torch.int16)
    data = tac.encode_float_cdf(cdf, sym, check_input_bounds=True)
    assert data == oracle.encode_float_p(x, p)
    out = tac.decode_float_cdf(cdf, data)
    np.testing.assert_array_equal(out.cpu().numpy().astype(np.float32) * 2 - 1, x)


def test_grid_encoder_module_matches_oracle(cuda, oracle):
    """operator boundary (ngp.py:228-315): forward / forward_diff_levels / forward_given_params + autograd."""
    from cnc_b200.gridencoder import GridEncoder

    torch.manual_seed(3)
    enc = GridEncoder(num_dim=3, n_features=8, resolutions_list=R3, log2_hashmap_size=19, ste_binary=True).to(cuda)
    with torch.no_grad():
        enc.params.copy_(torch.randn_like(enc.params) * 0.7)
    x = torch.rand(2, 1000, 3, device=cuda)
    y = enc(x)
    assert y.shape == (2, 1000, 96)
    tab = oracle.ste_binary(enc.params.detach().cpu().numpy())
    offs, res = enc.offsets_list.cpu().numpy(), enc.resolutions_list.cpu().numpy()
    ref = oracle.grid_encode_fwd(x.view(-1, 3).cpu().numpy(), tab, offs, res, 12)  # [L,N,F]
    np.testing.assert_array_equal(y.detach().view(-1, 12, 8).permute(1, 0, 2).cpu().numpy(), ref)
    # partial levels + gradient through STE
    y2 = enc(x.view(-1, 3), 3, 6)
    np.testing.assert_array_equal(y2.detach().view(-1, 3, 8).permute(1, 0, 2).cpu().numpy(), ref[3:6])
    g = torch.randn_like(y2)
    y2.backward(g)
    gref = oracle.grid_encode_bwd(g.view(-1, 3, 8).permute(1, 0, 2).contiguous().cpu().numpy(),
                                  x.view(-1, 3).cpu().numpy(), int(offs[-1]), offs[3:7], res[3:6], 3)
    p = enc.params.detach().cpu().numpy()
    gref = gref * (np.abs(p) <= 1)
    np.testing.assert_allclose(enc.params.grad.cpu().numpy(), gref, rtol=1e-5, atol=1e-5 * np.abs(gref).max())
    # the sign table follows in-place parameter updates
    with torch.no_grad():
        enc.params.neg_()
    np.testing.assert_array_equal(enc(x).detach().view(-1, 12, 8).permute(1, 0, 2).cpu().numpy(), -ref)
    # per-point levels
    ml = torch.randint(0, 9, (2000,), device=cuda, dtype=torch.int32)
    vx = torch.from_numpy(ball_occupancy(128)).to(cuda)
    y3 = enc.forward_diff_levels(x.view(-1, 3), ml, 3, binary_vxl=vx, PV=1001)
    ref3 = oracle.grid_encode_fwd(x.view(-1, 3).cpu().numpy(), -tab, offs, res, 3, binary_vxl=vx.cpu().numpy(),
                                  min_level_id=ml.cpu().numpy())
    np.testing.assert_array_equal(y3.detach().view(-1, 3, 8).permute(1, 0, 2).cpu().numpy(), ref3)
    # forward_given_params: single 2D level, raw fp32 table (vote-fraction plane)
    enc2 = GridEncoder(num_dim=2, n_features=8, resolutions_list=R2, log2_hashmap_size=17, ste_binary=True).to(cuda)
    plane = torch.rand(130 * 130, 8, device=cuda)
    o2 = torch.tensor([0, 130 * 130], dtype=torch.int32, device=cuda)
    r2 = torch.tensor([130], dtype=torch.int32, device=cuda)
    xy = torch.rand(3000, 2, device=cuda)
    y4 = enc2.forward_given_params(xy, o2, r2, plane)
    ref4 = oracle.grid_encode_fwd(xy.cpu().numpy(), plane.cpu().numpy(), [0, 130 * 130], [130], 1)[0]
    np.testing.assert_allclose(y4.detach().cpu().numpy(), ref4, rtol=1e-5, atol=1e-7)
