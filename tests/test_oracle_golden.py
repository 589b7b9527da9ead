"""CPU suite: pins the oracle against every golden vector the reference offers for the path
(reference python executed by tests/golden/make_golden.py + nerfacc docstring KATs) and checks
the domain properties the GPU tests rely on."""
import numpy as np
import pytest

from conftest import R2, R3, R16


def test_hash_rows_match_reference_python(golden, oracle):
    """a1: gridencoder.cu:61-87 restatement == examples/utils.py:492-511 (reference python run on CPU)."""
    cases = golden["hash_cases"]
    assert len(cases) == 12 + 4 + 16
    n_hashed = 0
    for k, (D, res, T) in enumerate(cases):
        pos = golden[f"hash_pos_{k}"].astype(np.uint32)
        rows = oracle.grid_rows(pos, int(T), int(res))
        np.testing.assert_array_equal(rows.astype(np.int64), golden[f"hash_idx_{k}"])
        if res ** D > T:
            n_hashed += 1
            assert T & (T - 1) == 0, "hashed levels must have power-of-two T (SURVEY 8c)"
    assert n_hashed >= 6 + 2 + 14


def test_table_layout_matches_reference(golden, oracle):
    for name, (D, res, log2T) in dict(xyz=(3, R3, 19), plane=(2, R2, 17), cfg1=(3, R16, 14)).items():
        offs = oracle.grid_layout(D, res, log2T)
        np.testing.assert_array_equal(offs, golden[f"layout_{name}_offsets"])
        np.testing.assert_array_equal(np.array(res, np.int32), golden[f"layout_{name}_res"])
        assert offs[-1] == golden[f"layout_{name}_rows"]
    assert oracle.grid_layout(3, R3, 19)[-1] == 4003896
    assert oracle.grid_layout(2, R2, 17)[-1] == 345616


def test_ste_binary(golden, oracle):
    np.testing.assert_array_equal(oracle.ste_binary(golden["ste_in"]), golden["ste_out"])
    x = golden["ste_in"]
    mask = (np.clip(x, -1, 1) == x).astype(np.float32)
    np.testing.assert_array_equal(golden["ste_gin"] * mask, golden["ste_gout"])


def test_freq_embed(golden, oracle):
    np.testing.assert_allclose(oracle.freq_embed(golden["embed_in"]), golden["embed_out"], atol=2e-5)


def test_nerfacc_scan_kats(golden, oracle):
    pk = oracle.pack_info(golden["kat_ray_indices_9"], 3)
    np.testing.assert_array_equal(pk, golden["kat_packed_info"])
    x = golden["kat_scan_in"]
    np.testing.assert_array_equal(oracle.packed_scan(x, pk, "sum", True), golden["kat_inclusive_sum"])
    np.testing.assert_array_equal(oracle.packed_scan(x, pk, "sum", False), golden["kat_exclusive_sum"])
    np.testing.assert_array_equal(oracle.packed_scan(x, pk, "prod", True), golden["kat_inclusive_prod"])
    np.testing.assert_array_equal(oracle.packed_scan(x, pk, "prod", False), golden["kat_exclusive_prod"])


def test_nerfacc_volrend_kats(golden, oracle):
    pk = oracle.pack_info(golden["kat_ray_indices_7"], 3)
    a = golden["kat_alphas"]
    trans = oracle.packed_scan(1 - a, pk, "prod", False)  # volrend.py:208
    np.testing.assert_allclose(trans, golden["kat_trans_from_alpha"], atol=1e-6)
    np.testing.assert_allclose(trans * a, golden["kat_weights_from_alpha"], atol=1e-6)
    r = oracle.render_from_density(golden["kat_t_starts"], golden["kat_t_ends"], golden["kat_sigmas"], pk)
    np.testing.assert_allclose(r["trans"], golden["kat_trans_from_density"], atol=6e-3)
    np.testing.assert_allclose(r["alphas"], golden["kat_alphas_from_density"], atol=6e-3)
    np.testing.assert_allclose(r["weights"], golden["kat_weights_from_density"], atol=6e-3)
    vis = (r["trans"] >= 0.3) & (r["alphas"] >= 0.2)  # volrend.py:479-482
    np.testing.assert_array_equal(vis.astype(np.uint8), golden["kat_visibility"])
    vis_a = (trans >= 0.3) & (a >= 0.2)
    np.testing.assert_array_equal(vis_a.astype(np.uint8), golden["kat_visibility"])


# ------------------------------------------------------------------------------------------ grid
def _cfg1_points():
    rng = np.random.default_rng(0)
    x = rng.random((4096, 3), dtype=np.float32)
    corners = np.array([[i, j, k] for i in (0, 1) for j in (0, 1) for k in (0, 1)], np.float32)
    cells = (np.arange(64, dtype=np.float32)[:, None] / 64.0).repeat(3, 1)
    return np.concatenate([x, corners, cells]).astype(np.float32)


def test_grid_forward_properties(oracle):
    x = _cfg1_points()
    offs = oracle.grid_layout(3, R16, 14)
    ones = np.ones((offs[-1], 2), np.float32)
    out, rows = oracle.grid_encode_fwd(x, ones, offs, R16, 16, return_rows=True)
    assert out.shape == (16, len(x), 2)
    # renormalised weights of the valid corners sum to 1
    interior = (rows >= 0).any(-1)
    np.testing.assert_allclose(out[interior], 1.0, atol=3e-7)
    assert (out[~interior] == 0).all()
    # rows stay inside their level
    for l in range(16):
        v = rows[l][rows[l] >= 0]
        assert v.max() < offs[l + 1] - offs[l]
    # x in {0,1}^3 sits half-way between the zeroed border vertex and the first interior one:
    # exactly one of the 8 corners survives and renormalisation makes its weight 1
    assert ((rows[:, 4096:4104] >= 0).sum(-1) == 1).all()
    assert (out[:, 4096:4104] == 1).all()
    # out of range -> zeros
    bad = np.array([[1.5, 0.5, 0.5], [0.5, -0.1, 0.5]], np.float32)
    assert (oracle.grid_encode_fwd(bad, ones, offs, R16, 16) == 0).all()
    # +-1 table: features are exact convex combinations, |f| <= 1
    rng = np.random.default_rng(1)
    pm = np.where(rng.random(ones.shape) < 0.5, -1.0, 1.0).astype(np.float32)
    o2 = oracle.grid_encode_fwd(x, pm, offs, R16, 16)
    assert np.abs(o2).max() <= 1.0 + 3e-7


def test_grid_backward_is_adjoint(oracle):
    rng = np.random.default_rng(2)
    x = rng.random((513, 3), dtype=np.float32)
    offs = oracle.grid_layout(3, R16[:8], 14)
    tab = rng.normal(size=(offs[-1], 2)).astype(np.float32)
    g = rng.normal(size=(8, 513, 2)).astype(np.float32)
    out = oracle.grid_encode_fwd(x, tab, offs, R16[:8], 8)
    gt = oracle.grid_encode_bwd(g, x, offs[-1], offs, R16[:8], 8)
    lhs = float((out.astype(np.float64) * g).sum())
    rhs = float((gt.astype(np.float64) * tab).sum())
    assert abs(lhs - rhs) <= 1e-4 * max(1.0, abs(lhs))


def test_grid_occupancy_mask_and_levels(oracle):
    rng = np.random.default_rng(3)
    x = rng.random((300, 3), dtype=np.float32)
    offs = oracle.grid_layout(3, R3, 19)
    tab = np.where(rng.random((offs[-1], 4)) < 0.5, -1.0, 1.0).astype(np.float32)
    full = np.ones((16, 16, 16), np.uint8)
    a = oracle.grid_encode_fwd(x, tab, offs[3:7], R3[3:6], 3)
    b = oracle.grid_encode_fwd(x, tab, offs[3:7], R3[3:6], 3, Rb=16, binary_vxl=full)
    np.testing.assert_array_equal(a, b)  # all-occupied grid == no mask
    empty = np.zeros((16, 16, 16), np.uint8)
    assert (oracle.grid_encode_fwd(x, tab, offs[3:7], R3[3:6], 3, Rb=16, binary_vxl=empty) == 0).all()
    # per-point min level == sliced lists
    ml = np.full(300, 3, np.int32)
    c = oracle.grid_encode_fwd(x, tab, offs, R3, 3, min_level_id=ml)
    np.testing.assert_array_equal(a, c)


def test_query_mask_properties(oracle):
    rng = np.random.default_rng(4)
    res = 148
    pts = rng.integers(0, res, (2000, 3)).astype(np.int16)
    full = np.ones((128, 128, 128), np.uint8)
    m, ov = oracle.query_mask(pts, full, resolution=res)
    assert (m == 1).all()
    # box is 2/(res-2) wide per axis (clipped at the grid border) -> <= 1000*(2*128/146)^3
    assert ov.max() <= int(1000 * (2 * 128 / 146) ** 3) + 1 and ov.min() > 0
    interior = ((pts > 2) & (pts < res - 3)).all(-1)
    assert abs(ov[interior].astype(np.float64) - 1000 * (2 * 128 / 146) ** 3).max() < 2.0
    m0, ov0 = oracle.query_mask(pts, np.zeros_like(full), resolution=res)
    assert (m0 == 0).all() and (ov0 == 0).all()
    m1, ov1 = oracle.query_mask(pts, full, resolution_list=np.full(2000, res, np.int64))
    np.testing.assert_array_equal(m, m1)
    np.testing.assert_array_equal(ov, ov1)


def test_align_pack_roundtrip(oracle):
    rng = np.random.default_rng(5)
    cnt = rng.integers(0, 7, 50)
    cnt[0] = 6
    feat = rng.normal(size=(cnt.sum(), 3)).astype(np.float32)
    packed = oracle.align_pack_fwd(feat, cnt, V=-7.0)
    assert packed.shape == (50, 6, 3)
    for i in range(50):
        assert (packed[i, cnt[i]:] == -7.0).all()
    np.testing.assert_array_equal(oracle.align_pack_bwd(packed, cnt), feat)


def test_vote_planes(oracle):
    rng = np.random.default_rng(6)
    res, T, F = 34, 2 ** 12, 4
    tab = np.where(rng.random((T, F)) < 0.5, -1.0, 1.0).astype(np.float32)
    g = np.stack(np.meshgrid(*[np.arange(res)] * 3, indexing="ij"), -1).reshape(-1, 3).astype(np.int16)
    for axis in range(3):
        out = oracle.vote_planes_fwd(g, tab, res, T, axis)
        # every interior column holds res-2 votes per channel
        np.testing.assert_array_equal(out.sum(-1), np.full((res - 2, res - 2, F), res - 2, np.float32))


# ------------------------------------------------------------------------------------------ coder
@pytest.mark.parametrize("n", [0, 1, 2, 7, 8, 9, 1000, 100003])
def test_coder_roundtrip(oracle, n):
    rng = np.random.default_rng(n)
    p = rng.uniform(0, 1, n).astype(np.float32).clip(1e-6, 1 - 1e-6)
    x = np.where(rng.random(n) < p, 1.0, -1.0).astype(np.float32)
    data = oracle.encode_float_p(x, p)
    np.testing.assert_array_equal(oracle.decode_float_p(p, data), x)
    if n >= 1000:
        bits = -(np.where(x > 0, np.log2(p), np.log2(1 - p))).sum()
        assert len(data) * 8 <= bits * 1.01 + 64  # within 1% of the ideal code length


def test_coder_extremes_and_cdf(oracle):
    # clamp range used by the reference (utils_bpp_acc.py:751,852)
    p = np.array([1e-6, 1 - 1e-6, 0.5, 0.25, 1e-6, 1 - 1e-6] * 50, np.float32)
    c1 = oracle.cdf_from_p(p)
    assert c1.min() >= 1 and c1.max() <= 65535
    assert c1[2] == 32768 and c1[3] == int(np.rint(np.float32(0.75) * np.float32(65534))) + 1
    for sym in (np.zeros(300, np.uint8), np.ones(300, np.uint8)):
        data = oracle.ac_encode(c1, sym)
        np.testing.assert_array_equal(oracle.ac_decode(c1, data), sym)
    # hand-traced vectors (Appendix B): p=1/2, one symbol -> E1/E2 emits the symbol bit, the
    # finish step emits 0 then one pending 1; the empty stream is just the finish step
    assert oracle.ac_encode(np.array([0x8000], np.uint16), np.array([0], np.uint8)) == bytes([0b00100000])
    assert oracle.ac_encode(np.array([0x8000], np.uint16), np.array([1], np.uint8)) == bytes([0b10100000])
    assert oracle.ac_encode(np.zeros(0, np.uint16), np.zeros(0, np.uint8)) == bytes([0b01000000])


def test_sh16(oracle):
    rng = np.random.default_rng(7)
    d = rng.normal(size=(100, 3))
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    o = oracle.sh16((d + 1) / 2, fp16_round=False)
    np.testing.assert_allclose(o[:, 0], 0.28209479, rtol=1e-6)
    np.testing.assert_allclose(o[:, 2], 0.48860251 * d[:, 2], atol=1e-6)
    # orthonormality-ish: sum_l<=3 of Y^2 is (16)/(4 pi) on the unit sphere
    np.testing.assert_allclose((o.astype(np.float64) ** 2).sum(1), 16 / (4 * np.pi), rtol=1e-4)
    o16 = oracle.sh16((d + 1) / 2, fp16_round=True)
    assert np.abs(o16 - o).max() < 2e-3 and (o16 == o16.astype(np.float16)).all()


def test_marching_oracle(oracle):
    # rays through a fully occupied 8^3 grid over [-1,1]^3: uniform samples, continuous intervals
    o = np.array([[-3.0, 0.1, 0.2], [0.0, 0.0, -5.0], [5, 5, 5]], np.float32)
    d = np.array([[1.0, 0, 0], [0, 0, 1.0], [1.0, 0, 0]], np.float32)
    bins = np.ones((1, 8, 8, 8), np.uint8)
    t0, t1, ri, pk, term = oracle.traverse_grids(o, d, bins, [-1, -1, -1, 1, 1, 1], step_size=0.05)
    assert pk[2, 1] == 0  # the third ray misses the box
    for r, (tin, tout) in enumerate([(2.0, 4.0), (4.0, 6.0)]):
        s, c = pk[r]
        assert 38 <= c <= 41
        np.testing.assert_allclose(t1[s:s + c] - t0[s:s + c], 0.05, atol=1e-5)
        np.testing.assert_allclose(t0[s + 1:s + c], t1[s:s + c - 1], atol=0)  # continuous
        mid = (t0[s:s + c] + t1[s:s + c]) / 2
        assert mid.min() >= tin - 1e-4 and mid.max() <= tout + 1e-4
    # empty grid -> no samples
    assert oracle.traverse_grids(o, d, np.zeros_like(bins), [-1, -1, -1, 1, 1, 1], step_size=0.05)[3][:, 1].sum() == 0


def _renorm_loop(low, high):
    """the oracle's bit-at-a-time renormalisation (oracle/cnc_oracle.c cnc_o_ac_encode), vectorised: -> (low, high, S)"""
    low, high = low.astype(np.uint64), high.astype(np.uint64)
    S = np.zeros(low.shape, np.int64)
    M = np.uint64(0xFFFFFFFF)
    for _ in range(34):
        e12 = (high < 0x80000000) | (low >= 0x80000000)
        e3 = ~e12 & (low >= 0x40000000) & (high < 0xC0000000)
        any_ = e12 | e3
        l1 = (low << np.uint64(1)) & M
        h1 = ((high << np.uint64(1)) | np.uint64(1)) & M
        l3 = l1 & np.uint64(0x7FFFFFFF)
        h3 = h1 | np.uint64(0x80000000)
        low = np.where(e12, l1, np.where(e3, l3, low))
        high = np.where(e12, h1, np.where(e3, h3, high))
        S += any_
    return low, high, S


def test_closed_form_renormalisation_equals_reference_loop():
    """The GPU coder replaces the E1/E2/E3 loop by S = clz(r2) - 1 + [top bits differ by <= 1] (csrc/coder.cu):
    same shift count and same (low, high - low) afterwards, on random and on boundary-hugging intervals."""
    rng = np.random.default_rng(5)
    n = 400000
    B = (rng.integers(0, 2 ** 32, n, dtype=np.uint64) >> rng.integers(0, 32, n).astype(np.uint64)) << rng.integers(0, 32, n).astype(np.uint64)
    B = np.where(rng.random(n) < 0.4, rng.choice(np.array([2 ** 31, 2 ** 30, 3 * 2 ** 30], np.uint64), n), B & np.uint64(0xFFFFFFFF))
    x = rng.integers(0, 2 ** 32, n, dtype=np.uint64) >> rng.integers(1, 33, n).astype(np.uint64)
    y = rng.integers(0, 2 ** 32, n, dtype=np.uint64) >> rng.integers(1, 33, n).astype(np.uint64)
    lo = np.maximum(B.astype(np.int64) - x.astype(np.int64), 0)
    hi = np.minimum(B.astype(np.int64) + y.astype(np.int64), 2 ** 32 - 1)
    lo2 = np.concatenate([lo, rng.integers(0, 2 ** 31, n)])
    hi2 = np.concatenate([hi, rng.integers(2 ** 31, 2 ** 32, n)])
    keep = hi2 - lo2 >= 1
    lo2, hi2 = lo2[keep], hi2[keep]
    r2 = hi2 - lo2
    c0 = 32 - np.floor(np.log2(r2.astype(np.float64))).astype(np.int64) - 1
    c0 = np.where((r2 >> (31 - c0)) == 0, c0 + 1, np.where((r2 >> (31 - c0)) > 1, c0 - 1, c0))   # exact clz
    sh = 31 - c0
    S = c0 - 1 + (((hi2 >> sh) - (lo2 >> sh)) <= 1)
    low_n = ((lo2 << S) & 0x7FFFFFFF)
    r_n = ((r2 << S) | ((1 << S) - 1)) & 0xFFFFFFFF
    l, h, Sref = _renorm_loop(lo2, hi2)
    np.testing.assert_array_equal(S, Sref)
    np.testing.assert_array_equal(low_n.astype(np.uint64), l)
    np.testing.assert_array_equal(r_n.astype(np.uint64), h - l)
    assert (Sref >= 10).sum() > 1000 and (Sref == 0).sum() > 1000
