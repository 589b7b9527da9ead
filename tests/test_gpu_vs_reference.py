"""GPU suite (B200): our kernels against the reference's OWN CUDA kernels, compiled unmodified
from /root/reference into oracle/_ref by oracle/build_ref.py (the .so files travel with the repo
snapshot).  This pins K1/K2/K4-K9 parity on the real reference binary, not only on the C oracle."""
import numpy as np
import pytest
import torch

from conftest import R2, R3
from test_gpu_parity import T, ball_occupancy

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref_grid(cuda):
    from oracle import ref_ext

    m = ref_ext.load("_gridencoder")
    if m is None:
        pytest.skip("oracle/_ref/_gridencoder not built")
    return m


@pytest.fixture(scope="module")
def ref_pack(cuda):
    from oracle import ref_ext

    m = ref_ext.load("pack_and_align")
    if m is None:
        pytest.skip("oracle/_ref/pack_and_align not built")
    return m


@pytest.mark.parametrize("D,F,use_vxl,use_ml", [(3, 8, False, False), (3, 8, True, False), (3, 8, True, True),
                                                 (2, 8, False, False), (2, 8, True, False), (3, 2, False, False),
                                                 (3, 4, True, True)])
def test_forward_bit_exact_vs_reference_kernel(cuda, oracle, ref_grid, D, F, use_vxl, use_ml):
    from cnc_b200 import _gridencoder as G

    torch.manual_seed(D * 100 + F)
    res = R3 if D == 3 else R2
    offs = T(oracle.grid_layout(D, res, 19 if D == 3 else 17), cuda)
    rl = T(np.asarray(res, np.int32), cuda)
    N = 100003
    L = 3 if use_ml else len(res)
    x = torch.rand(N, D, device=cuda)
    x[:8] = torch.tensor([[i, j, k][:D] for i in (0., 1.) for j in (0., 1.) for k in (0., 1.)], device=cuda)
    tab = torch.randn(int(offs[-1]), F, device=cuda)
    vx = T(ball_occupancy(128) if D == 3 else ball_occupancy(128).any(2), cuda) if use_vxl else None
    ml = torch.randint(0, len(res) - 2, (N,), device=cuda, dtype=torch.int32) if use_ml else None
    a = torch.empty(L, N, F, device=cuda)
    b = torch.empty(L, N, F, device=cuda)
    ref_grid.grid_encode_forward(x, tab, offs, rl, a, N, D, F, L, 0, 128, 0.0, None, vx, ml)
    G.grid_encode_forward(x, tab, offs, rl, b, N, D, F, L, 0, 128, 0.0, None, vx, ml)
    torch.cuda.synchronize()
    # Contract (BASELINE north_star): features within 1e-5 relative; indices / masks exact.  A wrong corner,
    # floor() or occupancy decision would show up as an O(1) difference with this randn table, so the
    # allclose below also pins the integer side against the reference binary.  Bit-identity of the fp32
    # features is NOT attainable in general: nvcc contracts `wn += w` with the last weight multiply into
    # FFMA differently in every <D,F> instantiation of the reference (seen in the SASS of oracle/_ref), so
    # the normaliser differs in the last bit for a fraction of the points; the fraction is asserted loosely.
    assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)
    frac = (a == b).float().mean().item()
    print(f"D={D} F={F} vxl={use_vxl} ml={use_ml}: {frac:.4f} of the features bit-identical to the reference binary")
    assert frac > 0.5
    # +-1 tables: every product is exact, only the normaliser's last bit can differ
    pm = torch.where(tab >= 0, 1.0, -1.0)
    ref_grid.grid_encode_forward(x, pm, offs, rl, a, N, D, F, L, 0, 128, 0.0, None, vx, ml)
    G.grid_encode_forward(x, pm, offs, rl, b, N, D, F, L, 0, 128, 0.0, None, vx, ml)
    assert torch.allclose(a, b, rtol=1e-6, atol=2e-7)
    c = torch.empty_like(b)
    G.grid_encode_forward_bits(x, G.sign_pack(tab), offs, rl, c, N, D, F, L, 128, vx, ml)
    assert torch.equal(b, c)  # our fp32-table and 1-bit-table paths are the same arithmetic


@pytest.mark.parametrize("D,F,use_vxl", [(3, 8, False), (3, 8, True), (2, 8, False), (3, 2, False)])
def test_backward_vs_reference_kernel(cuda, oracle, ref_grid, D, F, use_vxl):
    from cnc_b200 import _gridencoder as G

    torch.manual_seed(7)
    res = R3 if D == 3 else R2
    offs = T(oracle.grid_layout(D, res, 19 if D == 3 else 17), cuda)
    rl = T(np.asarray(res, np.int32), cuda)
    N, L = 50000, len(res)
    x = torch.rand(N, D, device=cuda)
    g = torch.randn(L, N, F, device=cuda)
    tab = torch.zeros(int(offs[-1]), F, device=cuda)
    vx = T(ball_occupancy(128) if D == 3 else ball_occupancy(128).any(2), cuda) if use_vxl else None
    a, b = torch.zeros_like(tab), torch.zeros_like(tab)
    ref_grid.grid_encode_backward(g, x, tab, offs, rl, a, N, D, F, L, 0, 128, None, None, vx, None)
    G.grid_encode_backward(g, x, tab, offs, rl, b, N, D, F, L, 0, 128, None, None, vx, None)
    torch.cuda.synchronize()
    assert torch.equal(a != 0, b != 0)
    assert torch.allclose(a, b, rtol=1e-4, atol=1e-5 * a.abs().max().item())  # atomics: order differs in both


def test_query_mask_vs_reference_kernel(cuda, ref_pack):
    from cnc_b200 import pack_and_align as P

    torch.manual_seed(9)
    vx = T(ball_occupancy(128), cuda)
    for res in (18, 44, 148, 514):
        N = 200000
        pts = torch.randint(0, res, (N, 3), device=cuda).to(torch.int16)
        m0, o0 = torch.zeros(N, dtype=torch.int16, device=cuda), torch.zeros(N, dtype=torch.int32, device=cuda)
        m1, o1 = torch.zeros_like(m0), torch.zeros_like(o0)
        ref_pack.query_mask_3D(pts, vx, m0, o0, res, N)
        P.query_mask_3D(pts, vx, m1, o1, res, N)
        assert torch.equal(m0, m1)
        assert torch.equal(o0, o1), f"overlap ints differ at res {res}: {(o0 != o1).sum().item()} of {N}"
    lv = torch.randint(3, 12, (300000,), device=cuda)
    rl = torch.tensor(R3, device=cuda)[lv].to(torch.int64)
    pts = (torch.rand(300000, 3, device=cuda) * rl[:, None]).to(torch.int16)
    m0, o0 = torch.zeros(300000, dtype=torch.int16, device=cuda), torch.zeros(300000, dtype=torch.int32, device=cuda)
    m1, o1 = torch.zeros_like(m0), torch.zeros_like(o0)
    ref_pack.query_mask_3D_qlist(pts, vx, m0, o0, rl, 300000)
    P.query_mask_3D_qlist(pts, vx, m1, o1, rl, 300000)
    assert torch.equal(m0, m1) and torch.equal(o0, o1)
    vx2 = vx.any(2)
    pts2 = torch.randint(0, 1026, (200000, 2), device=cuda).to(torch.int16)
    m0, o0 = torch.zeros(200000, dtype=torch.int16, device=cuda), torch.zeros(200000, dtype=torch.int32, device=cuda)
    m1, o1 = torch.zeros_like(m0), torch.zeros_like(o0)
    ref_pack.query_mask_3D(pts2, vx2, m0, o0, 1026, 200000)
    P.query_mask_3D(pts2, vx2, m1, o1, 1026, 200000)
    assert torch.equal(m0, m1) and torch.equal(o0, o1)


def test_align_pack_vs_reference_kernel(cuda, ref_pack):
    from cnc_b200 import pack_and_align as P

    torch.manual_seed(10)
    cnt = torch.randint(0, 60, (5000,), device=cuda)
    cs = torch.cat([torch.zeros(1, dtype=torch.int64, device=cuda), cnt.cumsum(0)])
    feat = torch.randn(int(cs[-1]), 8, device=cuda)
    M = int(cnt.max())
    for dim in (2, 3):
        a = ref_pack.align_and_pack_forward(feat, cnt, cs, 5000, M, 8, 0.0, dim)
        b = P.align_and_pack_forward(feat, cnt, cs, 5000, M, 8, 0.0, dim)
        assert torch.equal(a, b)
        ga = ref_pack.align_and_pack_backward(a, feat, cnt, cs, 5000, M, 8, int(cs[-1]), dim)
        gb = P.align_and_pack_backward(a, feat, cnt, cs, 5000, M, 8, int(cs[-1]), dim)
        assert torch.equal(ga, gb) and torch.equal(gb, feat)


def test_vote_planes_vs_reference_kernel(cuda, ref_grid):
    from cnc_b200 import _gridencoder as G

    torch.manual_seed(11)
    res, Tn, F = 514, 2 ** 19, 8
    tab = torch.where(torch.rand(Tn, F, device=cuda) < 0.45, -1.0, 1.0)
    pts = torch.randint(0, res, (1000000, 3), device=cuda).to(torch.int16)
    for axis in range(3):
        a = torch.zeros(res - 2, res - 2, F, 2, device=cuda)
        b = torch.zeros_like(a)
        ref_grid.cnt_np_embed(pts, tab, a, len(pts), res, F, Tn, axis)
        G.cnt_np_embed(pts, tab, b, len(pts), res, F, Tn, axis)
        assert torch.equal(a, b)
        s = a.sum(-1, keepdim=True) + 1e-6
        g = torch.randn_like(a)
        ga, gb = torch.zeros_like(tab), torch.zeros_like(tab)
        ref_grid.cnt_np_embed_backward(pts, tab, s, g, ga, len(pts), res, F, Tn, axis)
        G.cnt_np_embed_backward(pts, tab, s, g, gb, len(pts), res, F, Tn, axis)
        assert torch.allclose(ga, gb, rtol=1e-4, atol=1e-4 * ga.abs().max().item())
