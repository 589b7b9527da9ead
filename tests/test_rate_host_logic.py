"""CPU suite: the host-side pieces of the rate term that replace per-level autograd slicing (cnc_b200/context_models.py:
`_LevelSums`, `level_entropies`) against the per-level function that mirrors utils_bpp_acc.py:472-486, values and gradients."""
import types

import torch

from cnc_b200.context_models import CNC_context_models, _CtxMLP3, _LevelSums, _level_const


def _stub(offs):
    return types.SimpleNamespace(offs=offs)


def test_level_sums_forward_and_backward_match_slicing():
    torch.manual_seed(0)
    offs = [0, 5, 12, 13, 40]
    p = torch.randn(40, 8, dtype=torch.float64, requires_grad=True)
    s = _LevelSums.apply(p, offs)
    want = torch.stack([p[a:b].sum() for a, b in zip(offs[:-1], offs[1:])])
    torch.testing.assert_close(s, want)
    g = torch.randn(4, dtype=torch.float64)
    (ga,) = torch.autograd.grad(s, p, g)
    (gb,) = torch.autograd.grad(want, p, g)
    torch.testing.assert_close(ga, gb)
    assert torch.autograd.gradcheck(lambda t: _LevelSums.apply(t, offs), (p,))
    # a table whose levels do not start at row 0 / end at the last row
    offs2 = [3, 10, 30]
    (gc,) = torch.autograd.grad(_LevelSums.apply(p, offs2), p, torch.tensor([2.0, -1.0], dtype=torch.float64))
    assert (gc[:3] == 0).all() and (gc[30:] == 0).all() and (gc[3:10] == 2).all() and (gc[10:30] == -1).all()


def test_level_constants_are_cached_per_layout():
    a = _level_const([0, 4, 9], 8, torch.device("cpu"))
    b = _level_const([0, 4, 9], 8, torch.device("cpu"))
    assert a[0] is b[0] and a[1] is b[1]
    assert a[0].tolist() == [0] * 4 + [1] * 5 and a[1].tolist() == [32.0, 40.0]
    c = _level_const([0, 4, 9], 2, torch.device("cpu"))
    assert c[1].tolist() == [8.0, 10.0]


def test_level_entropies_equal_the_per_level_function():
    """same Pg and bits as get_BiRF_wentropy_leveln level by level, same gradient of their sum w.r.t. the table"""
    torch.manual_seed(1)
    offs = [0, 64, 200, 1000]
    table = torch.where(torch.rand(1000, 8) < 0.7, 1.0, -1.0).requires_grad_(True)
    me = _stub(offs)
    Pgs, bits = CNC_context_models.level_entropies(me, table)
    tot_a = sum(bits)
    tot_b = 0
    for n in range(3):
        Pg_n, bit_n, ttl = CNC_context_models.get_BiRF_wentropy_leveln(me, table, n)
        assert ttl == (offs[n + 1] - offs[n]) * 8
        torch.testing.assert_close(Pgs[n], Pg_n, rtol=1e-6, atol=0)
        torch.testing.assert_close(bits[n], bit_n, rtol=1e-5, atol=0)
        tot_b = tot_b + bit_n
    (ga,) = torch.autograd.grad(tot_a, table)
    (gb,) = torch.autograd.grad(tot_b, table)
    torch.testing.assert_close(ga, gb, rtol=1e-5, atol=1e-7)


def test_fused_context_mlp_with_no_voxels():
    """no voxel of the sampled entries touches the occupancy: empty output, zero weight gradients, no kernel call"""
    net = torch.nn.Sequential(torch.nn.Linear(25, 32), torch.nn.LeakyReLU(), torch.nn.Linear(32, 32), torch.nn.LeakyReLU(), torch.nn.Linear(32, 8))
    ps = [net[0].weight, net[0].bias, net[2].weight, net[2].bias, net[4].weight, net[4].bias]
    x = torch.zeros(0, 25, requires_grad=True)
    y = _CtxMLP3.apply(x, *ps)
    assert y.shape == (0, 8)
    grads = torch.autograd.grad(y.sum(), [x] + ps)
    assert grads[0].shape == (0, 25)
    for g, p_ in zip(grads[1:], ps):
        assert g.shape == p_.shape and not g.any()


def test_rows_gather_and_bits_sum_host_paths():
    """the torch branches of `_RowsGather` (distinct rows: scatter backward) and `CNC_context_models._bits_sum` (CPU tensors
    take the op-by-op Bernoulli_entropy) -- the CUDA branches are checked against these expressions in tests/test_gpu_codec.py"""
    from cnc_b200.context_models import Bernoulli_entropy, _RowsGather

    torch.manual_seed(1)
    table = torch.randn(50, 8, dtype=torch.float64, requires_grad=True)
    rows = torch.randperm(50)[:17]
    out = _RowsGather.apply(table, rows)
    assert torch.equal(out, table[rows])
    w = torch.randn(17, 8, dtype=torch.float64)
    (ga,) = torch.autograd.grad((out * w).sum(), table)
    (gb,) = torch.autograd.grad((table[rows] * w).sum(), table)
    assert torch.equal(ga, gb)
    stub = types.SimpleNamespace(entropy_model=Bernoulli_entropy())
    x = torch.where(torch.rand(9, 8) < 0.5, -1.0, 1.0)
    p = torch.rand(9, 8)
    got = CNC_context_models._bits_sum(stub, x, p)
    torch.testing.assert_close(got, torch.sum(Bernoulli_entropy()(x, p)))


def test_level_sums_binary_flag_is_ignored_off_the_gpu():
    offs = [0, 8, 24]
    q = torch.where(torch.rand(24, 8) < 0.4, -1.0, 1.0)
    assert torch.equal(_LevelSums.apply(q, offs, True), _LevelSums.apply(q, offs, False))
