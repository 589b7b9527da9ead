"""GPU suite: context model + range coder end to end (cnc_b200.context_models.CNC_context_models).

  * encode -> decode round trip restores every coded table row, bit for bit (utils_bpp_acc.py:709-999);
  * the fused chunk kernel (cnc_context3d_probs) agrees with the reference's op-by-op data flow on the
    drop-in kernels (K6 mask, K1 masked gather, nn.Linear, K8 pack) to 1e-5, and the int16 CDF entries
    it produces differ on a reported, tiny fraction only;
  * every emitted stream equals the CPU oracle coder's bytes for the same (cdf, symbol) input.
"""
import numpy as np
import pytest
import torch

from conftest import R2, R3

pytestmark = pytest.mark.gpu


def ball(Rb, radius=0.8):
    c = (np.arange(Rb) + 0.5) / Rb * 2 - 1
    X, Y, Z = np.meshgrid(c, c, c, indexing="ij")
    return torch.from_numpy(X * X + Y * Y + Z * Z <= radius * radius)


def make(dev, res3, log2T, res2, log2T2, Rb, skip3=(0, 1, 2), seed=0, fused=True, smooth=True, tables="full"):
    from cnc_b200.context_models import CNC_context_models
    from cnc_b200.gridencoder import GridEncoder

    torch.manual_seed(seed)
    encs = [GridEncoder(num_dim=3, n_features=8, resolutions_list=res3, log2_hashmap_size=log2T, ste_binary=True).to(dev)] + \
           [GridEncoder(num_dim=2, n_features=8, resolutions_list=res2, log2_hashmap_size=log2T2, ste_binary=True).to(dev) for _ in range(3)]
    with torch.no_grad():
        for e in encs:
            # a biased sign field (P(+1) ~ 0.7) so that the context model has something to predict
            e.params.copy_(torch.where(torch.rand_like(e.params) < 0.7, 0.5, -0.5))
    cm = CNC_context_models(num_dim=3, resolutions_list=res3, resolutions_list_2D=res2, log2_hashmap_size=log2T,
                            log2_hashmap_size_2D=log2T2, n_features=8, sample_num=4000, max_context_layer_num=3,
                            ste_binary=True, Rb=Rb, skip_levels_3D=skip3, skip_levels_2D=(0,), device=dev, fused=fused, tables=tables)
    with torch.no_grad():   # context models that give non-trivial, valid probabilities
        for m in list(cm.context_model_3D) + [l for s in cm.context_model_2D for l in s]:
            if isinstance(m, torch.nn.Linear):
                m.weight.mul_(0.5)
        cm.context_model_3D[4].bias.fill_(0.6)
        for s in cm.context_model_2D:
            s[0].bias.fill_(0.6)
    vxl = ball(Rb).to(dev).unsqueeze(0)
    return cm, encs, vxl


SMALL = dict(res3=[6, 10, 18, 34, 66], log2T=12, res2=[18, 34, 66], log2T2=10, Rb=16)


def roundtrip(cm, encs, vxl, oracle=None):
    dev = vxl.device
    Pgs, est_MB, coded_MB, streams = cm.encode_binary_vxl_mixPg_3D2D(*encs, vxl, "t", return_streams=True)
    recs = [torch.ones_like(e.params) for e in encs]
    out = cm.decode_binary_vxl_mixPg_3D2D(*encs, *recs, vxl, Pgs, "t", streams=streams)
    return Pgs, est_MB, coded_MB, streams, out


def test_codec_roundtrip_small(cuda, oracle):
    cm, encs, vxl = make(cuda, **SMALL)
    Pgs, est_MB, coded_MB, streams, rec = roundtrip(cm, encs, vxl)
    assert len(streams) == 3 * 3 + 5 and all(k.endswith(".b") for k in streams)
    q = [torch.where(e.params >= 0, 1.0, -1.0) for e in encs]
    # 3D: skip levels are coded completely; context levels only rows with an occupied voxel, the rest stay +1
    offs = cm.offs
    for n in range(cm.n_levels):
        a, b = q[0][offs[n]:offs[n + 1]], rec[0][offs[n]:offs[n + 1]]
        if n in cm.skip_levels_3D:
            assert torch.equal(a, b)
        else:
            differs = (a != b).any(-1)
            assert (b[differs] == 1).all()           # untouched rows
            assert differs.float().mean() < 0.9       # (with 4096-row tables nearly every row has a voxel in the ball)
    assert abs(est_MB - coded_MB) / coded_MB < 0.05       # the entropy estimate matches what the coder emits
    print(f"small config: estimated {est_MB * 1024:.2f} KiB, coded {coded_MB * 1024:.2f} KiB in {len(streams)} streams")


def test_fused_context_kernel_vs_reference_flow(cuda, oracle):
    cm, encs, vxl = make(cuda, **SMALL)
    pq = cm.get_STE_params(encs[0]).detach()
    for n in (3, 4):
        Pg_n, _, _ = cm.get_BiRF_wentropy_leveln(pq, n)
        E = int(cm.hashparams_num_levels[n])
        pf, ef = cm._probs_3D_fused(encs[0], pq, vxl, n, 0, E, Pg_n)
        pu, eu = cm._probs_3D_unfused(encs[0], pq, vxl, n, 0, E, Pg_n)
        assert torch.equal(ef, eu)                           # mask_exist: integer side, bit exact
        torch.testing.assert_close(pf, pu, rtol=1e-5, atol=1e-6)
        from cnc_b200 import torchac as tac
        c_f, c_u = tac.cdf_from_p(pf), tac.cdf_from_p(pu)
        frac = (c_f != c_u).float().mean().item()
        print(f"level {n}: {ef.sum().item()} coded rows, {frac * 100:.3f}% of the int16 CDF entries differ fused vs op-by-op")
        assert frac < 0.02
        # sub-range call == slice of the full call: bit for bit when the cut falls on the kernel's batch grid (64
        # entries, absolute), to rounding of the overlap-weighted sum otherwise (the chunking of a level is part of
        # the stream layout and identical on both sides of the codec, utils_bpp_acc.py:798-802 == :929-933)
        for lo, hi, exact in ((E // 3 // 64 * 64, 2 * E // 3 // 64 * 64, True), (E // 3 + 1, 2 * E // 3 + 5, False)):
            ps, es = cm._probs_3D_fused(encs[0], pq, vxl, n, lo, hi, Pg_n)
            assert torch.equal(es, ef[lo:hi])
            first = int(ef[:lo].sum())
            want = pf[first:first + int(es.sum())]
            if exact:
                assert torch.equal(ps, want)
            else:
                torch.testing.assert_close(ps, want, rtol=2e-6, atol=1e-7)
        # and the kernel is deterministic: same call, same bits
        p2, _ = cm._probs_3D_fused(encs[0], pq, vxl, n, 0, E, Pg_n)
        assert torch.equal(p2, pf)


def test_streams_equal_oracle_coder(cuda, oracle):
    cm, encs, vxl = make(cuda, **SMALL)
    from cnc_b200 import torchac as tac

    captured = {"c1": [], "sym": []}
    orig = tac.encode_streams_async

    def spy(c1s, syms):
        captured["c1"] += list(c1s)
        captured["sym"] += list(syms)
        return orig(c1s, syms)

    tac.encode_streams_async = spy
    try:
        _, _, _, streams = cm.encode_binary_vxl_mixPg_3D2D(*encs, vxl, "t", return_streams=True)
    finally:
        tac.encode_streams_async = orig
    assert len(captured["c1"]) == len(streams) > 0
    for c1, sym, data in zip(captured["c1"], captured["sym"], streams.values()):   # dicts keep the emission order
        want = oracle.ac_encode(c1.cpu().numpy().view(np.uint16), sym.cpu().numpy())
        assert data == want
        assert np.array_equal(oracle.ac_decode(c1.cpu().numpy().view(np.uint16), data), sym.cpu().numpy())


def test_training_loss_runs_and_backprops(cuda):
    cm, encs, vxl = make(cuda, **SMALL)
    bpp, MB = cm.forward_binary_vxl_mixPg_3D2D(*encs, vxl, step=0)
    assert torch.isfinite(bpp) and 0.3 < float(bpp) < 1.2
    bpp.backward()
    assert encs[0].params.grad is not None and torch.isfinite(encs[0].params.grad).all()
    assert cm.context_model_3D[0].weight.grad.abs().sum() > 0
    assert any(p.grad is not None and p.grad.abs().sum() > 0 for p in cm.context_model_2D.parameters())
    # estimate and codec agree on the order of magnitude of the rate
    _, est_MB, coded_MB, _ = cm.encode_binary_vxl_mixPg_3D2D(*encs, vxl, "t", return_streams=True)
    assert abs(MB - est_MB) / est_MB < 0.25


@pytest.mark.timeout(900)
def test_codec_roundtrip_product_layout(cuda):
    """BASELINE configs[2]: L=12 (res 18..514, T=2^19) + 3 planes x 4 levels (T=2^17), F=8: 33 streams,
    decode(encode(table)) == table on every coded row."""
    cm, encs, vxl = make(cuda, res3=R3, log2T=19, res2=R2, log2T2=17, Rb=128, seed=1)
    Pgs, est_MB, coded_MB, streams, rec = roundtrip(cm, encs, vxl)
    assert len(streams) == 33
    assert sorted(k for k in streams if "3D11" in k) == [f"t_3D11_{i}.b" for i in range(7)]
    q = [torch.where(e.params >= 0, 1.0, -1.0) for e in encs]
    for k in range(4):
        differs = (q[k] != rec[k]).any(-1)
        assert (rec[k][differs] == 1).all()
    for n in cm.skip_levels_3D:
        assert torch.equal(q[0][cm.offs[n]:cm.offs[n + 1]], rec[0][cm.offs[n]:cm.offs[n + 1]])
    n_sym = sum(8 * int(cm.hashparams_num_levels[n]) for n in range(12))
    print(f"product layout: {coded_MB:.3f} MiB coded (estimate {est_MB:.3f}) for {n_sym} 3D symbols max")
    assert abs(est_MB - coded_MB) / coded_MB < 0.02


def test_container_roundtrip_feeds_the_decoder(cuda):
    """encode -> one blob (streams + Pgs + occupancy bits + 13-bit context-model weights + layout incl. the symbol-order
    seed) -> a decoder built from NOTHING but the blob (fresh CNC_context_models, fresh GridEncoders, a global generator
    in another state: ADVICE r1) -> the coded tables"""
    from cnc_b200 import container as C
    from cnc_b200.context_models import CNC_context_models
    from cnc_b200.gridencoder import GridEncoder

    # level 3 is dense AND context-coded: its symbol order is the random permutation
    cm, encs, vxl = make(cuda, res3=[6, 8, 10, 12, 18, 34], log2T=12, res2=[18, 34, 66], log2T2=10, Rb=16)
    assert cm.res[3] <= cm.resolution_thresh and 3 not in cm.skip_levels_3D
    qs = C.quantize_state(cm.state_dict(), digits=13)
    cm.load_state_dict(C.dequantize_state(qs, device=cuda))      # the encoder runs with the weights the decoder will have
    Pgs, _, coded_MB, streams = cm.encode_binary_vxl_mixPg_3D2D(*encs, vxl, "c", return_streams=True)
    blob = C.pack(streams, Pgs, vxl, qs, cm.layout())
    assert len(blob) < coded_MB * 1024 * 1024 + vxl.numel() / 8 + 16384
    q = [torch.where(e.params >= 0, 1.0, -1.0) for e in encs]
    offs, skip = cm.offs, cm.skip_levels_3D
    del cm, encs
    torch.manual_seed(987654)
    torch.rand(11)
    got = C.unpack(blob, device=cuda)
    lay = got["layout"]
    cm2 = CNC_context_models.from_layout(lay, device=cuda)
    cm2.load_state_dict(got["mlp_state"])
    encs2 = [GridEncoder(num_dim=3, n_features=8, resolutions_list=lay["resolutions_list"], log2_hashmap_size=lay["log2_hashmap_size"],
                         ste_binary=True).to(cuda)] + \
            [GridEncoder(num_dim=2, n_features=8, resolutions_list=lay["resolutions_list_2D"], log2_hashmap_size=lay["log2_hashmap_size_2D"],
                         ste_binary=True).to(cuda) for _ in range(3)]
    recs = [torch.ones_like(e.params) for e in encs2]
    out = cm2.decode_binary_vxl_mixPg_3D2D(*encs2, *recs, got["binary_vxl"], got["Pgs_dict"], "c", streams=got["streams"])
    n_coded = 0
    for k in range(4):
        differs = (q[k] != out[k]).any(-1)
        assert (out[k][differs] == 1).all()                       # rows that were never coded stay +1, everything else is back
        n_coded += int((~differs).sum())
    for n in skip:
        assert torch.equal(q[0][offs[n]:offs[n + 1]], out[0][offs[n]:offs[n + 1]])
    # the dense context-coded level: a wrong permutation would scatter its symbols to the wrong rows
    n = 3
    lvl_q, lvl_o = q[0][offs[n]:offs[n + 1]], out[0][offs[n]:offs[n + 1]]
    coded = ~(lvl_q != lvl_o).any(-1)
    assert float(coded.float().mean()) > 0.3 and float((lvl_q[coded] == -1).float().mean()) > 0.2


def test_fused_vote_planes_equal_the_list_based_kernels(cuda):
    """cnc_vote3_fwd / cnc_vote3_bwd (occupancy closed form, no voxel list, no atomics) against get_idx_coords2 +
    3 x cnt_np_embed(_backward): vote counts are exact integers -> identical planes; the gradient w.r.t. the finest
    level agrees to fp32 summation order."""
    cm, encs, vxl = make(cuda, **SMALL)
    assert cm._vote3_ready()
    finest = cm.get_STE_params(encs[0]).detach()[cm.offs[-2]:cm.offs[-1]].clone().requires_grad_(True)
    idx = cm.get_idx_coords2(vxl)
    old = {a: cm.get_pn_embed_frac(finest, idx, axis=a) for a in ("xy", "xz", "yz")}
    g = torch.Generator(device="cpu").manual_seed(4)
    w = {a: torch.randn(old[a].shape, generator=g).to(cuda) for a in old}
    sum(((old[a] * w[a]).sum() for a in old)).backward()
    g_old = finest.grad.clone()
    finest.grad = None
    new = cm.get_pn_embed_frac3(finest, vxl)
    for a in old:
        assert torch.equal(new[a], old[a]), a
    sum(((new[a] * w[a]).sum() for a in new)).backward()
    g_new = finest.grad
    assert torch.equal(g_new != 0, g_old != 0)
    torch.testing.assert_close(g_new, g_old, rtol=1e-4, atol=1e-6 * float(g_old.abs().max()))
    assert float(g_old.abs().max()) > 0


def test_segment_sum_matches_padded_flow(cuda):
    """segment_sum (cnc_segment_wsum_idx / _bwd) == index_select -> align_and_pack -> * weights -> sum(dim=1) of the
    reference's rate term (utils_bpp_acc.py:563-566,741-745), values and gradients"""
    from cnc_b200.context_models import align_and_pack, segment_sum

    g = torch.Generator(device="cpu").manual_seed(3)
    cnt = torch.randint(0, 7, (500,), generator=g).to(cuda)
    cnt[::50] = 0
    M = int(cnt.sum())
    cs = torch.cat([torch.zeros(1, dtype=torch.int64, device=cuda), torch.cumsum(cnt, 0)])
    perm = torch.randperm(M + 13, generator=g)[:M].to(cuda)          # rows of feat that are used, in segment order
    w = torch.rand(M, generator=g).to(cuda)
    for use_idx, use_w in ((True, True), (True, False), (False, True)):
        feat = torch.randn(M + 13, 8, generator=g).to(cuda).requires_grad_(True)
        gout = torch.randn(500, 8, generator=g).to(cuda)
        got = segment_sum.apply(feat, cs, w if use_w else None, perm if use_idx else None)
        got.backward(gout)
        g_got = feat.grad.clone()
        feat.grad = None
        rows = torch.index_select(feat, 0, perm) if use_idx else feat[:M]
        if use_w:
            rows = rows * w[:, None]
        want = torch.sum(align_and_pack.apply(rows, cnt, 0.0, 2), dim=1)
        want.backward(gout)
        torch.testing.assert_close(got, want, rtol=1e-6, atol=1e-6)
        torch.testing.assert_close(g_got, feat.grad, rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("M", [1, 255, 256, 257, 40013])
def test_fused_context_mlp_matches_fp64_autograd(cuda, M):
    """cnc_ctx_mlp_fwd / _bwd (context_model_3D as one forward and one backward kernel) against fp64 autograd of the same
    nn.Sequential: outputs, input gradient and all weight / bias gradients to fp32 rounding (torch's fp32 path shown for scale)"""
    from cnc_b200.context_models import _CtxMLP3

    torch.manual_seed(M)
    def mk():
        return torch.nn.Sequential(torch.nn.Linear(25, 32), torch.nn.LeakyReLU(), torch.nn.Linear(32, 32), torch.nn.LeakyReLU(),
                                   torch.nn.Linear(32, 8)).to(cuda)
    net, net64 = mk(), mk().double()
    net64.load_state_dict({k: v.double() for k, v in net.state_dict().items()})
    x = torch.randn(M, 25, device=cuda, requires_grad=True)
    gy = torch.randn(M, 8, device=cuda)
    ps = [net[0].weight, net[0].bias, net[2].weight, net[2].bias, net[4].weight, net[4].bias]
    ps64 = [net64[0].weight, net64[0].bias, net64[2].weight, net64[2].bias, net64[4].weight, net64[4].bias]
    y = _CtxMLP3.apply(x, *ps)
    got = torch.autograd.grad(y, [x] + ps, gy)
    x64 = x.detach().double().requires_grad_(True)
    y64 = net64(x64)
    want = torch.autograd.grad(y64, [x64] + ps64, gy.double())
    rel = lambda a, b: ((a.double() - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()
    assert y.shape == (M, 8) and all(a.shape == b.shape for a, b in zip(got, want))
    assert rel(y, y64) < 5e-6
    for a, b in zip(got, want):
        assert rel(a, b) < 5e-6


def test_rate_term_same_with_and_without_the_fused_context_mlp(cuda):
    """forward_binary_vxl_mixPg_3D2D: the fused context MLP changes neither the loss nor the gradients (fp32 noise)"""
    cm, encs, vxl = make(cuda, **SMALL)
    params = [e.params for e in encs] + list(cm.parameters())
    out = {}
    for fused in (False, True):
        cm.fused_mlp_train = fused
        torch.manual_seed(3)
        for p in params:
            p.grad = None
        bpp, _ = cm.forward_binary_vxl_mixPg_3D2D(*encs, vxl, step=1)
        bpp.backward()
        out[fused] = (bpp.detach().clone(), [None if p.grad is None else p.grad.clone() for p in params])
    torch.testing.assert_close(out[True][0], out[False][0], rtol=1e-5, atol=0)
    for a, b in zip(out[True][1], out[False][1]):
        assert (a is None) == (b is None)
        if a is not None:
            assert ((a - b).norm() / b.norm().clamp_min(1e-30)).item() < 1e-4


# ------------------------------------------------------------------------------------------------ SURVEY 8f.4
DENSE3 = dict(res3=[6, 8, 10, 12, 18, 34], log2T=12, res2=[18, 34, 66], log2T2=10, Rb=16)   # level 3 dense + context-coded


@pytest.mark.parametrize("cfg", [SMALL, DENSE3, dict(res3=R3, log2T=19, res2=R2, log2T2=17, Rb=128)], ids=["small", "dense3", "product"])
def test_pruned_tables_are_the_masked_full_tables(cuda, cfg):
    """tables='pruned' (histogram pass + occupancy-pruned key sort, csrc/table_build.cu) against the reference construction
    (meshgrid -> hash -> sort -> unique, utils_bpp_acc.py:294-335) followed by the reference's mask (K6, :811-833):
    identical entry numbering, per-entry vertex lists in the same order, identical existing-entry set -- all integers."""
    cm_f, _, vxl = make(cuda, **cfg, seed=3, tables="full")
    cm_p, _, _ = make(cuda, **cfg, seed=3, tables="pruned")
    assert all(p is None for p in cm_p.pos_grid_sorted_list)
    for n in range(cm_f.n_levels):
        assert torch.equal(cm_f.unique_value_list[n], cm_p.unique_value_list[n]), n       # incl. the dense levels' shuffle
        assert torch.equal(cm_f.unique_count_list[n], cm_p.unique_count_list[n]), n
        assert torch.equal(cm_f.unique_count_cumsum_list[n], cm_p.unique_count_cumsum_list[n]), n
    assert torch.equal(cm_f.hashparams_num_levels, cm_p.hashparams_num_levels)
    assert cm_f.utils_points_per_param_levels == cm_p.utils_points_per_param_levels
    vx = vxl.squeeze(0).contiguous()
    total_f = total_p = 0
    for n in range(3, cm_f.n_levels):
        pts_f = cm_f.pos_grid_sorted_list[n]
        mask = cm_f.query_binary_vxl(pts_f, vxl, n)
        E = cm_f.unique_value_list[n].numel()
        entry_f = torch.repeat_interleave(torch.arange(E, device=cuda), cm_f.unique_count_list[n])[mask]
        pts_p, ent, seg, ent_h, seg_h = cm_p._pruned_level(n, vx)
        assert torch.equal(pts_p, pts_f[mask]), n
        e_want, c_want = torch.unique_consecutive(entry_f, return_counts=True)
        assert torch.equal(ent, e_want) and torch.equal(seg[1:] - seg[:-1], c_want), n
        assert np.array_equal(ent_h, ent.cpu().numpy()) and int(seg_h[-1]) == pts_p.shape[0]
        total_f += pts_f.numel() * 2
        total_p += pts_p.numel() * 2
    print(f"vertex lists of the coded levels: full {total_f / 1e6:.1f} MB, pruned {total_p / 1e6:.1f} MB; "
          f"all table state: full {cm_f.table_bytes() / 1e6:.1f} MB, pruned {cm_p.table_bytes() / 1e6:.1f} MB")
    assert total_p < 0.6 * total_f
    # the training loss still works from a pruned object: it builds the full lists on first use, identical to the reference's
    cm_p._ensure_full_tables()
    for n in range(cm_f.n_levels):
        assert torch.equal(cm_f.pos_grid_sorted_list[n], cm_p.pos_grid_sorted_list[n]), n


@pytest.mark.parametrize("cfg", [SMALL, DENSE3], ids=["small", "dense3"])
def test_codec_pruned_mode_roundtrip_and_vs_full(cuda, cfg):
    """encode / decode with pruned tables: round trip restores every coded row; against full-table mode the same files
    carry the same symbols in the same order, the int16 CDFs are at most one step apart (same vertices, same order per
    entry, different batching inside the kernel), and a decoder rebuilt from `layout()` is in pruned mode too."""
    from cnc_b200 import torchac as tac
    from cnc_b200.context_models import CNC_context_models

    def encode(cm, encs, vxl):
        cap, orig = {"c1": [], "sym": []}, tac.encode_streams_async

        def spy(c1s, syms):
            cap["c1"] += list(c1s)
            cap["sym"] += list(syms)
            return orig(c1s, syms)

        tac.encode_streams_async = spy
        try:
            Pgs, est, coded, streams = cm.encode_binary_vxl_mixPg_3D2D(*encs, vxl, "p", return_streams=True)
        finally:
            tac.encode_streams_async = orig
        return Pgs, streams, {k: (c, s_) for k, c, s_ in zip(streams, cap["c1"], cap["sym"])}, coded

    cm_f, encs, vxl = make(cuda, **cfg, seed=5, tables="full")
    cm_p, _, _ = make(cuda, **cfg, seed=5, tables="pruned")
    cm_p.load_state_dict(cm_f.state_dict())
    Pgs_f, st_f, cap_f, coded_f = encode(cm_f, encs, vxl)
    Pgs_p, st_p, cap_p, coded_p = encode(cm_p, encs, vxl)
    assert list(st_f) == list(st_p)
    n_diff = n_tot = 0
    for k in st_f:
        assert torch.equal(cap_f[k][1], cap_p[k][1]), k                       # symbols: same rows, same order
        d = ((cap_f[k][0].to(torch.int32) & 0xFFFF) - (cap_p[k][0].to(torch.int32) & 0xFFFF)).abs()
        assert int(d.max()) <= 1, k
        n_diff, n_tot = n_diff + int((d != 0).sum()), n_tot + d.numel()
    print(f"pruned vs full tables: {n_diff} of {n_tot} int16 CDF entries differ; coded {coded_p * 1024:.2f} vs {coded_f * 1024:.2f} KiB")
    assert n_diff < 0.01 * n_tot
    lay = cm_p.layout()
    assert lay["tables"] == "pruned"
    cm_d = CNC_context_models.from_layout(lay, device=cuda)
    assert cm_d.tables == "pruned"
    cm_d.load_state_dict(cm_p.state_dict())
    out = cm_d.decode_binary_vxl_mixPg_3D2D(*encs, *[torch.ones_like(e.params) for e in encs], vxl, Pgs_p, "p", streams=st_p)
    ref = cm_f.decode_binary_vxl_mixPg_3D2D(*encs, *[torch.ones_like(e.params) for e in encs], vxl, Pgs_f, "p", streams=st_f)
    for a, b, e in zip(out, ref, encs):
        assert torch.equal(a, b)                                             # both modes reconstruct the same tables
        q = torch.where(e.params >= 0, 1.0, -1.0)
        assert ((a == q) | (a == 1)).all()
    assert all(p is None for p in cm_d.pos_grid_sorted_list)                 # the decoder never built a full vertex list


def test_fused_context_gather_equals_forward_diff_levels(cuda):
    """rate term: `_Ctx3DGather` (bitmap-masked 3-level gather + Pg column, csrc/context_train.cu) against
    GridEncoder.forward_diff_levels + cat on the drop-in K1/K2 (per-corner occupancy boxes): identical features on the
    +-1 table -> identical loss; table / Pg / MLP gradients to fp32 atomics order."""
    cm, encs, vxl = make(cuda, res3=R3, log2T=19, res2=R2, log2T2=17, Rb=128, seed=2)
    out = {}
    for fast in (True, False):
        cm.fused_gather_train = fast
        for p in [e.params for e in encs] + list(cm.parameters()):
            p.grad = None
        torch.manual_seed(9)
        bpp, _ = cm.forward_binary_vxl_mixPg_3D2D(*encs, vxl, step=0, sample_num=20000)
        bpp.backward()
        out[fast] = (float(bpp), [e.params.grad.clone() for e in encs], [p.grad.clone() for p in cm.parameters()])
    assert abs(out[True][0] - out[False][0]) <= 1e-6 * abs(out[False][0]), (out[True][0], out[False][0])
    for a, b in zip(out[True][1] + out[True][2], out[False][1] + out[False][2]):
        assert float((a - b).abs().max()) <= 1e-5 * float(b.abs().max()) + 1e-12
    assert float(out[True][1][0].abs().max()) > 0


@pytest.mark.parametrize("K,N", [(17, 1000), (25, 70001), (33, 256), (33, 300000), (9, 5)])
def test_lin8_kernels_vs_torch(cuda, K, N):
    """cnc_lin8_fwd / _bwd (plane context models, csrc/context_lin8.cu) against nn.Linear + autograd in fp64"""
    from cnc_b200.context_models import _linear8

    torch.manual_seed(K + N)
    lin = torch.nn.Linear(K, 8).to(cuda)
    x = torch.randn(N, K, device=cuda, requires_grad=True)
    gy = torch.randn(N, 8, device=cuda)
    y = _linear8(lin, x)
    assert "Lin8" in type(y.grad_fn).__name__
    y.backward(gy)
    got = (y.detach(), x.grad.clone(), lin.weight.grad.clone(), lin.bias.grad.clone())
    lin64 = torch.nn.Linear(K, 8).to(cuda).double()
    lin64.load_state_dict({k: v.double() for k, v in lin.state_dict().items()})
    x64 = x.detach().double().requires_grad_(True)
    y64 = lin64(x64)
    y64.backward(gy.double())
    for a, b in zip(got, (y64.detach(), x64.grad, lin64.weight.grad, lin64.bias.grad)):
        assert float((a.double() - b).abs().max()) <= 2e-6 * float(b.abs().max()) * max(1.0, (N / 1e4) ** 0.5)


def test_fused_plane_gather_equals_the_two_k1_flow(cuda):
    """`_Ctx2DGather` (c plane levels | vote fraction plane | Pg in one kernel, csrc/context_train.cu) against
    Encoding_2D(...) + forward_given_params(...) + cat on the drop-in K1/K2: same loss, gradients of the plane tables, of the
    finest 3D level (through the vote planes) and of the context models to fp32 atomics order; same coded streams."""
    cm, encs, vxl = make(cuda, res3=R3, log2T=19, res2=R2, log2T2=17, Rb=128, seed=4)
    with torch.no_grad():   # keep the plane probabilities away from the clamp (see the A/B suite)
        for s_ in cm.context_model_2D:
            s_[0].weight.mul_(0.5)
    out, streams = {}, {}
    for fast in (True, False):
        cm.fused_gather2d = fast
        for p in [e.params for e in encs] + list(cm.parameters()):
            p.grad = None
        torch.manual_seed(9)
        bpp, _ = cm.forward_binary_vxl_mixPg_3D2D(*encs, vxl, step=0, sample_num=20000)
        bpp.backward()
        out[fast] = (float(bpp), [e.params.grad.clone() for e in encs], [p.grad.clone() for p in cm.parameters()])
        streams[fast] = cm.encode_binary_vxl_mixPg_3D2D(*encs, vxl, "g", return_streams=True)[3]
    assert abs(out[True][0] - out[False][0]) <= 1e-6 * abs(out[False][0]), (out[True][0], out[False][0])
    for a, b in zip(out[True][1] + out[True][2], out[False][1] + out[False][2]):
        assert float((a - b).abs().max()) <= 2e-5 * float(b.abs().max()) + 1e-12
    assert all(float(g.abs().max()) > 0 for g in out[True][1])
    assert list(streams[True]) == list(streams[False])
    assert all(streams[True][k] == streams[False][k] for k in streams[True])    # identical features -> identical bytes


def test_rate_term_shared_among_data_parallel_ranks(cuda):
    """CNC_context_models.set_data_parallel: the plane terms dealt over the ranks (times world), the sampled 3D entries split
    among them -- the AVERAGE over the ranks of loss and gradients is the single-process rate term.  Checked with every entry
    sampled (so that the 3D part is exact, too); and with the training sample size, that each rank samples its share and the
    windows of different ranks differ."""
    cm, encs, vxl = make(cuda, **SMALL)
    params = [e.params for e in encs] + [p for p in cm.parameters() if p.requires_grad]

    def run(rank, world, sample_num=None):
        cm.set_data_parallel(rank, world)
        for p in params:
            p.grad = None
        torch.manual_seed(0)
        bpp, _ = cm.forward_binary_vxl_mixPg_3D2D(*encs, vxl, step=0, sample_num=sample_num)
        bpp.backward()
        return float(bpp), [torch.zeros_like(p) if p.grad is None else p.grad.clone() for p in params]

    everything = 10 ** 9
    full_bpp, full_g = run(0, 1, sample_num=everything)
    cm.sample_num = everything * 3          # (a third of it is still every entry)
    world = 3
    parts = [run(r, world) for r in range(world)]
    np.testing.assert_allclose(np.mean([b for b, _ in parts]), full_bpp, rtol=2e-5)
    assert max(abs(b - full_bpp) for b, _ in parts) > 1e-3 * full_bpp          # no rank computes the whole thing
    for k, g in enumerate(full_g):
        avg = sum(p[1][k] for p in parts) / world
        torch.testing.assert_close(avg, g, rtol=2e-4, atol=2e-6 * float(g.abs().max()) + 1e-12)
    # training sample size: a rank's share, windows shifted by rank / world
    cm.sample_num = 4000
    cm.set_data_parallel(1, 4)
    snl, n_valid = cm._dp_snl
    assert abs(int(snl.sum()) - 1000) <= cm.n_levels and 0 < n_valid <= int(snl.sum())
    b1, _ = run(1, 4)
    b2, _ = run(2, 4)
    assert np.isfinite([b1, b2]).all() and b1 != b2
    cm.set_data_parallel(0, 1)


def test_rate_term_fused_pieces_match_the_op_by_op_expressions(cuda):
    """the small-op chains of the rate term as single kernels: sum of Bernoulli_entropy (value and both gradients), gather of
    distinct rows with a scatter backward, level sums of a +-1 table from its sign plane"""
    from cnc_b200.context_models import Bernoulli_entropy, _BernoulliBitsSum, _LevelSums, _RowsGather

    g = torch.Generator().manual_seed(2)
    n = 70001
    x = torch.where(torch.rand(n, 8, generator=g) < 0.5, -1.0, 1.0).to(cuda).requires_grad_(True)
    p = torch.rand(n, 8, generator=g).to(cuda)
    with torch.no_grad():
        p[:50] = torch.tensor([0.0, 1.0, 1e-7, 1 - 1e-7, 1e-6, 1 - 1e-6, 0.5, 2.0], device=cuda)      # on and beyond the clamp
    p.requires_grad_(True)
    want = torch.sum(Bernoulli_entropy()(x, p)) * 0.37
    want.backward()
    gx, gp = x.grad.clone(), p.grad.clone()
    x.grad = p.grad = None
    got = _BernoulliBitsSum.apply(x, p) * 0.37
    got.backward()
    torch.testing.assert_close(got, want, rtol=2e-6, atol=0)
    torch.testing.assert_close(x.grad, gx, rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(p.grad, gp, rtol=1e-5, atol=1e-7)
    # rows gather
    table = torch.randn(5000, 8, generator=g).to(cuda).requires_grad_(True)
    rows = torch.randperm(5000, generator=g)[:1234].to(cuda)
    w = torch.randn(1234, 8, generator=g).to(cuda)
    (table[rows] * w).sum().backward()
    g0 = table.grad.clone()
    table.grad = None
    out = _RowsGather.apply(table, rows)
    assert torch.equal(out, table[rows])
    (out * w).sum().backward()
    assert torch.equal(table.grad, g0)
    # level sums
    offs = [0, 8, 1000, 1008, 4000]
    q = torch.where(torch.rand(4000, 8, generator=g) < 0.3, -1.0, 1.0).to(cuda).requires_grad_(True)
    a = _LevelSums.apply(q, offs, True)
    b = _LevelSums.apply(q, offs, False)
    assert torch.equal(a, b)
    (a * torch.arange(1, 5, device=cuda)).sum().backward()
    assert torch.equal(q.grad[:8], torch.ones(8, 8, device=cuda)) and torch.equal(q.grad[1008:], torch.full((2992, 8), 4.0, device=cuda))
