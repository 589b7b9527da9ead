"""GPU suite: occupancy-grid marching, packed scans and the volume-rendering tail (cnc_b200.nerfacc) against
the CPU oracle (oracle/cnc_oracle_march.c, pinned by the nerfacc docstring KATs in tests/golden) and, where the
reference extension is present, against the reference's own nerfacc binary.  Sample layout (counts, ray indices)
and interval edges are compared exactly; transmittance / weights to 1e-5 (the reference's scan tree order differs)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def T(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def scene(n_rays=3000, Rb=64, seed=0, levels=1):
    rng = np.random.default_rng(seed)
    # camera positions on a radius-4 sphere looking roughly at the origin
    o = rng.normal(size=(n_rays, 3))
    o = (o / np.linalg.norm(o, axis=1, keepdims=True) * 4).astype(np.float32)
    tgt = rng.uniform(-0.8, 0.8, (n_rays, 3))
    d = tgt - o
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    d[0] = [0, 0, 1]; o[0] = [0.1, 0.2, -4]       # axis-aligned ray (zero direction components)
    d[1] = [1, 0, 0]; o[1] = [9, 9, 9]            # misses everything
    c = (np.arange(Rb) + 0.5) / Rb * 3 - 1.5
    X, Y, Z = np.meshgrid(c, c, c, indexing="ij")
    bins = np.stack([(X * X + Y * Y + Z * Z <= 1.0)] + [np.ones_like(X, bool)] * (levels - 1)).astype(np.uint8)
    aabbs = np.stack([np.array([-1.5, -1.5, -1.5, 1.5, 1.5, 1.5], np.float32) * 2 ** l for l in range(levels)])
    return o, d, bins, aabbs


def test_ray_aabb_and_kats(cuda, oracle, golden):
    from cnc_b200 import nerfacc as N

    o, d, bins, aabbs = scene(levels=2)
    tmin, tmax, hits = N.ray_aabb_intersect(T(o, cuda), T(d, cuda), T(aabbs, cuda))
    rmin, rmax, rh = oracle.ray_aabb_intersect(o, d, aabbs)
    np.testing.assert_array_equal(hits.cpu().numpy(), rh)
    np.testing.assert_array_equal(tmin.cpu().numpy(), rmin)
    np.testing.assert_array_equal(tmax.cpu().numpy(), rmax)
    # nerfacc docstring known answers (pack.py:29-32, scan.py:36-39,78-81,127-130,170-173)
    ri = T(golden["kat_ray_indices_9"], cuda)
    pk = N.pack_info(ri, 3)
    np.testing.assert_array_equal(pk.cpu().numpy(), golden["kat_packed_info"])
    x = T(golden["kat_scan_in"], cuda)
    for fn, key in ((N.inclusive_sum, "kat_inclusive_sum"), (N.exclusive_sum, "kat_exclusive_sum"),
                    (N.inclusive_prod, "kat_inclusive_prod"), (N.exclusive_prod, "kat_exclusive_prod")):
        np.testing.assert_array_equal(fn(x, pk).cpu().numpy(), golden[key])
    # volrend.py:349-357 / :463-473
    pk7 = N.pack_info(T(golden["kat_ray_indices_7"], cuda), 3)
    w, tr, al = N.render_weight_from_density(T(golden["kat_t_starts"], cuda), T(golden["kat_t_ends"], cuda),
                                             T(golden["kat_sigmas"], cuda), packed_info=pk7)
    np.testing.assert_allclose(w.cpu().numpy(), golden["kat_weights_from_density"], atol=6e-3)
    np.testing.assert_allclose(tr.cpu().numpy(), golden["kat_trans_from_density"], atol=6e-3)
    vis = N.render_visibility_from_density(T(golden["kat_t_starts"], cuda), T(golden["kat_t_ends"], cuda),
                                           T(golden["kat_sigmas"], cuda), packed_info=pk7, early_stop_eps=0.3, alpha_thre=0.2)
    np.testing.assert_array_equal(vis.cpu().numpy().astype(np.uint8), golden["kat_visibility"])


@pytest.mark.parametrize("levels,step,cone,limit", [(1, 5e-3, 0.0, None), (2, 1e-2, 0.0, None), (1, 1e-2, 4e-3, None), (1, 5e-3, 0.0, 17)])
def test_traverse_grids_exact_vs_oracle(cuda, oracle, levels, step, cone, limit):
    from cnc_b200 import nerfacc as N

    o, d, bins, aabbs = scene(levels=levels, seed=levels)
    rng = np.random.default_rng(3)
    near = (rng.random(len(o)) * step).astype(np.float32)           # stratified jitter (occ_grid.py:172-173)
    far = np.full(len(o), 1e10, np.float32)
    mask = None if limit is None else (rng.random(len(o)) < 0.7)
    iv, sm, term = N.traverse_grids(T(o, cuda), T(d, cuda), T(bins.astype(bool), cuda), T(aabbs, cuda), T(near, cuda), T(far, cuda),
                                    step_size=step, cone_angle=cone, traverse_steps_limit=limit,
                                    rays_mask=None if mask is None else T(mask, cuda))
    t0, t1, ri, pk, rterm = oracle.traverse_grids(o, d, bins, aabbs, near, far, step, cone, 0 if limit is None else limit, mask)
    np.testing.assert_array_equal(sm.packed_info.cpu().numpy(), pk)            # sample counts per ray: exact
    np.testing.assert_array_equal(sm.ray_indices.cpu().numpy(), ri)
    np.testing.assert_array_equal(iv.vals[iv.is_left].cpu().numpy(), t0)        # interval edges: exact
    np.testing.assert_array_equal(iv.vals[iv.is_right].cpu().numpy(), t1)
    live = np.ones(len(o), bool) if mask is None else mask
    np.testing.assert_array_equal(term.cpu().numpy()[live], rterm[live])
    assert pk[1, 1] == 0 and pk[:, 1].sum() > 10 * len(o)
    if limit is not None:
        assert pk[:, 1].max() == limit and (pk[~mask, 1] == 0).all()


def test_render_tail_vs_oracle_and_autograd(cuda, oracle):
    from cnc_b200 import nerfacc as N

    o, d, bins, aabbs = scene(n_rays=800, seed=5)
    t0, t1, ri, pk, _ = oracle.traverse_grids(o, d, bins, aabbs, None, None, 1e-2)
    rng = np.random.default_rng(6)
    sig = (rng.random(len(t0)) * 8).astype(np.float32)
    rgb = rng.random((len(t0), 3)).astype(np.float32)
    ref = oracle.render_from_density(t0, t1, sig, pk, rgb)
    tt0, tt1, tsig, trgb, tri, tpk = (T(x, cuda) for x in (t0, t1, sig, rgb, ri, pk))
    w, tr, al = N.render_weight_from_density(tt0, tt1, tsig, ray_indices=tri, n_rays=len(o))
    # (the oracle sums a ray sequentially, the kernel in a fixed 32-wide shuffle tree, the reference in its own smem tree:
    #  1e-5, the contract for fp32 results)
    np.testing.assert_allclose(w.cpu().numpy(), ref["weights"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(tr.cpu().numpy(), ref["trans"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(al.cpu().numpy(), ref["alphas"], rtol=1e-5, atol=1e-7)
    col, op, dep = N.render_fused(tt0, tt1, tsig, trgb, tpk)
    np.testing.assert_allclose(col.cpu().numpy(), ref["colors"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(op[:, 0].cpu().numpy(), ref["opacities"], rtol=1e-5, atol=1e-6)
    # rendering(): patched 3-tuple callback, extras, background
    def rgb_sigma_fn(a, b, r):
        return trgb, tsig, torch.zeros(len(a), 3, device=cuda)
    bk = torch.ones(3, device=cuda)
    c2, o2, d2, ex = N.rendering(tt0, tt1, tri, n_rays=len(o), rgb_sigma_fn=rgb_sigma_fn, render_bkgd=bk)
    cb, ob, db = N.render_fused(tt0, tt1, tsig, trgb, tpk, render_bkgd=bk)
    torch.testing.assert_close(c2, cb, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(o2, ob, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(d2, db, rtol=1e-4, atol=1e-5)
    assert set(ex) >= {"weights", "alphas", "trans", "sigmas", "rgbs", "positions"}
    # autograd: analytic backward of the fused weights kernel == torch composition (double)
    s = tsig.double().requires_grad_(True)
    dt = (tt1 - tt0).double()
    sd = s * dt
    cs = torch.cumsum(sd, 0)
    starts = tpk[:, 0][tri]
    base = torch.where(starts > 0, cs[(starts - 1).clamp_min(0)], torch.zeros_like(cs))
    excl = cs - sd - base
    w_ref = torch.exp(-excl) * (1 - torch.exp(-sd))
    g = torch.randn(len(t0), device=cuda, dtype=torch.float64)
    (w_ref * g).sum().backward()
    s32 = tsig.clone().requires_grad_(True)
    w32, _, _ = N.render_weight_from_density(tt0, tt1, s32, ray_indices=tri, n_rays=len(o))
    (w32 * g.float()).sum().backward()
    torch.testing.assert_close(s32.grad.double(), s.grad, rtol=2e-4, atol=1e-6)


def test_occ_grid_estimator_sampling_and_update(cuda):
    from cnc_b200 import nerfacc as N

    torch.manual_seed(0)
    est = N.OccGridEstimator(roi_aabb=[-1.5, -1.5, -1.5, 1.5, 1.5, 1.5], resolution=64, levels=1).to(cuda)
    assert est.binaries.shape == (1, 64, 64, 64) and est.aabbs.shape == (1, 6) and est.occs.shape == (64 ** 3,)

    def density(x):   # a soft ball of radius 1
        return (20.0 * torch.sigmoid((1.0 - x.norm(dim=-1, keepdim=True)) * 20)).float()

    est.train()
    for step in range(0, 48, 16):
        est.update_every_n_steps(step, occ_eval_fn=lambda x: density(x) * 5e-3, occ_thre=1e-2)
    frac = est.binaries.float().mean().item()
    assert 0.10 < frac < 0.25       # ~ volume of the unit ball in the [-1.5,1.5]^3 box (15.5 %)
    o, d, _, _ = scene(n_rays=2000, seed=9)
    ro, rd = T(o, cuda), T(d, cuda)

    def sigma_fn(t0, t1, ri):
        return density(ro[ri] + rd[ri] * ((t0 + t1) / 2)[:, None]).squeeze(-1)

    ri, t0, t1 = est.sampling(ro, rd, sigma_fn=sigma_fn, render_step_size=5e-3, stratified=True)
    ri_all, t0_all, _ = est.sampling(ro, rd, render_step_size=5e-3)
    assert 0 < ri.numel() < ri_all.numel()           # visibility filtering removed the occluded samples
    assert (ri[1:] >= ri[:-1]).all() and torch.allclose(t1 - t0, torch.full_like(t0, 5e-3), atol=1e-6)
    mid = ro[ri] + rd[ri] * ((t0 + t1) / 2)[:, None]
    assert (mid.norm(dim=-1) < 1.35).all()               # samples only in cells near the (soft) ball, none in empty space
    est.eval()
    with pytest.raises(RuntimeError):
        est.update_every_n_steps(0, occ_eval_fn=density)


class _BallField(torch.nn.Module):
    """analytic radiance field: a soft ball of radius 1, colour from position and direction"""

    def query_density(self, x):
        return (20.0 * torch.sigmoid((1.0 - x.norm(dim=-1, keepdim=True)) * 20)).float()

    def forward(self, x, d):
        return torch.sigmoid(3.0 * x + d), self.query_density(x)


def _ball_estimator(cuda, res=64):
    from cnc_b200 import nerfacc as N

    est = N.OccGridEstimator(roi_aabb=[-1.5, -1.5, -1.5, 1.5, 1.5, 1.5], resolution=res, levels=1).to(cuda)
    c = (torch.arange(res, device=cuda) + 0.5) / res * 3 - 1.5
    X, Y, Z = torch.meshgrid(c, c, c, indexing="ij")
    est.binaries.copy_(((X * X + Y * Y + Z * Z) <= 1.2 ** 2).unsqueeze(0))
    return est.eval()


def test_test_time_renderer_matches_the_full_march(cuda):
    """examples/utils.py:316-489 against :83-216 on the same rays: without early stopping the wavefront visits exactly
    the samples of the one-shot march (the step-limited march resumes on the same t lattice); with it the image moves by
    less than early_stop_eps and fewer samples are evaluated."""
    from cnc_b200.render import Rays, render_image_with_occgrid, render_image_with_occgrid_test

    est, field = _ball_estimator(cuda), _BallField().to(cuda).eval()
    o, d, _, _ = scene(n_rays=1200, seed=4)
    rays = Rays(origins=T(o, cuda), viewdirs=T(d, cuda))
    bkgd = torch.ones(3, device=cuda)
    kw = dict(render_step_size=5e-3, render_bkgd=bkgd)
    with torch.no_grad():
        rgb_f, opa_f, dep_f, n_f = render_image_with_occgrid(field, est, rays, test_chunk_size=500, **kw)
    rgb_w, opa_w, dep_w, n_w = render_image_with_occgrid_test(4096, field, est, rays, early_stop_eps=0.0, **kw)
    assert n_w >= n_f > 0       # the one-shot path drops samples behind T < 1e-4 (visibility pass), the wavefront keeps them
    torch.testing.assert_close(rgb_w, rgb_f, rtol=0, atol=2e-4)
    torch.testing.assert_close(opa_w, opa_f, rtol=0, atol=2e-4)
    hit = opa_f.view(-1) > 0.5
    torch.testing.assert_close(dep_w[hit], dep_f[hit], rtol=0, atol=2e-3)
    rgb_e, opa_e, dep_e, n_e = render_image_with_occgrid_test(4096, field, est, rays, early_stop_eps=1e-4, **kw)
    assert 0 < n_e < n_w
    torch.testing.assert_close(rgb_e, rgb_w, rtol=0, atol=3e-4)
    assert rgb_e.shape == (1200, 3) and opa_e.shape == (1200, 1) and dep_e.shape == (1200, 1)
    assert (opa_e[1] == 0).all() and torch.equal(rgb_e[1], bkgd)          # the ray that misses the box shows the background
    # fixed round size: same image within the early-stop epsilon, at most (round - 1) samples more per retired ray
    rgb_k, opa_k, dep_k, n_k = render_image_with_occgrid_test(4096, field, est, rays, early_stop_eps=1e-4, samples_per_round=32, **kw)
    assert n_e <= n_k <= n_e + 31 * 1200
    torch.testing.assert_close(rgb_k, rgb_e, rtol=0, atol=3e-4)
    torch.testing.assert_close(opa_k, opa_e, rtol=0, atol=3e-4)
    # image-shaped rays and a sample budget that ends the loop early
    img = Rays(origins=rays.origins[:1000].view(25, 40, 3), viewdirs=rays.viewdirs[:1000].view(25, 40, 3))
    rgb_i, opa_i, _, n_i = render_image_with_occgrid_test(8, field, est, img, early_stop_eps=1e-4, **kw)
    assert rgb_i.shape == (25, 40, 3) and opa_i.shape == (25, 40, 1) and 0 < n_i <= 8 * 1000


def test_test_time_renderer_on_the_fused_field(cuda):
    """the product field behind the wavefront renderer: same image as the chunked one-shot renderer"""
    from test_gpu_field import make_field
    from cnc_b200.render import Rays, render_image_with_occgrid, render_image_with_occgrid_test

    f = make_field(cuda).eval()
    est = _ball_estimator(cuda, res=128)
    o, d, _, _ = scene(n_rays=800, seed=5)
    rays = Rays(origins=T(o, cuda), viewdirs=T(d, cuda))
    kw = dict(render_step_size=5e-3, render_bkgd=torch.zeros(3, device=cuda))
    with torch.no_grad():
        rgb_f, opa_f, _, n_f = render_image_with_occgrid(f, est, rays, test_chunk_size=8192, **kw)
    rgb_w, opa_w, _, n_w = render_image_with_occgrid_test(4096, f, est, rays, early_stop_eps=1e-4, **kw)
    assert n_w > 0 and n_f > 0
    # A resumed march recomputes the cell boundaries from its new start, so a sample whose midpoint sits within an ulp of
    # a boundary can fall on the other side (the reference's algorithm has the same property); with this field's small
    # alphas (~2e-3 per sample) one such sample moves a ray by ~1e-3.  Everything else agrees to the early-stop epsilon.
    for w, f_ in ((rgb_w, rgb_f), (opa_w, opa_f)):
        diff = (w - f_).abs().amax(dim=-1)
        assert diff.max().item() < 4e-3
        assert torch.quantile(diff, 0.98).item() < 3e-4


@pytest.mark.parametrize("cone,step,alpha_thre", [(0.0, 5e-3, 0.0), (4e-3, 1e-2, 0.0), (0.0, 5e-3, 1e-3)])
def test_device_loop_equals_the_host_loop(cuda, cone, step, alpha_thre):
    """the sync-free wavefront renderer (`device_loop=True`: wf_begin / wf_march / cnc_field_fwd_n / wf_composite queued in
    batches) against the python loop of examples/utils.py:395-479 on the same kernels: the same rounds take the same
    samples (total_samples equal), images to 1e-5; image-shaped rays, a dense trained-like field (weights scaled up so
    that rays saturate and the early stop and the growing round size are exercised), a sample budget that cuts the loop."""
    from test_gpu_field import make_field
    from cnc_b200.render import Rays, render_image_with_occgrid_test

    f = make_field(cuda).eval()
    with torch.no_grad():
        f.mlp_base.network[2].bias[0] += 4.0          # densities of e^3: opaque after a few dozen samples
    f.invalidate_caches()
    est = _ball_estimator(cuda, res=128)
    o, d, _, _ = scene(n_rays=48 * 40, seed=6)
    rays = Rays(origins=T(o, cuda).view(48, 40, 3), viewdirs=T(d, cuda).view(48, 40, 3))
    kw = dict(render_step_size=step, render_bkgd=torch.ones(3, device=cuda), cone_angle=cone, alpha_thre=alpha_thre)
    for budget in (24, 1024):
        a = render_image_with_occgrid_test(budget, f, est, rays, device_loop=False, **kw)
        b = render_image_with_occgrid_test(budget, f, est, rays, device_loop=True, rounds_per_check=7, **kw)
        c = render_image_with_occgrid_test(budget, f, est, rays, **kw)                 # default: the device loop
        # The host loop accumulates with index_add_ (atomics, unordered), the device loop sequentially per ray: a ray whose
        # opacity lands within rounding of the 1 - 1e-4 threshold may be retired one round apart (a handful of samples in
        # 1.6e5, a change of the image below the early-stop epsilon).  The device loop itself is deterministic.
        assert b[3] == c[3] > 0 and abs(a[3] - b[3]) <= 1e-4 * a[3], (a[3], b[3], c[3])
        for x, y, z in zip(a[:3], b[:3], c[:3]):
            assert x.shape == y.shape
            close = ((y - x).abs() <= 1e-5 * x.abs() + 2e-6).all(-1)
            assert float(close.float().mean()) > 0.998, float(close.float().mean())
            torch.testing.assert_close(y, x, rtol=0, atol=2e-4)
            assert torch.equal(y, z)
    assert float(a[1].mean()) > 0.3
    if cone == 0.0:
        assert float((a[1] > 1 - 1e-4).float().mean()) > 0.05   # rays did saturate


def test_fused_render_backward_sample_points_and_compaction(cuda):
    """the one-kernel pieces of the training step against the torch expressions they replace: the analytic backward of the
    rendering tail (volrend.py:14-160 through autograd), the query points of a sample batch (examples/utils.py:250-262) and
    the boolean-mask compaction of OccGridEstimator.sampling (occ_grid.py:192-197) with its pack_info"""
    from cnc_b200 import nerfacc as N
    from cnc_b200.render import sample_points

    g = torch.Generator().manual_seed(3)
    cnt = torch.randint(0, 90, (700,), generator=g)
    cnt[5] = 0
    cnt[-1] = 0
    n, R = int(cnt.sum()), cnt.numel()
    ri = torch.repeat_interleave(torch.arange(R), cnt).to(cuda)
    t0 = (torch.rand(n, generator=g) * 3).to(cuda)
    t1 = t0 + 5e-3
    sig = (torch.rand(n, generator=g) * 30).to(cuda).requires_grad_(True)
    rgb = torch.rand(n, 3, generator=g).to(cuda).requires_grad_(True)
    pk = N.pack_info(ri, R)
    # --- backward: the fused Function vs the op-by-op composition (scans + index_add), same upstream gradients
    w, T_, al, col, op, dp = N._RenderAll.apply(t0, t1, sig, rgb, pk, ri)
    gC, gO, gD, gW = (torch.randn(R, 3, generator=g).to(cuda), torch.randn(R, generator=g).to(cuda), torch.randn(R, generator=g).to(cuda),
                      torch.randn(n, generator=g).to(cuda))
    (col * gC).sum().add((op * gO).sum()).add((dp * gD).sum()).add((w * gW).sum()).backward()
    gs1, gr1 = sig.grad.clone(), rgb.grad.clone()
    sig.grad = rgb.grad = None
    w2, _, _ = N.render_weight_from_density(t0, t1, sig, ray_indices=ri, n_rays=R)
    col2 = N.accumulate_along_rays(w2, rgb, ri, R)
    op2 = N.accumulate_along_rays(w2, None, ri, R).squeeze(-1)
    dp2 = N.accumulate_along_rays(w2, ((t0 + t1) / 2)[:, None], ri, R).squeeze(-1)
    (col2 * gC).sum().add((op2 * gO).sum()).add((dp2 * gD).sum()).add((w2 * gW).sum()).backward()
    torch.testing.assert_close(gr1, rgb.grad, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(gs1, sig.grad, rtol=2e-4, atol=2e-5 * float(sig.grad.abs().max()))
    # --- query points
    o, d = torch.randn(R, 3, generator=g).to(cuda), torch.randn(R, 3, generator=g).to(cuda)
    pos, dirs = sample_points(o, d, ri, t0, t1)
    assert torch.equal(dirs, d[ri])
    assert torch.equal(pos, o[ri] + d[ri] * (t0 + t1)[:, None] / 2.0)
    # --- compaction
    keep = (torch.rand(n, generator=g) < 0.4).to(cuda)
    ri2, a0, a1 = N._compact(keep, t0, t1, ri, pk)
    assert torch.equal(ri2, ri[keep]) and torch.equal(a0, t0[keep]) and torch.equal(a1, t1[keep])
    assert torch.equal(ri2._cnc_packed, N.pack_info(ri[keep], R))
    assert N._packed(None, ri2, R, None) is ri2._cnc_packed or torch.equal(N._packed(None, ri2, R, None), ri2._cnc_packed)
    none = torch.zeros(n, dtype=torch.bool, device=cuda)
    ri3, b0, b1 = N._compact(none, t0, t1, ri, pk)
    assert ri3.numel() == 0 and torch.equal(ri3._cnc_packed, torch.zeros(R, 2, dtype=torch.int64, device=cuda))
