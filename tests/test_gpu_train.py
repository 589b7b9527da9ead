"""GPU suite: the training-step driver (cnc_b200.trainer.TrainStep) and the table passes under it
(csrc/train_ops.cu through cnc_b200.train_ops).

  * planes / stand-in latents / fused Adam kernels against the torch arithmetic of the same module (bit planes exact,
    Adam to fp32 rounding against torch.optim.Adam), at the product table size and at sizes that are not multiples of
    the kernels' 1024-element chunks;
  * TrainStep on the product field: the photometric loss of a fixed ray batch falls over 50 steps, the sign planes the
    next forward reads are the signs of the updated latents, and lambda > 0 (rate term) steps run and stay finite.
"""
import numpy as np
import pytest
import torch

from conftest import R2, R3

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [32, 992, 1024, 4003896 * 8, 345616 * 8 // 2 + 32])
def test_table_passes_vs_torch(cuda, n):
    from cnc_b200.train_ops import adam_planes, planes_pack, surrogate_fill

    g = torch.Generator().manual_seed(n % 1000)
    small = n <= 1 << 20
    p = (torch.randn(n, generator=g) * 0.9) if small else (torch.randn(n, device=cuda) * 0.9).cpu()
    p[:4] = torch.tensor([0.0, -0.0, 1.0, -1.0])
    pc = p.to(cuda)
    s_c, m_c = planes_pack(pc)
    s_h, m_h = planes_pack(p.clone())
    assert torch.equal(s_c.cpu(), s_h) and torch.equal(m_c.cpu(), m_h)
    lo, hi = (n // 3) // 32 * 32, (2 * n // 3) // 32 * 32
    q_c, q_h = pc.clone(), p.clone()
    surrogate_fill(q_c, s_c, m_c, lo, hi)
    surrogate_fill(q_h, s_h, m_h, lo, hi)
    assert torch.equal(q_c.cpu(), q_h)
    assert torch.equal(q_c[lo:hi], pc[lo:hi])
    # Adam: three steps against torch.optim.Adam (fused) on the same gradients, with a loss scale to divide out
    a = torch.nn.Parameter(pc.clone())
    opt = torch.optim.Adam([a], lr=6e-3, eps=1e-15, weight_decay=2e-6, fused=True)
    b, m1, v2 = pc.clone(), torch.zeros(n, device=cuda), torch.zeros(n, device=cuda)
    sign, mask = torch.empty(n // 8, dtype=torch.uint8, device=cuda), torch.empty(n // 8, dtype=torch.uint8, device=cuda)
    for step in range(1, 4):
        gr = torch.randn(n, device=cuda) * (10.0 ** float(np.random.default_rng(step).integers(-6, 1)))
        gr[::7] = 0                                               # untouched table rows: zero gradient
        a.grad = gr.clone()
        opt.step()
        adam_planes(b, gr * 1024.0, m1, v2, step=step, lr=6e-3, eps=1e-15, weight_decay=2e-6, grad_scale=1024.0, sign=sign, mask=mask)
        torch.testing.assert_close(b, a.detach(), rtol=2e-5, atol=1e-7)
    s2, m2 = planes_pack(b)
    assert torch.equal(sign, s2) and torch.equal(mask, m2)       # the planes the kernel emitted are those of what it wrote
    # STE window folded into the pass: the gradient of latents outside [-1, 1] is dropped (weight decay still acts on them)
    with torch.no_grad():
        a[: n // 2] *= 1.5
        b[: n // 2] *= 1.5
    gr = torch.randn(n, device=cuda)
    a.grad = gr * ((b >= -1) & (b <= 1))     # (the window of the kernel's own copy: the two copies differ by rounding, and |p| = 1 is a step)
    opt.step()
    adam_planes(b, gr, m1, v2, step=4, lr=6e-3, eps=1e-15, weight_decay=2e-6, ste_window=True)
    torch.testing.assert_close(b, a.detach(), rtol=2e-5, atol=1e-7)
    assert float(((a.detach().abs() > 1)).float().mean()) > 0.05


def _scene(dev, n_rays=600, seed=0):
    from cnc_b200.field import NGPRadianceField_mygrid_2D3D
    from cnc_b200.nerfacc import OccGridEstimator
    from cnc_b200.render import Rays

    torch.manual_seed(seed)
    field = NGPRadianceField_mygrid_2D3D(aabb=[-1.5] * 3 + [1.5] * 3, n_features_per_level=8, n_neurons=160, resolutions_list=R3,
                                         log2_hashmap_size=19, resolutions_list_2D=R2, log2_hashmap_size_2D=17, ste_binary=True).to(dev)
    est = OccGridEstimator(roi_aabb=[-1.5] * 3 + [1.5] * 3, resolution=128, levels=1).to(dev)
    c = (torch.arange(128, device=dev) + 0.5) / 128 * 3 - 1.5
    X, Y, Z = torch.meshgrid(c, c, c, indexing="ij")
    est.binaries = (X * X + Y * Y + Z * Z <= 0.6 ** 2).unsqueeze(0)
    est.occs = est.binaries.reshape(-1).float()
    g = torch.Generator().manual_seed(seed + 1)
    o = torch.randn(n_rays, 3, generator=g)
    o = o / o.norm(dim=-1, keepdim=True) * 4
    d = (torch.rand(n_rays, 3, generator=g) - 0.5) * 0.8 - o
    d = d / d.norm(dim=-1, keepdim=True)
    rays = Rays(o.to(dev), d.to(dev))

    class Teacher(torch.nn.Module):
        """a closed-form scene: the target pixels are its rendering through the same occupancy grid"""

        def query_density(self, x):
            return 30.0 * torch.exp(-8.0 * (x * x).sum(-1, keepdim=True))

        def forward(self, x, dd):
            return torch.sigmoid(4.0 * x), self.query_density(x)

    from cnc_b200.render import render_image_with_occgrid

    teacher = Teacher().to(dev).eval()
    with torch.no_grad():
        pixels = render_image_with_occgrid(teacher, est, rays, render_step_size=5e-3, render_bkgd=torch.zeros(3, device=dev))[0]
    return field, est, rays, pixels


def test_train_step_loss_falls_and_planes_follow(cuda):
    from cnc_b200 import _gridencoder as G
    from cnc_b200.trainer import TrainStep

    field, est, rays, pixels = _scene(cuda)
    ts = TrainStep(field, est, lr=2e-3)
    assert ts.sharded and ts.comm_bytes_per_step() == 0
    bk = torch.zeros(3, device=cuda)
    losses = []
    p0 = field.mlp_base.encoding_xyz.params.detach().clone()
    for _ in range(60):
        loss, n = ts(rays, pixels, render_bkgd=bk, refresh_occupancy=False)
        assert n > 0
        losses.append(float(loss))
    assert all(np.isfinite(losses))
    print("loss, first and last five of 60 steps:", [round(l, 5) for l in losses[:5]], [round(l, 5) for l in losses[-5:]])
    # (the reference's own step on its own kernels goes 0.148 -> 0.001 on this scene in 60 steps: scripts/train_compare.py)
    assert np.mean(losses[-5:]) < 0.05 * losses[0], (losses[:5], losses[-5:])
    assert losses[-1] < 0.9 * float((pixels ** 2).mean()), "no better than a transparent field"
    mb = field.mlp_base
    for enc in (mb.encoding_xyz, mb.encoding_xy, mb.encoding_xz, mb.encoding_yz):
        cached = enc.sign_bits()                              # what the next forward will gather from
        assert torch.equal(cached, G.sign_pack(enc.params.detach().contiguous()))
    moved = (mb.encoding_xyz.params.detach() != p0).any(-1).float().mean()
    assert 0.001 < float(moved) <= 1.0                         # rows the rays touched (plus weight decay on all rows)
    # inference through the fused kernel sees the trained tables (same planes) -> the rendered colours match the loss
    field.eval()
    with torch.no_grad():
        from cnc_b200.render import render_image_with_occgrid

        rgb, _, _, _ = render_image_with_occgrid(field, est, rays, render_step_size=5e-3, render_bkgd=bk)
    assert float(torch.nn.functional.mse_loss(rgb, pixels)) < 1.5 * np.mean(losses[-5:]) + 1e-3


def test_caches_follow_fused_optimizers(cuda):
    """torch.optim.Adam(fused=True) (and any kernel writing through raw pointers) updates parameters without bumping the
    autograd version counter: the packed MLP blob and the sign planes must not be trusted by version during training,
    and a train() -> eval() switch must drop them (round-1 bug: every forward after the first saw step-0 weights)."""
    from cnc_b200.render import render_image_with_occgrid

    field, est, rays, pixels = _scene(cuda, seed=5)
    bk = torch.zeros(3, device=cuda)
    opt = torch.optim.Adam(field.parameters(), lr=5e-3, eps=1e-15, fused=True)

    def render():
        return render_image_with_occgrid(field, est, rays, render_step_size=5e-3, render_bkgd=bk)[0]

    first = None
    for _ in range(3):
        field.train()
        opt.zero_grad(set_to_none=True)
        torch.nn.functional.mse_loss(render(), pixels).backward()
        opt.step()
        field.eval()
        with torch.no_grad():
            got = render()                                   # through the caches
            field.invalidate_caches()
            want = render()                                  # everything re-packed from the parameters
        torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-6)   # (index_add_ accumulation order is not fixed)
        assert float((got - first).abs().max()) > 1e-3 if first is not None else True
        first = got if first is None else first


def test_train_step_with_rate_term(cuda):
    from cnc_b200.context_models import CNC_context_models
    from cnc_b200.trainer import TrainStep

    field, est, rays, pixels = _scene(cuda, seed=3)
    cm = CNC_context_models(num_dim=3, resolutions_list=R3, resolutions_list_2D=R2, log2_hashmap_size=19, log2_hashmap_size_2D=17,
                            n_features=8, sample_num=20000, max_context_layer_num=3, ste_binary=True, Rb=128,
                            skip_levels_3D=(0, 1, 2), skip_levels_2D=(0,), device=cuda)
    ts = TrainStep(field, est, context_model=cm, lmbda=1e-3, lr=1e-3)
    w0 = cm.context_model_3D[0].weight.detach().clone()
    for _ in range(3):
        loss, n = ts(rays, pixels, refresh_occupancy=False)
        assert torch.isfinite(loss) and n > 0
    assert not torch.equal(cm.context_model_3D[0].weight.detach(), w0)      # the context model trains with the field
    assert all(torch.isfinite(p).all() for p in field.parameters())


def test_march_ahead_gives_the_same_samples_and_the_same_training(cuda):
    """nerfacc.Premarch (the occupancy march of the next batch on a side stream) hands `sampling` exactly what the in-line
    march produces -- same jitter draw, same kernels -- also when the batch outgrows its buffer; TrainStep(next_rays=...)
    uses it on every step that does not refresh the grid and trains like the plain loop."""
    from cnc_b200.nerfacc import Premarch
    from cnc_b200.trainer import TrainStep

    field, est, rays, pixels = _scene(cuda, n_rays=500)
    field.train()
    sigma_fn = lambda t0, t1, ri: field.query_density(rays.origins[ri] + rays.viewdirs[ri] * ((t0 + t1) / 2)[:, None]).squeeze(-1)
    for cap in (1 << 20, 1000):                      # the second one is outgrown at once
        torch.manual_seed(11)
        want = est.sampling(rays.origins, rays.viewdirs, sigma_fn=sigma_fn, render_step_size=5e-3, stratified=True)
        torch.manual_seed(11)
        pm = Premarch(cuda, capacity=cap)
        pm.issue(est, rays.origins, rays.viewdirs, render_step_size=5e-3, stratified=True)
        assert pm.matches(est, rays.origins, rays.viewdirs, 0.0, 1e10, 5e-3, True, 0.0)
        assert not pm.matches(est, rays.origins, rays.viewdirs, 0.0, 1e10, 1e-2, True, 0.0)
        got = est.sampling(rays.origins, rays.viewdirs, sigma_fn=sigma_fn, render_step_size=5e-3, stratified=True, premarched=pm)
        assert want[0].numel() > 10000
        for a, b in zip(want, got):
            assert torch.equal(a, b)
        assert pm.pending is None and pm.taken == 1
    # the training loop with the look-ahead: used on the steps between grid refreshes, and the loss falls the same way
    curves = []
    for ahead in (False, True):
        field, est, rays, pixels = _scene(cuda, n_rays=500)
        torch.manual_seed(3)
        ts = TrainStep(field, est, lr=2e-3, occ_refresh_every=4)
        ts.step_id = 1024       # past the estimator's warm-up: a refresh touches a random quarter of the cells
        losses = []
        for _ in range(24):
            loss, n = ts(rays, pixels, render_bkgd=torch.zeros(3, device=cuda), refresh_occupancy=True,
                         next_rays=(lambda n_samples: rays) if ahead else None)
            losses.append(float(loss))
        if ahead:
            assert ts._premarch.taken == 18      # every step but the first and the five that follow a refresh decision
        curves.append(losses)
    assert np.isfinite(curves[1]).all()
    np.testing.assert_allclose(curves[1][0], curves[0][0], rtol=1e-4)            # same first step
    assert np.mean(curves[1][-4:]) < 0.9 * curves[1][0]
    np.testing.assert_allclose(np.mean(curves[1][-4:]), np.mean(curves[0][-4:]), rtol=0.2)
