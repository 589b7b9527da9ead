"""GPU suite: the fused field forward (cnc_field_fwd: encode + 3xTF32 tcgen05 MLPs) against
(a) the unfused path (our gather kernels + torch fp32 nn.Linear, the reference's data flow,
ngp.py:514-566) and (b) a float64 evaluation of the same network on the exact fp32 features.
Tolerance: 1e-5 relative (BASELINE north_star) on sigma / rgb / geo."""
import numpy as np
import pytest
import torch

from conftest import R2, R3

pytestmark = pytest.mark.gpu


def make_field(dev, seed=0, wscale=1.0):
    from cnc_b200.field import NGPRadianceField_mygrid_2D3D

    torch.manual_seed(seed)
    f = NGPRadianceField_mygrid_2D3D(aabb=[-1.5, -1.5, -1.5, 1.5, 1.5, 1.5], n_features_per_level=8, n_neurons=160,
                                     resolutions_list=R3, log2_hashmap_size=19, resolutions_list_2D=R2,
                                     log2_hashmap_size_2D=17, ste_binary=True).to(dev)
    with torch.no_grad():
        for k in ("xyz", "xy", "xz", "yz"):
            p = getattr(f.mlp_base, f"encoding_{k}").params
            p.copy_(torch.where(torch.rand_like(p) < 0.5, -0.5, 0.5))
        if wscale != 1.0:
            for m in list(f.mlp_base.network) + list(f.mlp_head):
                if isinstance(m, torch.nn.Linear):
                    m.weight.mul_(wscale)
                    m.bias.mul_(wscale)
    return f.eval()


def inputs(n, dev, seed=1):
    g = torch.Generator(device="cpu").manual_seed(seed)
    pos = (torch.rand(n, 3, generator=g) * 3.2 - 1.6)  # a few percent outside the +-1.5 aabb
    pos[:8] = torch.tensor([[-1.5, 0, 0], [1.5, 0, 0], [0, 0, 0], [1.5, 1.5, 1.5], [-1.5, -1.5, -1.5],
                            [0.3, -1.7, 0.2], [1.4999, 0.1, -0.2], [0, 0, 1.6]])[: min(8, n)]
    d = torch.randn(n, 3, generator=g)
    d = d / d.norm(dim=-1, keepdim=True)
    return pos.to(dev), d.to(dev)


def double_reference(f, pos, dirs):
    """fp64 evaluation of ngp.py:514-566 on the exact fp32 features / SH values."""
    with torch.no_grad():
        x = f._normalise(pos)
        sel = ((x > 0.0) & (x < 1.0)).all(dim=-1)
        feats = f.mlp_base.features(x.view(-1, 3)).double()
        net = f.mlp_base.network
        h = torch.relu(feats @ net[0].weight.double().T + net[0].bias.double()) @ net[2].weight.double().T + net[2].bias.double()
        sigma = torch.exp(h[:, :1] - 1) * sel[:, None]
        geo = h[:, 1:]
        sh = f.direction_encoding((dirs + 1.0) / 2.0).double()
        hd = f.mlp_head
        z = torch.cat([sh, geo], -1)
        z = torch.relu(z @ hd[0].weight.double().T + hd[0].bias.double())
        z = torch.relu(z @ hd[2].weight.double().T + hd[2].bias.double())
        rgb = torch.sigmoid(z @ hd[4].weight.double().T + hd[4].bias.double())
        return rgb, sigma, geo


def relerr(a, b):
    """max |a-b| / max(|b|, typical magnitude of b): relative error that does not blow up on values
    that happen to cancel to ~0."""
    scale = b.abs().mean().clamp_min(1e-30)
    return ((a.double() - b).abs() / torch.maximum(b.abs(), scale)).max().item()


@pytest.mark.parametrize("n,wscale", [(1, 1.0), (127, 1.0), (1000, 1.0), (4096 + 77, 3.0)])
def test_fused_forward_matches_unfused_and_fp64(cuda, n, wscale):
    f = make_field(cuda, wscale=wscale)
    assert f.fused_available()
    pos, dirs = inputs(n, cuda)
    rgb_f, sig_f, geo_f = f.fused_forward(pos, dirs, return_feat=True)
    f.fused = False
    with torch.no_grad():
        rgb_u, sig_u = f(pos, dirs)
    f.fused = True
    rgb_d, sig_d, geo_d = double_reference(f, pos, dirs)
    torch.cuda.synchronize()
    # contract: 1e-5 relative against the exact result; the fp32 cuBLAS path is shown for scale
    e_f = (relerr(rgb_f, rgb_d), relerr(sig_f, sig_d), relerr(geo_f, geo_d))
    e_u = (relerr(rgb_u, rgb_d), relerr(sig_u, sig_d))
    print(f"n={n} fused rel err rgb/sigma/geo = {e_f}, unfused fp32 rgb/sigma = {e_u}")
    assert max(e_f) <= 1e-5
    torch.testing.assert_close(rgb_f, rgb_u, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(sig_f, sig_u, rtol=2e-5, atol=1e-7)
    # selector: samples outside the aabb have zero density on both paths
    x = f._normalise(pos)
    out = ~((x > 0) & (x < 1)).all(-1)
    assert out.any() or n < 8
    assert (sig_f[out] == 0).all()


def test_fused_density_only_and_module_dispatch(cuda):
    f = make_field(cuda, seed=3)
    pos, dirs = inputs(3000, cuda, seed=5)
    with torch.no_grad():
        d1 = f.query_density(pos)                      # fused, density-only kernel
        d2, g2 = f.query_density(pos, return_feat=True)
        rgb, d3 = f(pos, dirs)                         # fused, full kernel
    assert d1.shape == (3000, 1) and g2.shape == (3000, 79) and rgb.shape == (3000, 3)
    assert torch.equal(d1, d2) and torch.equal(d1, d3)  # same arithmetic for layers 1-2 in both kernels
    _, sig_d, geo_d = double_reference(f, pos, dirs)
    assert relerr(d1, sig_d) <= 1e-5 and relerr(g2, geo_d) <= 1e-5
    # with autograd enabled the module takes the differentiable (unfused) path and agrees
    rgb_g, d_g = f(pos, dirs)
    assert rgb_g.requires_grad
    torch.testing.assert_close(rgb_g.detach(), rgb, rtol=1e-5, atol=1e-6)
    # weights change -> the packed blob follows
    with torch.no_grad():
        f.mlp_head[4].bias.add_(0.5)
        rgb2, _ = f(pos, dirs)
    assert (rgb2 > rgb).all()


def test_fused_full_size_properties(cuda):
    """product size (262144 samples, 2048 tiles over the persistent grid): identical results for a
    permuted batch (tile/CTA assignment independent) and run-to-run determinism."""
    f = make_field(cuda, seed=7)
    pos, dirs = inputs(262144, cuda, seed=11)
    rgb, sig, _ = f.fused_forward(pos, dirs)
    rgb_b, sig_b, _ = f.fused_forward(pos, dirs)
    assert torch.equal(rgb, rgb_b) and torch.equal(sig, sig_b)
    perm = torch.randperm(pos.shape[0], device=cuda)
    rgb_p, sig_p, _ = f.fused_forward(pos[perm], dirs[perm])
    assert torch.equal(rgb_p, rgb[perm]) and torch.equal(sig_p, sig[perm])
    sub = slice(100000, 101000)
    rgb_d, sig_d, _ = double_reference(f, pos[sub], dirs[sub])
    assert relerr(rgb[sub], rgb_d) <= 1e-5 and relerr(sig[sub], sig_d) <= 1e-5
    assert torch.isfinite(rgb).all() and torch.isfinite(sig).all()


def test_forward_host_pipeline_matches_device_call(cuda):
    """host-buffer entry point (chunked, three streams) == one device call, for sizes around the chunk boundaries"""
    f = make_field(cuda, seed=9)
    n_sm = torch.cuda.get_device_properties(cuda).multi_processor_count
    for n in (1000, n_sm * 128, 2 * n_sm * 128 + 77):
        pos, dirs = inputs(n, cuda, seed=n)
        rgb, sig, _ = f.fused_forward(pos, dirs)
        hp, hd = pos.cpu().pin_memory(), dirs.cpu().pin_memory()
        r1, s1 = f.forward_host(hp, hd, chunk_waves=1)
        torch.cuda.synchronize()
        assert torch.equal(r1, rgb.cpu()) and torch.equal(s1, sig.cpu())
        r2, s2 = f.forward_host(hp, hd)   # buffers are reused across calls
        torch.cuda.synchronize()
        assert torch.equal(r2, rgb.cpu()) and torch.equal(s2, sig.cpu())


def test_forward_host_two_slots_on_two_streams(cuda):
    """independent batches issued alternately with slot 0 / 1 from two streams overlap (upload of one beside the kernel of
    the other) and still return exactly the device results, call after call"""
    f = make_field(cuda, seed=11)
    n_sm = torch.cuda.get_device_properties(cuda).multi_processor_count
    n = 5 * n_sm * 128 + 333
    batches = []
    for k in range(2):
        pos, dirs = inputs(n, cuda, seed=50 + k)
        rgb, sig, _ = f.fused_forward(pos, dirs)
        batches.append((pos.cpu().pin_memory(), dirs.cpu().pin_memory(), rgb.cpu(), sig.cpu(),
                        torch.empty(n, 3).pin_memory(), torch.empty(n, 1).pin_memory()))
    streams = [torch.cuda.Stream(cuda), torch.cuda.Stream(cuda)]
    torch.cuda.synchronize()
    for rep in range(6):
        k = rep & 1
        hp, hd, _, _, o_rgb, o_sig = batches[k]
        o_rgb.zero_(); o_sig.zero_()
        with torch.cuda.stream(streams[k]):
            f.forward_host(hp, hd, o_rgb, o_sig, slot=k)
        if rep >= 1:   # the previous batch (other stream, other slot) is checked while this one is in flight
            j = 1 - k
            streams[j].synchronize()
            assert torch.equal(batches[j][4], batches[j][2]) and torch.equal(batches[j][5], batches[j][3])
    torch.cuda.synchronize()


@pytest.mark.parametrize("n", [1000, 20000])
def test_fused_training_path_matches_unfused_autograd(cuda, n):
    """`_FusedFieldTrain` (fused-kernel forward + explicit backward) against torch autograd through the op-by-op
    path (K1/K2 + nn.Linear, the reference's data flow): same outputs (1e-5) and the same gradient for every
    table and MLP parameter (fp32 GEMMs in a different association: 2e-4 of the gradient's scale)."""
    f = make_field(cuda, seed=3)
    f.train()
    pos, dirs = inputs(n, cuda, seed=5)
    g = torch.Generator(device="cpu").manual_seed(9)
    w_rgb, w_sig = torch.randn(n, 3, generator=g).to(cuda), torch.randn(n, 1, generator=g).to(cuda)
    params = dict(f.named_parameters())
    grads = {}
    for mode in (True, False):
        f.fused_train = mode
        for p in params.values():
            p.grad = None
        rgb, sigma = f(pos, dirs)
        (rgb * w_rgb).sum().add((torch.log1p(sigma) * w_sig).sum()).backward()
        grads[mode] = ({k: p.grad.clone() for k, p in params.items()}, rgb.detach().clone(), sigma.detach().clone())
    (ga, rgb_a, sig_a), (gb, rgb_b, sig_b) = grads[True], grads[False]
    assert torch.allclose(rgb_a, rgb_b, rtol=1e-5, atol=1e-6)
    assert torch.allclose(sig_a, sig_b, rtol=1e-5, atol=1e-7)
    for k in params:
        a, b = ga[k], gb[k]
        scale = float(b.abs().max())
        assert scale > 0, k
        # a ReLU whose pre-activation is within rounding of zero may be open in one path and closed in the other
        # (3xTF32 vs cuBLAS summation order): that moves one sample's contribution to a whole weight row, never
        # the bulk -> entry-wise comparison on the small batch (no such sample), norm-wise on the large one
        if n <= 1000:
            assert float((a - b).abs().max()) <= 2e-4 * scale, (k, float((a - b).abs().max()), scale)
        assert float((a - b).norm()) <= 2e-3 * float(b.norm()), (k, float((a - b).norm()), float(b.norm()))
        if "encoding" in k:   # the same rows are touched
            assert float(((a != 0) ^ (b != 0)).float().mean()) < 1e-4, k


@pytest.mark.parametrize("ns,mi,no,ones", [(1, 32, 16, False), (31, 96, 160, True), (1000, 256, 160, False), (4097, 160, 80, True),
                                           (50000, 160, 160, True), (300000, 256, 160, False), (70, 64, 32, True)])
def test_wgrad_matches_fp64(cuda, ns, mi, no, ones):
    """cnc_wgrad (3xTF32 tcgen05, MN-major operands straight from the sample-major activations) against an fp64
    matmul: fp32-equivalent (the error of an fp32 GEMM with this contraction length), leading dimensions honoured."""
    from cnc_b200.field import wgrad

    g = torch.Generator(device="cpu").manual_seed(ns + mi)
    x = torch.randn(ns, mi + 32, generator=g).to(cuda) * torch.rand(1, mi + 32, generator=g).to(cuda) * 3
    z = torch.randn(ns, no + 16, generator=g).to(cuda) * 0.1
    x[:, 0] = 1.0      # a constant column: same-sign sums show any accumulation bias
    z[:, 0] = 0.25
    got = wgrad(x, z, mi, no, with_ones=ones)
    if ones:
        np.testing.assert_allclose(got[mi].double().cpu().numpy(), z[:, :no].double().sum(0).cpu().numpy(),
                                   rtol=0, atol=4e-6 * float(z[:, :no].abs().sum(0).max()))
        got = got[:mi]
    want = x[:, :mi].double().t() @ z[:, :no].double()
    ref32 = x[:, :mi].t() @ z[:, :no]
    scale = (x[:, :mi].double().abs().t() @ z[:, :no].double().abs())       # sum of |terms|: the fp32 error scale
    err = ((got.double() - want).abs() / scale).max().item()
    err32 = ((ref32.double() - want).abs() / scale).max().item()
    assert got.shape == (mi, no)
    assert err <= max(4 * err32, 2e-6), (err, err32)


@pytest.mark.parametrize("ns,no,n,mask", [(1, 4, 32, False), (127, 160, 160, True), (129, 80, 160, True), (5000, 160, 80, False),
                                          (70000, 160, 192, False), (300000, 160, 160, True)])
def test_dgrad_matches_fp64(cuda, ns, no, n, mask):
    """cnc_dgrad (error-compensated tf32 tcgen05, ReLU mask fused) against an fp64 matmul + mask"""
    from cnc_b200.field import dgrad

    g = torch.Generator(device="cpu").manual_seed(ns + n)
    z = (torch.randn(ns, no, generator=g) * 0.3).to(cuda)
    W = (torch.randn(no, n + 20, generator=g) * 0.2).to(cuda)
    h = torch.relu(torch.randn(ns, n, generator=g)).to(cuda) if mask else None
    got = dgrad(z, W, n, h=h, col_off=3, n_first=1)
    want = z.double() @ W[:, 3:3 + n].double()
    want[:, 0] = 0
    scale = z.double().abs() @ W[:, 3:3 + n].double().abs() + 1e-30
    ref32 = z @ W[:, 3:3 + n]
    if mask:
        want = want * (h > 0)
    err = ((got.double() - want).abs() / scale).max().item()
    err32 = (((ref32.double() * ((h > 0) if mask else 1.0)) - z.double() @ W[:, 3:3 + n].double() * ((h > 0) if mask else 1.0)).abs() / scale)[:, 1:].max().item()
    assert err <= max(4 * err32, 2e-6), (err, err32)


def test_fused_training_density_gradient_is_capped_like_trunc_exp(cuda):
    """trunc_exp's backward uses exp(min(h - 1, 15)) (ngp.py:328-334); the fused backward takes the density
    itself as d density / d h, which must be capped at e^15 the same way once the pre-activation exceeds 16."""
    f = make_field(cuda, seed=4)
    f.train()
    with torch.no_grad():
        f.mlp_base.network[2].bias[0] = 19.0      # h - 1 ~ 18 > 15 for every sample
    f.invalidate_caches()
    pos, dirs = inputs(512, cuda, seed=6)
    grads = {}
    for mode in (True, False):
        f.fused_train = mode
        for p in f.parameters():
            p.grad = None
        rgb, sigma = f(pos, dirs)
        assert float(sigma.max()) > 3.3e6          # beyond e^15
        (sigma * 1e-7).sum().backward()
        grads[mode] = f.mlp_base.network[2].weight.grad[0].clone(), f.mlp_base.network[2].bias.grad[0].clone()
    (wa, ba), (wb, bb) = grads[True], grads[False]
    assert torch.isfinite(wa).all() and float(bb) > 0
    assert abs(float(ba) - float(bb)) <= 1e-4 * float(bb)
    assert float((wa - wb).abs().max()) <= 1e-4 * float(wb.abs().max())
    # with the uncapped derivative the bias gradient would be the sum of the densities themselves: ~e^18 / e^15 = 20x larger
    assert float(ba) < 0.2 * float((sigma.detach() * 1e-7).sum())
