"""Validation + timing aid for the fused context MLP (cnc_ctx_mlp_fwd / _bwd): op-level errors against fp64 autograd,
then the rate term with and without it (same loss, same gradients, time per step)."""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import torch
from cnc_b200.context_models import _CtxMLP3
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = torch.nn.Sequential(torch.nn.Linear(25, 32), torch.nn.LeakyReLU(), torch.nn.Linear(32, 32), torch.nn.LeakyReLU(), torch.nn.Linear(32, 8)).to(dev)
ok = True
for M in (1, 255, 256, 257, 100077):
    x = torch.randn(M, 25, device=dev, requires_grad=True)
    gy = torch.randn(M, 8, device=dev)
    ps = [net[0].weight, net[0].bias, net[2].weight, net[2].bias, net[4].weight, net[4].bias]
    y = _CtxMLP3.apply(x, *ps)
    got = torch.autograd.grad(y, [x] + ps, gy)
    net64 = torch.nn.Sequential(torch.nn.Linear(25, 32), torch.nn.LeakyReLU(), torch.nn.Linear(32, 32), torch.nn.LeakyReLU(), torch.nn.Linear(32, 8)).to(dev).double()
    net64.load_state_dict({k: v.double() for k, v in net.state_dict().items()})
    x64 = x.detach().double().requires_grad_(True)
    y64 = net64(x64)
    ps64 = [net64[0].weight, net64[0].bias, net64[2].weight, net64[2].bias, net64[4].weight, net64[4].bias]
    want = torch.autograd.grad(y64, [x64] + ps64, gy.double())
    y32 = net(x)
    ref32 = torch.autograd.grad(y32, [x] + ps, gy)
    def rel(a, b): return ((a.double() - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()
    errs = [rel(y, y64)] + [rel(a, b) for a, b in zip(got, want)]
    errs32 = [rel(y32, y64)] + [rel(a, b) for a, b in zip(ref32, want)]
    good = max(errs) < 2e-5
    ok = ok and good
    print(f"M={M}: fused max rel err {max(errs):.2e} (y {errs[0]:.1e}, gx {errs[1]:.1e}, gW {max(errs[2:]):.1e}); torch fp32 {max(errs32):.2e}  {'ok' if good else 'FAIL'}")
print("OP_CHECK", "PASS" if ok else "FAIL")

from conftest import R2, R3
from test_gpu_codec import make
cm, encs, vxl = make(dev, res3=R3, log2T=19, res2=R2, log2T2=17, Rb=128, seed=1)
params = [e.params for e in encs] + list(cm.parameters())
SN = int(os.environ.get("SAMPLE_NUM", "150000"))
def step(i, seed):
    torch.manual_seed(seed)
    for p in params: p.grad = None
    bpp, mb = cm.forward_binary_vxl_mixPg_3D2D(*encs, vxl, step=i, sample_num=SN)
    bpp.backward()
    return bpp.detach().clone(), [None if p.grad is None else p.grad.detach().clone() for p in params]
res = {}
for fused in (False, True):
    cm.fused_mlp_train = fused
    for i in range(3): step(i + 1, 100 + i)
    res[fused] = step(17, 7)
    torch.cuda.synchronize(); t = time.perf_counter()
    for i in range(5): step(i + 18, 200 + i)
    torch.cuda.synchronize()
    print(f"fused_mlp_train={fused}: rate term {(time.perf_counter() - t) / 5 * 1e3:.2f} ms/step at sample_num={SN}")
b0, g0 = res[False]; b1, g1 = res[True]
print("bpp", b0.item(), b1.item(), "rel diff", abs(b0.item() - b1.item()) / abs(b0.item()))
worst = 0.0
for a, b in zip(g0, g1):
    if a is None: continue
    worst = max(worst, ((a - b).norm() / a.norm().clamp_min(1e-30)).item())
print("gradient rel diff (norm-wise, worst parameter):", worst)
print("RATE_CHECK", "PASS" if abs(b0.item() - b1.item()) / abs(b0.item()) < 1e-5 and worst < 1e-3 else "FAIL")
