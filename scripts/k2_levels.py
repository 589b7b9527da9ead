"""K2 (3D scatter-add) per level group, on ray-ordered samples (what training produces) and on shuffled samples"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import torch, bench
from cnc_b200 import _gridencoder as G
dev = torch.device("cuda:0")
arm = bench.Arm("ours", dev)
f, est = arm.field(), arm.estimator()
g = torch.Generator(device="cpu").manual_seed(7)
n_rays = 1100
o = torch.randn(n_rays, 3, generator=g); o = o / o.norm(dim=-1, keepdim=True) * 4
d = (torch.rand(n_rays, 3, generator=g) - 0.5) * 1.2 - o; d = d / d.norm(dim=-1, keepdim=True)
o, d = o.to(dev), d.to(dev)
ri, t0, t1 = est.sampling(o, d, render_step_size=5e-3)
pos = o[ri] + d[ri] * ((t0 + t1) / 2)[:, None]
x = ((pos + 1.5) / 3.0).contiguous()
N = x.shape[0]
enc = f.mlp_base.encoding_xyz
offs, res = enc.offsets_list, enc.resolutions_list
def run(xx, lo, hi, reps=10):
    L = hi - lo
    grad = torch.randn(L, N, 8, device=dev)
    gt = torch.zeros_like(enc.params)
    for _ in range(2):
        G.grid_encode_backward(grad, xx, enc.params, offs[lo:hi + 1].contiguous(), res[lo:hi].contiguous(), gt, N, 3, 8, L, 0, 128)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        G.grid_encode_backward(grad, xx, enc.params, offs[lo:hi + 1].contiguous(), res[lo:hi].contiguous(), gt, N, 3, 8, L, 0, 128)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
xs = x[torch.randperm(N, device=dev)].contiguous()
print("samples", N)
for lo, hi in ((0, 12), (0, 1), (1, 2), (2, 3), (3, 6), (6, 9), (9, 12), (11, 12)):
    print(f"levels [{lo},{hi}): ray order {run(x, lo, hi):.3f} ms   shuffled {run(xs, lo, hi):.3f} ms")
