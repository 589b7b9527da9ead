"""device-loop test renderer on the bench view (timing + launches for ncu)"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import torch, bench
dev = torch.device("cuda:0")
arm = bench.Arm("ours", dev)
field, est = arm.field().eval(), arm.estimator()
H = W = 256
v, u = torch.meshgrid(torch.linspace(-0.3, 0.3, H, device=dev), torch.linspace(-0.3, 0.3, W, device=dev), indexing="ij")
dirs = torch.stack([u, v, torch.ones_like(u)], -1); dirs = dirs / dirs.norm(dim=-1, keepdim=True)
img = arm.Rays(torch.tensor([0.0, 0.0, -4.0], device=dev).expand(H, W, 3).contiguous(), dirs.contiguous())
kw = dict(render_step_size=5e-3, render_bkgd=torch.ones(3, device=dev))
for rep in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = arm.render_test(1024, field, est, img, **kw); e1.record(); torch.cuda.synchronize()
    print(f"device loop: {e0.elapsed_time(e1):.2f} ms, {out[3]} samples")
