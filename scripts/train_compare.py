"""Diagnostic: loss curves of TrainStep (ours) and of the reference step (bench.RefTrainStep on the unmodified reference
objects) on the same closed-form teacher scene.  Adam's first steps are sign-like, so the trajectories differ in detail;
the curves should agree statistically."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import torch
import bench
from test_gpu_train import _scene

dev = torch.device("cuda:0")
steps, lr = int(os.environ.get("STEPS", 60)), float(os.environ.get("LR", 2e-3))
field, est, rays, pixels = _scene(dev)
print("mean(pixels^2) =", float((pixels ** 2).mean()), " var =", float(pixels.var()))
bk = torch.zeros(3, device=dev)
from cnc_b200.trainer import TrainStep
ts = TrainStep(field, est, lr=lr)
ours = [float(ts(rays, pixels, render_bkgd=bk, refresh_occupancy=False)[0]) for _ in range(steps)]
print("ours     ", [round(x, 4) for x in ours[:8]], "...", [round(x, 4) for x in ours[-4:]])
arm = bench.Arm("reference", dev)
torch.manual_seed(0)
rf = arm.Field(aabb=[-1.5] * 3 + [1.5] * 3, n_features_per_level=8, n_neurons=160, resolutions_list=bench.R3, log2_hashmap_size=19,
               resolutions_list_2D=bench.R2, log2_hashmap_size_2D=17, ste_binary=True).to(dev)
re = arm.estimator()
re.binaries = est.binaries.clone(); re.occs = est.occs.clone()
rr = arm.Rays(rays.origins, rays.viewdirs)
class T(bench.RefTrainStep):
    pass
rt = bench.RefTrainStep(arm, rf, re, lr=lr)
# RefTrainStep renders with a white background: use black like ours
import types
def call(self, rays, pixels):
    self.field.train(); self.est.train()
    rgb, acc, depth, n = self.arm.render_train(self.field, self.est, rays, render_step_size=5e-3, render_bkgd=bk)
    loss = torch.nn.functional.mse_loss(rgb, pixels)
    self.opt.zero_grad(); (loss * 1024.0).backward()
    for p in self.field.parameters():
        if p.grad is not None: p.grad /= 1024.0
    self.opt.step(); return loss.detach()
ref = [float(call(rt, rr, pixels)) for _ in range(steps)]
print("reference", [round(x, 4) for x in ref[:8]], "...", [round(x, 4) for x in ref[-4:]])
