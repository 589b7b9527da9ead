"""Profiling aid: where the host time of the training step goes (cProfile over a few steps, same batch as train_profile.py)."""
import cProfile, os, pstats, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import torch
import bench
from cnc_b200.render import Rays
from cnc_b200.trainer import TrainStep
dev = torch.device("cuda", 0)
arm = bench.Arm("ours", dev)
field, est = arm.field(seed=0), arm.estimator()
g = torch.Generator(device="cpu").manual_seed(7)
n_rays = 1100
o = torch.randn(n_rays, 3, generator=g); o = o / o.norm(dim=-1, keepdim=True) * 4
tgt = (torch.rand(n_rays, 3, generator=g) - 0.5) * 1.2
d = tgt - o; d = d / d.norm(dim=-1, keepdim=True)
rays = Rays(o.to(dev), d.to(dev)); pixels = torch.rand(n_rays, 3, generator=g).to(dev)
ts = TrainStep(field, est, lr=1e-4)
AHEAD = (lambda n: rays) if os.environ.get("AHEAD", "1") == "1" else None
for _ in range(5):
    ts(rays, pixels, refresh_occupancy=False, next_rays=AHEAD)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(10):
    ts(rays, pixels, refresh_occupancy=False, next_rays=AHEAD)
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr); st.sort_stats(os.environ.get("SORT", "tottime")).print_stats(40)
