"""Profiling aid: kernel-level breakdown of one training step (torch.profiler, CUDA activities)."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import torch
from torch.profiler import ProfilerActivity, profile
import bench
from cnc_b200.nerfacc import OccGridEstimator
from cnc_b200.render import Rays
from cnc_b200.trainer import TrainStep

dev = torch.device("cuda:0")
field = bench.build_field(dev)
est = OccGridEstimator(roi_aabb=[-1.5] * 3 + [1.5] * 3, resolution=128, levels=1).to(dev)
c = (torch.arange(128, device=dev) + 0.5) / 128 * 3 - 1.5
X, Y, Z = torch.meshgrid(c, c, c, indexing="ij")
est.binaries = (X * X + Y * Y + Z * Z <= 1.0).unsqueeze(0)
est.occs = est.binaries.reshape(-1).float()
g = torch.Generator(device="cpu").manual_seed(7)
n_rays = int(sys.argv[1]) if len(sys.argv) > 1 else 1100
o = torch.randn(n_rays, 3, generator=g); o = o / o.norm(dim=-1, keepdim=True) * 4
tgt = (torch.rand(n_rays, 3, generator=g) - 0.5) * 1.2
d = tgt - o; d = d / d.norm(dim=-1, keepdim=True)
rays = Rays(o.to(dev), d.to(dev)); pixels = torch.rand(n_rays, 3, generator=g).to(dev)
ts = TrainStep(field, est, lr=1e-4)
for _ in range(3):
    _, n_s = ts(rays, pixels, refresh_occupancy=False)
torch.cuda.synchronize()
import time
t = time.perf_counter()
for _ in range(5):
    ts(rays, pixels, refresh_occupancy=False)
torch.cuda.synchronize()
print(f"samples {n_s}; wall {1e3 * (time.perf_counter() - t) / 5:.2f} ms/step")
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        ts(rays, pixels, refresh_occupancy=False)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
