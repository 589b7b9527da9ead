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

# (works under torchrun too: every rank trains its own ray shard, rank 0 prints)
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    torch.distributed.init_process_group("nccl", device_id=dev)
arm = bench.Arm("ours", dev)
field = arm.field(seed=0)
est = arm.estimator()
g = torch.Generator(device="cpu").manual_seed(7 + rank)
n_rays = int(sys.argv[1]) if len(sys.argv) > 1 else 1100
o = torch.randn(n_rays, 3, generator=g); o = o / o.norm(dim=-1, keepdim=True) * 4
tgt = (torch.rand(n_rays, 3, generator=g) - 0.5) * 1.2
d = tgt - o; d = d / d.norm(dim=-1, keepdim=True)
rays = Rays(o.to(dev), d.to(dev)); pixels = torch.rand(n_rays, 3, generator=g).to(dev)
ts = TrainStep(field, est, lr=1e-4, exchange=os.environ.get("CNC_EXCHANGE", "auto"))
if rank == 0:
    print("exchange:", ts.comm_description())
AHEAD = (lambda n: rays) if os.environ.get("AHEAD", "1") == "1" else None
for _ in range(3):
    _, n_s = ts(rays, pixels, refresh_occupancy=False, next_rays=AHEAD)
torch.cuda.synchronize()
import time
t = time.perf_counter()
for _ in range(5):
    ts(rays, pixels, refresh_occupancy=False, next_rays=AHEAD)
torch.cuda.synchronize()
if rank == 0:
    print(f"samples {n_s}; wall {1e3 * (time.perf_counter() - t) / 5:.2f} ms/step; world {world}")
if os.environ.get("PROFILE", "1") == "0":
    if world > 1:
        torch.distributed.destroy_process_group()
    sys.exit(0)
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        ts(rays, pixels, refresh_occupancy=False, next_rays=AHEAD)
    torch.cuda.synchronize()
if rank == 0 and os.environ.get("TRACE"):
    prof.export_chrome_trace(os.environ["TRACE"])
if rank == 0:
    print(prof.key_averages().table(sort_by=os.environ.get("SORT", "self_cuda_time_total"), row_limit=45, max_name_column_width=70))
    print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=15, max_name_column_width=70))
if world > 1:
    torch.distributed.destroy_process_group()
