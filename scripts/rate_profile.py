"""Profiling aid: the rate term of the training loss (forward_binary_vxl_mixPg_3D2D + backward) at the product layout."""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import torch
from torch.profiler import ProfilerActivity, profile
from conftest import R2, R3
from test_gpu_codec import make
dev = torch.device("cuda:0")
cm, encs, vxl = make(dev, res3=R3, log2T=19, res2=R2, log2T2=17, Rb=128, seed=1)
SN = int(os.environ.get("SAMPLE_NUM", "0"))   # the training scripts use 150000 (make(): 4000)
params = [e.params for e in encs] + list(cm.parameters())
def step(i):
    for p in params: p.grad = None
    bpp, mb = cm.forward_binary_vxl_mixPg_3D2D(*encs, vxl, step=i, sample_num=SN or None)
    bpp.backward()
    return bpp
for i in range(3): step(i + 1)
torch.cuda.synchronize(); t = time.perf_counter()
for i in range(5): step(i + 17)
torch.cuda.synchronize(); print(f"rate term fwd+bwd: {(time.perf_counter() - t) / 5 * 1e3:.2f} ms/step (steps not refreshing idx_coords2)")
t = time.perf_counter(); step(16); torch.cuda.synchronize(); print(f"refresh step (step % 16 == 0): {(time.perf_counter() - t) * 1e3:.2f} ms")
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(3): step(i + 33)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by=os.environ.get("SORT", "cuda_time_total"), row_limit=40, max_name_column_width=60))
if os.environ.get("TRACE"):
    prof.export_chrome_trace(os.environ["TRACE"])
