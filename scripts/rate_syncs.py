"""Profiling aid: list the host synchronisations inside one rate-term step (torch.cuda.set_sync_debug_mode)."""
import os, sys, warnings, traceback, collections
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import torch
from conftest import R2, R3
from test_gpu_codec import make
dev = torch.device("cuda:0")
cm, encs, vxl = make(dev, res3=R3, log2T=19, res2=R2, log2T2=17, Rb=128, seed=1)
params = [e.params for e in encs] + list(cm.parameters())
def step(i):
    for p in params: p.grad = None
    bpp, mb = cm.forward_binary_vxl_mixPg_3D2D(*encs, vxl, step=i)
    bpp.backward()
for i in range(3): step(i + 1)
torch.cuda.synchronize()
sites = collections.Counter()
def hook(message, category, filename, lineno, file=None, line=None):
    st = [f for f in traceback.extract_stack() if "/cnc_b200/" in f.filename or "/scripts/" in f.filename]
    key = " <- ".join(f"{os.path.basename(f.filename)}:{f.lineno}" for f in reversed(st[-3:]))
    sites[key] += 1
warnings.showwarning = hook
warnings.simplefilter("always")
torch.cuda.set_sync_debug_mode("warn")
step(17)
torch.cuda.set_sync_debug_mode("default")
for k, v in sites.most_common():
    print(v, k)
print("total syncs:", sum(sites.values()))
