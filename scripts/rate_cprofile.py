"""Profiling aid: where the host time of the rate term goes (cProfile over a few steps, SAMPLE_NUM as in rate_profile.py)."""
import cProfile, os, pstats, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import torch
from conftest import R2, R3
from test_gpu_codec import make
dev = torch.device("cuda:0")
cm, encs, vxl = make(dev, res3=R3, log2T=19, res2=R2, log2T2=17, Rb=128, seed=1)
SN = int(os.environ.get("SAMPLE_NUM", "0"))
params = [e.params for e in encs] + list(cm.parameters())
def step(i):
    for p in params: p.grad = None
    bpp, mb = cm.forward_binary_vxl_mixPg_3D2D(*encs, vxl, step=i, sample_num=SN or None)
    bpp.backward()
for i in range(3): step(i + 1)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for i in range(5): step(i + 17)
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr); st.sort_stats("tottime").print_stats(28)
