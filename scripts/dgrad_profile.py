import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import torch
from cnc_b200.field import dgrad, wgrad
dev = torch.device("cuda:0")
ns = 380000
z = torch.randn(ns, 160, device=dev); W = torch.randn(160, 255, device=dev) * 0.1; h = torch.relu(torch.randn(ns, 160, device=dev))
for _ in range(3):
    dgrad(z, W, 160, h=h); dgrad(z, W, 192); wgrad(h, z, with_ones=True)
torch.cuda.synchronize()
for name, fn in (("dgrad160+mask", lambda: dgrad(z, W, 160, h=h)), ("dgrad192", lambda: dgrad(z, W, 192)), ("wgrad160x160", lambda: wgrad(h, z, with_ones=True))):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    print(name, e0.elapsed_time(e1) / 10, "ms")
