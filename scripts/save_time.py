"""Profiling aid: training-mode forward (fused kernel + activation stores) vs inference forward, 380 k samples."""
import os, sys
sys.path.insert(0, os.getcwd())
import torch, bench
dev = torch.device("cuda:0")
f = bench.build_field(dev).train()
pos, dirs = bench.make_inputs(380000, 1, dev)
for _ in range(3): f(pos, dirs)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): f(pos, dirs)
e1.record(); torch.cuda.synchronize()
print("train forward (SAVE) ms:", e0.elapsed_time(e1) / 10)
f.eval()
with torch.no_grad():
    for _ in range(3): f(pos, dirs)
    e0.record()
    for _ in range(10): f(pos, dirs)
    e1.record(); torch.cuda.synchronize()
print("inference forward ms:", e0.elapsed_time(e1) / 10)
