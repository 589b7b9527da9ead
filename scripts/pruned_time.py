"""tables='pruned': construction and per-occupancy list build at the product layout (timing + launches for ncu)"""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import torch, bench
dev = torch.device("cuda:0")
arm = bench.Arm("ours", dev)
torch.cuda.synchronize(); t = time.perf_counter()
cm = arm.context_model(tables="pruned")
torch.cuda.synchronize(); print(f"construction {time.perf_counter() - t:.4f} s")
vx = arm.estimator().binaries.squeeze(0).contiguous()
for rep in range(2):
    cm._pruned_cache = None
    torch.cuda.synchronize(); t = time.perf_counter()
    tot = 0
    for n in range(3, 12):
        tot += cm._pruned_level(n, vx)[0].shape[0]
    torch.cuda.synchronize(); print(f"vertex lists of levels 3..11: {time.perf_counter() - t:.4f} s, {tot} vertices, {cm.table_bytes() / 1e6:.1f} MB table state")
