"""standalone K1 (cnc_grid_encode_fwd_bits) on the 3D product table, 262 144 points: timing + a launch for ncu"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import torch, bench
dev = torch.device("cuda:0")
arm = bench.Arm("ours", dev)
f = arm.field()
enc = f.mlp_base.encoding_xyz.eval()
x = torch.rand(262144, 3, device=dev)
with torch.no_grad():
    for _ in range(3): enc(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): enc(x)
    e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(f"K1 3D (12 levels, F=8, sign planes) 262144 points: {ms:.3f} ms  -> {262144 * 3468 / ms / 1e6:.0f} GB/s algorithmic (3468 B/point)")
