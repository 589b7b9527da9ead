"""Debug helper: run the fused field kernel on a small batch and print error statistics."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from test_gpu_field import make_field, inputs, double_reference, relerr

dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
f = make_field(dev, wscale=float(sys.argv[2]) if len(sys.argv) > 2 else 1.0)
pos, dirs = inputs(n, dev)
_, sig1, geo1 = f.fused_forward(pos, None, return_feat=True)
torch.cuda.synchronize()
print("density-only kernel ran")
rgb_d, sig_d, geo_d = double_reference(f, pos, dirs)
print("sigma relerr", relerr(sig1, sig_d), "geo relerr", relerr(geo1, geo_d))
print("geo sample", geo1[2, :6].tolist(), geo_d[2, :6].tolist())
rgb, sig, geo = f.fused_forward(pos, dirs, return_feat=True)
torch.cuda.synchronize()
print("full kernel ran")
print("rgb relerr", relerr(rgb, rgb_d), "sigma", relerr(sig, sig_d), "geo", relerr(geo, geo_d))
print("rgb sample", rgb[2].tolist(), rgb_d[2].tolist())
bad = ((rgb.double() - rgb_d).abs().max(-1).values > 1e-4).nonzero().flatten()
print("bad rows", bad.numel(), bad[:20].tolist())
