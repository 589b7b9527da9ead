import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import torch
from conftest import R2, R3
from test_gpu_codec import make
dev = torch.device("cuda:0")
cm, encs, vxl = make(dev, res3=R3, log2T=19, res2=R2, log2T2=17, Rb=128, seed=2)
out = {}
for fast in (True, False, True):
    cm.fused_lin8 = fast
    for p in [e.params for e in encs] + list(cm.parameters()): p.grad = None
    torch.manual_seed(9)
    bpp, _ = cm.forward_binary_vxl_mixPg_3D2D(*encs, vxl, step=0, sample_num=20000)
    bpp.backward()
    cur = (float(bpp), [e.params.grad.clone() for e in encs], [p.grad.clone() for p in cm.parameters()])
    if fast in out:
        print("repeat (determinism):", [float((a - b).abs().max() / b.abs().max()) for a, b in zip(cur[1], out[fast][1])])
    out[fast] = cur
print("bpp", out[True][0], out[False][0])
print("tables", [float((a - b).abs().max() / b.abs().max()) for a, b in zip(out[True][1], out[False][1])])
print("ctx params", [float((a - b).abs().max() / b.abs().max().clamp_min(1e-30)) for a, b in zip(out[True][2], out[False][2])])
