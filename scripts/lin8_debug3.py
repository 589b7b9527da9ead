import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import torch
from conftest import R2, R3
from test_gpu_codec import make
from cnc_b200 import context_models as CM
dev = torch.device("cuda:0")
orig = CM._linear8
def dbg(lin, x):
    y = orig(lin, x)
    ref = torch.nn.functional.linear(x.double(), lin.weight.double(), lin.bias.double())
    y2 = lin(x)
    print("K", x.shape[1], "N", x.shape[0], "fwd err lin8", float((y.double() - ref).abs().max() / ref.abs().max()), "torch fp32", float((y2.double() - ref).abs().max() / ref.abs().max()),
          "x absmax", float(x.abs().max()), "finite", bool(torch.isfinite(x).all()))
    return y
CM._linear8 = dbg
cm, encs, vxl = make(dev, res3=R3, log2T=19, res2=R2, log2T2=17, Rb=128, seed=2)
torch.manual_seed(9)
bpp, _ = cm.forward_binary_vxl_mixPg_3D2D(*encs, vxl, step=0, sample_num=20000)
