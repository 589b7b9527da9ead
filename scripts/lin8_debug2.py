import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import torch
from conftest import R2, R3
from test_gpu_codec import make
from cnc_b200 import context_models as CM
dev = torch.device("cuda:0")
orig_bwd = CM._Lin8.backward
def dbg(ctx, gy):
    x, Wc = ctx.saved_tensors
    out = orig_bwd(ctx, gy)
    gx_ref = gy.double() @ Wc.double()
    gW_ref = gy.double().t() @ x.double()
    print("K", x.shape[1], "N", x.shape[0], "gy contiguous", gy.is_contiguous(), gy.stride(), "gx err", float((out[0].double() - gx_ref).abs().max() / gx_ref.abs().max()),
          "gW err", float((out[1].double() - gW_ref).abs().max() / gW_ref.abs().max()), "gy nan", bool(torch.isnan(gy).any()), "x nan", bool(torch.isnan(x).any()))
    return out
CM._Lin8.backward = staticmethod(dbg)
cm, encs, vxl = make(dev, res3=R3, log2T=19, res2=R2, log2T2=17, Rb=128, seed=2)
torch.manual_seed(9)
bpp, _ = cm.forward_binary_vxl_mixPg_3D2D(*encs, vxl, step=0, sample_num=20000)
bpp.backward()
