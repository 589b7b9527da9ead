import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import torch
from cnc_b200.field import wgrad
dev = torch.device("cuda:0")
torch.set_printoptions(linewidth=200, precision=3)
def probe(ns, mi, no, s0, i0):
    x = torch.zeros(ns, mi, device=dev); z = torch.zeros(ns, no, device=dev)
    x[s0, i0] = 1.0
    z[s0] = torch.arange(1, no + 1, device=dev).float()
    c = wgrad(x, z)
    nz = c.nonzero()
    print(f"ns={ns} mi={mi} no={no} s0={s0} i0={i0}: nonzero rows {sorted(set(nz[:,0].tolist()))[:10]} cols {sorted(set(nz[:,1].tolist()))[:20]}")
    if nz.numel():
        r = nz[0, 0].item(); print("   row", r, c[r, :min(no, 20)].tolist())
probe(8, 32, 16, 0, 0)
probe(8, 32, 16, 0, 5)
probe(8, 32, 16, 3, 0)
probe(8, 32, 16, 3, 9)
probe(64, 64, 32, 40, 37)
probe(64, 256, 160, 33, 200)
x = torch.randn(100, 32, device=dev); z = torch.randn(100, 16, device=dev)
c = wgrad(x, z); w = x.t() @ z
print("random small: max err", (c - w).abs().max().item(), "max |c|", c.abs().max().item(), "max |w|", w.abs().max().item())
x = torch.ones(32, 32, device=dev); z = torch.ones(32, 16, device=dev)
c = wgrad(x, z)
print("all ones 32x32x16:", c.unique().tolist()[:10], c.shape)
x = torch.ones(64, 256, device=dev); z = torch.ones(64, 160, device=dev)
c = wgrad(x, z)
print("all ones 64x256x160:", c.unique().tolist()[:10])
