import sys, os
R=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0,R); sys.path.insert(0,os.path.join(R,'tests'))
import numpy as np, torch
from conftest import R3
from oracle import oracle as o, ref_ext
from cnc_b200 import _gridencoder as G
ref=ref_ext.load("_gridencoder")
dev=torch.device('cuda:0')
torch.manual_seed(308)
offs_np=o.grid_layout(3,R3,19)
offs=torch.from_numpy(offs_np).to(dev); rl=torch.tensor(R3,dtype=torch.int32,device=dev)
N=100003; L=12; F=8
x=torch.rand(N,3,device=dev)
x[:8] = torch.tensor([[i, j, k] for i in (0., 1.) for j in (0., 1.) for k in (0., 1.)], device=dev)
tab=torch.randn(int(offs[-1]),F,device=dev)
a=torch.empty(L,N,F,device=dev); b=torch.empty(L,N,F,device=dev)
ref.grid_encode_forward(x,tab,offs,rl,a,N,3,F,L,0,128,0.0,None,None,None)
G.grid_encode_forward(x,tab,offs,rl,b,N,3,F,L,0,128,0.0,None,None,None)
torch.cuda.synchronize()
d=(a-b).abs()
print('max abs diff',d.max().item(),'frac equal',(a==b).float().mean().item())
for l in range(L):
    bad=(d[l].max(-1).values>1e-6).nonzero().flatten()
    print('level',l,'res',R3[l],'bad points',bad.numel(), 'maxdiff', d[l].max().item(), 'first',bad[:5].tolist())
bad=(d.max(-1).values>1e-6).nonzero()
if bad.numel():
    l,i=bad[0].tolist()
    print('example level',l,'pt',i,x[i].tolist(),'ref',a[l,i].tolist(),'ours',b[l,i].tolist())
    orc=o.grid_encode_fwd(x[i:i+1].cpu().numpy(),tab.cpu().numpy(),offs_np,R3,12)
    print('oracle',orc[l,0].tolist())
    xs=x[i].cpu().numpy(); print('pos*scale+0.5', xs*np.float32(R3[l]-2)+np.float32(0.5))
print("---- after fix: detailed example")
ne=(a!=b)
print('frac neq', ne.float().mean().item())
for l in range(L):
    print('level',l,'neq frac',ne[l].float().mean().item())
idx=ne.nonzero()
l,i,ch=idx[0].tolist()
print('example level',l,'pt',i,'ch',ch,'x',x[i].tolist())
print(' ref ',a[l,i].tolist()); print(' ours',b[l,i].tolist())
# recompute in numpy variants
res=R3[l]; s=np.float32(res-2); T=int(offs_np[l+1]-offs_np[l])
xs=x[i].cpu().numpy().astype(np.float32)
import math
pos=np.array([np.float32(math.fma(float(xs[d]),float(s),0.5)) if hasattr(math,'fma') else np.float32(np.float64(xs[d])*np.float64(s)+0.5) for d in range(3)],np.float32)
g=np.floor(pos).astype(np.uint32); f=(pos-g.astype(np.float32)).astype(np.float32)
print(' pos',pos,'g',g,'f',f)
tabc=tab.cpu().numpy()
ws=[];rows=[];valid=[]
for k in range(8):
    w=np.float32(1); c=[]
    for d in range(3):
        if (k>>d)&1: w=np.float32(w*f[d]); c.append(min(g[d]+1,res-1))
        else: w=np.float32(w*np.float32(np.float32(1)-f[d])); c.append(g[d])
    z=any(cc==0 or cc==res-1 for cc in c)
    ws.append(w); valid.append(not z)
    rows.append(int(o.grid_rows(np.array([c],np.uint32),T,res)[0]))
print(' w',ws,'valid',valid)
wn=np.float32(0)
for k in range(8):
    if valid[k]: wn=np.float32(wn+ws[k])
print(' wn',wn, 'sum64',sum(float(w) for w,v in zip(ws,valid) if v))
