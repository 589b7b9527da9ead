"""Profiling aid: wall-clock breakdown of encode/decode at the product layout."""
import sys, os, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import torch
from conftest import R2, R3
from test_gpu_codec import make
from cnc_b200 import torchac as tac
import cnc_b200.context_models as CM

dev = torch.device("cuda:0")
t0 = time.perf_counter()
cm, encs, vxl = make(dev, res3=R3, log2T=19, res2=R2, log2T2=17, Rb=128, seed=1)
torch.cuda.synchronize(); print(f"setup {time.perf_counter()-t0:.3f}s")

T = {}
def timed(name, fn):
    def w(*a, **k):
        torch.cuda.synchronize(); t = time.perf_counter()
        r = fn(*a, **k)
        torch.cuda.synchronize(); T[name] = T.get(name, 0) + time.perf_counter() - t
        return r
    return w
cm.get_idx_coords2 = timed("get_idx_coords2", cm.get_idx_coords2)
cm.get_pn_embed_frac = timed("get_pn_embed_frac", cm.get_pn_embed_frac)
cm._probs_2D = timed("probs_2D", cm._probs_2D)
cm._probs_3D = timed("probs_3D", cm._probs_3D)
cm.get_STE_params = timed("STE", cm.get_STE_params)
tac.decode_streams = timed("decode_streams", tac.decode_streams)
tac.cdf_from_p = timed("cdf_from_p", tac.cdf_from_p)
for it in range(2):
    T.clear()
    torch.cuda.synchronize(); t = time.perf_counter()
    Pgs, est, coded, streams = cm.encode_binary_vxl_mixPg_3D2D(*encs, vxl, "p", return_streams=True)
    torch.cuda.synchronize(); te = time.perf_counter() - t
    print(f"encode {te:.3f}s:", {k: round(v, 3) for k, v in T.items()})
    T.clear()
    recs = [torch.ones_like(e.params) for e in encs]
    torch.cuda.synchronize(); t = time.perf_counter()
    cm.decode_binary_vxl_mixPg_3D2D(*encs, *recs, vxl, Pgs, "p", streams=streams)
    torch.cuda.synchronize(); td = time.perf_counter() - t
    print(f"decode {td:.3f}s:", {k: round(v, 3) for k, v in T.items()})
print({k: len(v) for k, v in list(streams.items())})
