"""Wall-clock of the public encode/decode calls at the product layout (no instrumentation inside)."""
import sys, os, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import torch
from conftest import R2, R3
from test_gpu_codec import make
dev = torch.device("cuda:0")
cm, encs, vxl = make(dev, res3=R3, log2T=19, res2=R2, log2T2=17, Rb=128, seed=1)
for it in range(3):
    torch.cuda.synchronize(); t = time.perf_counter()
    Pgs, est, coded, streams = cm.encode_binary_vxl_mixPg_3D2D(*encs, vxl, "p", return_streams=True)
    torch.cuda.synchronize(); te = time.perf_counter() - t
    recs = [torch.ones_like(e.params) for e in encs]
    torch.cuda.synchronize(); t = time.perf_counter()
    out = cm.decode_binary_vxl_mixPg_3D2D(*encs, *recs, vxl, Pgs, "p", streams=streams)
    torch.cuda.synchronize(); td = time.perf_counter() - t
    ok = all(bool(((torch.where(e.params >= 0, 1.0, -1.0) == r) | (r == 1)).all()) for e, r in zip(encs, out))
    print(f"encode {te:.3f}s decode {td:.3f}s roundtrip {ok} coded {coded:.3f} MiB")
