"""Profiling aid: one long stream through the GPU range coder (encode + decode), timed with CUDA events."""
import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import torch
from cnc_b200 import torchac as tac
dev = torch.device("cuda:0")
torch.manual_seed(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ps = [torch.rand(n, device=dev).clamp_(0.02, 0.98) for _ in range(K)]
ps = [torch.where(torch.rand(n, device=dev) < 0.7, p * 0.2 + 0.8, p) for p in ps]   # skewed like a trained context model
syms = [(torch.rand(n, device=dev) < p).to(torch.uint8) for p in ps]
c1 = [tac.cdf_from_p(p) for p in ps]
for _ in range(2):
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record(); data = tac.encode_streams(c1, syms); e1.record(); dec = tac.decode_streams(c1, data); e2.record()
    torch.cuda.synchronize()
    assert all(torch.equal(d, s) for d, s in zip(dec, syms))
    print(f"n={n} x{K}: encode {e0.elapsed_time(e1):.1f} ms ({e0.elapsed_time(e1)*1e6/n:.0f} ns/sym), decode {e1.elapsed_time(e2):.1f} ms ({e1.elapsed_time(e2)*1e6/n:.0f} ns/sym), {len(data[0])*8/n:.3f} bits/sym")
