"""Profiling aid: inference forward (cnc_field_fwd) and density-only forward at 262 144 samples, plus an output checksum
(variants of the kernel that only change the schedule must reproduce it bit for bit)."""
import os, sys, hashlib
sys.path.insert(0, os.getcwd())
import torch, bench
dev = torch.device("cuda:0")
f = bench.build_field(dev).eval()
pos, dirs = bench.make_inputs(262144, 1, dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.no_grad():
    for _ in range(5): out = f(pos, dirs)
    torch.cuda.synchronize()
    ts = []
    for rep in range(3):
        e0.record()
        for _ in range(20): out = f(pos, dirs)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / 20)
    print("forward ms (back to back, 3 x 20):", " ".join(f"{t:.4f}" for t in ts))
    h = hashlib.sha1()
    for t in (out if isinstance(out, (tuple, list)) else [out]):
        h.update(t.detach().cpu().numpy().tobytes())
    print("output sha1:", h.hexdigest()[:16])
    for _ in range(3): d = f.query_density(pos)
    e0.record()
    for _ in range(20): d = f.query_density(pos)
    e1.record(); torch.cuda.synchronize()
    print(f"density-only ms: {e0.elapsed_time(e1) / 20:.4f}")
    dd = d[0] if isinstance(d, (tuple, list)) else d
    print("density sha1:", hashlib.sha1(dd.detach().cpu().numpy().tobytes()).hexdigest()[:16])
