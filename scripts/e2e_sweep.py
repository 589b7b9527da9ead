import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import torch, bench
dev = torch.device("cuda:0")
f = bench.build_field(dev).eval()
Ns = 262144
pos, dirs = bench.make_inputs(Ns, 1000, dev)
pp, pd = pos.cpu().pin_memory(), dirs.cpu().pin_memory()
o_rgb, o_sig = torch.empty(Ns, 3).pin_memory(), torch.empty(Ns, 1).pin_memory()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for cw in (1, 2, 3, 4, 6, 8, 14):
    for _ in range(3): f.forward_host(pp, pd, o_rgb, o_sig, chunk_waves=cw)
    evs = []
    for _ in range(20):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); f.forward_host(pp, pd, o_rgb, o_sig, chunk_waves=cw); e.record(); evs.append((s, e))
    torch.cuda.synchronize()
    ms = sum(s.elapsed_time(e) for s, e in evs) / 20
    print(f"chunk_waves={cw}: {ms:.3f} ms  {Ns / ms / 1e3:.1f} Msamples/s")
# (tried and dropped: letting the kernel read inputs from / write outputs to pinned host memory directly -- results are
# identical but the 4- and 12-byte accesses per row cross PCIe as tiny transactions: 1.2 ms instead of 0.5 ms)
