"""Profiling aid: host-buffer forward as a stream of independent batches -- one slot / one stream (calls serialise) against
two slots on two streams (the first upload and the last download of a batch hide behind the neighbouring kernels)."""
import os, sys
sys.path.insert(0, os.getcwd())
import torch, bench
dev = torch.device("cuda:0")
f = bench.build_field(dev).eval()
N = 262144
pos, dirs = bench.make_inputs(N, 1, dev)
pin = [(pos.cpu().pin_memory(), dirs.cpu().pin_memory(), torch.empty(N, 3).pin_memory(), torch.empty(N, 1).pin_memory()) for _ in range(2)]
S = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def run(K, two):
    cur = torch.cuda.current_stream(dev)
    e0.record()
    for s in S: s.wait_stream(cur)
    for k in range(K):
        j = (k & 1) if two else 0
        with torch.cuda.stream(S[j]):
            f.forward_host(*pin[j], slot=j)
    for s in S: cur.wait_stream(s)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K
for two in (False, True, False, True):
    run(6, two)
    ms = run(40, two)
    print(f"{'two slots / two streams' if two else 'one slot / one stream  '}: {ms:.4f} ms per 262144 samples = {N / ms * 1e-3:.1f} M samples/s")
for K in (4, 10, 20, 40, 80):
    run(4, True)
    print(f"two slots, K={K}: {run(K, True):.4f} ms per batch")
import time
t = time.perf_counter()
for k in range(40):
    with torch.cuda.stream(S[k & 1]):
        f.forward_host(*pin[k & 1], slot=k & 1)
t1 = time.perf_counter() - t
torch.cuda.synchronize()
print(f"host enqueue time per call: {t1 / 40 * 1e3:.3f} ms")
