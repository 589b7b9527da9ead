"""Profiling aid: GPU idle gaps of a torch.profiler chrome trace (which kernels ran, where the device waited for the host)."""
import json, sys
tr = json.load(open(sys.argv[1]))
ev = [e for e in tr["traceEvents"] if e.get("ph") == "X"]
k = sorted([e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")], key=lambda e: e["ts"])
cpu = sorted([e for e in ev if e.get("cat") in ("cpu_op", "user_annotation", "cuda_runtime", "cuda_driver")], key=lambda e: e["ts"])
t0, t1 = k[0]["ts"], max(e["ts"] + e["dur"] for e in k)
print(f"{len(k)} device activities over {(t1 - t0) / 1e3:.2f} ms")
# union of busy intervals over all streams
busy, cur_s, cur_e = 0.0, None, None
gaps = []
for e in k:
    s, f = e["ts"], e["ts"] + e["dur"]
    if cur_e is None:
        cur_s, cur_e = s, f
    elif s <= cur_e:
        cur_e = max(cur_e, f)
    else:
        busy += cur_e - cur_s
        gaps.append((s - cur_e, cur_e, s, e["name"][:60]))
        cur_s, cur_e = s, f
busy += cur_e - cur_s
print(f"busy {busy / 1e3:.2f} ms, idle {(t1 - t0 - busy) / 1e3:.2f} ms ({100 * (1 - busy / (t1 - t0)):.1f} %)")
gaps.sort(reverse=True)
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 15
print(f"gaps > {thr} us: {sum(1 for g in gaps if g[0] > thr)}, total {sum(g[0] for g in gaps if g[0] > thr) / 1e3:.2f} ms; "
      f"gaps <= {thr} us: {sum(1 for g in gaps if g[0] <= thr)}, total {sum(g[0] for g in gaps if g[0] <= thr) / 1e3:.2f} ms")
for g, a, b, name in sorted(gaps[:40], key=lambda x: x[1]):
    # what the host was doing during the gap
    ops = [c["name"][:40] for c in cpu if c["ts"] < b and c["ts"] + c["dur"] > a and c.get("cat") == "cpu_op"][:6]
    print(f"  {g:8.1f} us at +{(a - t0) / 1e3:7.3f} ms before {name:60s} host: {ops}")
