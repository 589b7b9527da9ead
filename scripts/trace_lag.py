"""Profiling aid: for every kernel of a chrome trace, how long it sat in the queue (device start - host launch).  A short
queue time means the device was waiting for the host at that point."""
import json, sys
tr = json.load(open(sys.argv[1]))
ev = [e for e in tr["traceEvents"] if e.get("ph") == "X"]
launch = {}
for e in ev:
    if e.get("cat") in ("cuda_runtime", "cuda_driver") and "correlation" in e.get("args", {}):
        launch[e["args"]["correlation"]] = e
ks = sorted([e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")], key=lambda e: e["ts"])
t0 = ks[0]["ts"]
lo, hi = float(sys.argv[2]) if len(sys.argv) > 2 else 0, float(sys.argv[3]) if len(sys.argv) > 3 else 1e9
prev_end = None
for k in ks:
    l = launch.get(k["args"].get("correlation"))
    rel = (k["ts"] - t0) / 1e3
    if lo <= rel <= hi:
        lag = k["ts"] - (l["ts"] + l["dur"]) if l else float("nan")
        gap = k["ts"] - prev_end if prev_end is not None else 0
        print(f"+{rel:8.3f} ms  dur {k['dur']:7.1f}  queued {lag:8.1f}  gap {gap:7.1f}  s{k['args'].get('stream')}  {k['name'][:70]}")
    prev_end = max(prev_end or 0, k["ts"] + k["dur"])
