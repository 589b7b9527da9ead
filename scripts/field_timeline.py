"""Profiling aid: per-phase clock64() timeline of one tile of the fused field kernel (CTA 0, 2nd tile)."""
import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import torch
from test_gpu_field import make_field, inputs
from cnc_b200._lib import lib
dev = torch.device("cuda:0")
f = make_field(dev)
pos, dirs = inputs(262144, dev)
for _ in range(3): f.fused_forward(pos, dirs)
buf = torch.zeros(128, dtype=torch.int64, device=dev)
lib().cnc_field_set_timeline_buffer(buf.data_ptr())
f.fused_forward(pos, dirs); torch.cuda.synchronize()
lib().cnc_field_set_timeline_buffer(None)
t = buf.cpu().tolist(); t0 = t[0]
names = {0: "tile start", **{1 + c: f"chunk {c} stored" for c in range(8)}, 10: "L1 done seen", 11: "ep1 done", 12: "L2 done seen",
         13: "ep2 done", 14: "L3 done seen", 15: "ep3 done", 16: "L4 done seen", 17: "ep4 done", 18: "L5 done seen", 19: "ep5 done",
         **{32 + k: f"MMA: A chunk {k} ready" for k in range(8)}, 40: "MMA: L1 issued", 41: "MMA: act1 seen", 42: "MMA: L2 issued",
         43: "MMA: act2 seen", 44: "MMA: act3 seen", 45: "MMA: act4 seen", 46: "MMA: L5 issued", **{48 + k: f"MMA: L2 chunk {k} weights in" for k in range(5)}, **{53 + k: f"MMA: L4 chunk {k} weights in" for k in range(5)}, **{58 + k: f"MMA: L5 chunk {k} weights in" for k in range(5)}, 63: "MMA: tile fully done", **{64 + g: f"PRODUCER: issues load {g}" for g in range(26)}, **{96 + g: f"PRODUCER: load {g} landed" for g in range(26)}, 47: "MMA: L1 done seen by MMA warp", 9: "compute t0 starts waiting for L1", 20: "t0 before load_x", 21: "t0 after load_x"}
for k, v in sorted(((k, v) for k, v in enumerate(t) if v and k != 22), key=lambda kv: kv[1]):
    print(f"{v - t0:8d}  {names.get(k, k)}")
