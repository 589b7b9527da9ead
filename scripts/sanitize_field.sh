cat > /tmp/san_field.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import torch
from test_gpu_field import make_field, inputs
dev = torch.device("cuda:0")
f = make_field(dev)
pos, dirs = inputs(3000, dev)
with torch.no_grad():
    rgb, sig = f(pos, dirs)
    d = f.query_density(pos)
f.train()
r, s = f(pos, dirs)
(r.sum() + s.sum()).backward()
hp, hd = pos.cpu().pin_memory(), dirs.cpu().pin_memory()
f.eval()
r2, s2 = f.forward_host(hp, hd)
torch.cuda.synchronize()
assert torch.equal(r2, rgb.cpu())
print("field sanitizer workload ok")
PY
for tool in memcheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_field.py > gpurun_out/sanitize_field_$tool.log 2>&1
  tail -3 gpurun_out/sanitize_field_$tool.log
done
