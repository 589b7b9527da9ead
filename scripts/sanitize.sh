#!/bin/bash
# compute-sanitizer passes over the kernels with shared-memory protocols (coder, context, wgrad/dgrad): small inputs.
set -x
mkdir -p gpurun_out
cat > /tmp/san_small.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch
from cnc_b200 import torchac as tac
from cnc_b200.field import wgrad, dgrad
from test_gpu_codec import make, SMALL
dev = torch.device("cuda:0")
torch.manual_seed(0)
p = torch.rand(5000, device=dev).clamp_(1e-6, 1 - 1e-6); s = (torch.rand(5000, device=dev) < p).to(torch.uint8)
c1 = tac.cdf_from_p(p); data = tac.encode_streams([c1, c1[:33]], [s, s[:33]]); dec = tac.decode_streams([c1, c1[:33]], data)
assert torch.equal(dec[0], s) and torch.equal(dec[1], s[:33])
x = torch.randn(300, 160, device=dev); z = torch.randn(300, 160, device=dev)
assert torch.allclose(wgrad(x, z, with_ones=True)[:160], x.t() @ z, rtol=1e-4, atol=1e-3)
W = torch.randn(160, 255, device=dev) * 0.1
assert torch.allclose(dgrad(z, W, 160, h=torch.relu(x)), (z @ W[:, :160]) * (x > 0), rtol=1e-4, atol=1e-3)
cm, encs, vxl = make(dev, **SMALL)
Pgs, est, coded, streams = cm.encode_binary_vxl_mixPg_3D2D(*encs, vxl, "s", return_streams=True)
out = cm.decode_binary_vxl_mixPg_3D2D(*encs, *[torch.ones_like(e.params) for e in encs], vxl, Pgs, "s", streams=streams)
print("sanitizer workload ok")
PY
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_small.py > gpurun_out/sanitize_$tool.log 2>&1
  tail -4 gpurun_out/sanitize_$tool.log
done
