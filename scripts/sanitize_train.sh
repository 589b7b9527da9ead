#!/bin/bash
# compute-sanitizer over the training step as it is in round 2 (march with warp look-ahead, one-walk packing, look-ahead march on a
# side stream, render backward, sample compaction, training forward with the 96-wide head input, dgrad / wgrad / K2, table Adam)
# on a small ray batch.
set -x
mkdir -p gpurun_out
cat > /tmp/san_train.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import torch
from test_gpu_train import _scene
from cnc_b200.trainer import TrainStep
dev = torch.device("cuda:0")
field, est, rays, pixels = _scene(dev, n_rays=96)
ts = TrainStep(field, est, lr=2e-3, occ_refresh_every=2)
ts.step_id = 1024
for i in range(4):
    loss, n = ts(rays, pixels, render_bkgd=torch.zeros(3, device=dev), refresh_occupancy=True, next_rays=lambda k: rays)
torch.cuda.synchronize()
print("sanitizer workload ok", float(loss), n, ts._premarch.taken)
PY
for tool in memcheck racecheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_train.py > gpurun_out/sanitize_train_$tool.log 2>&1
  tail -4 gpurun_out/sanitize_train_$tool.log
done
