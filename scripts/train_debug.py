import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import torch
from test_gpu_train import _scene
from cnc_b200.trainer import TrainStep
from cnc_b200.render import render_image_with_occgrid
dev = torch.device("cuda:0")
bk = torch.zeros(3, device=dev)
for shard in (True, False):
    field, est, rays, pixels = _scene(dev)
    ts = TrainStep(field, est, lr=2e-3, shard_tables=shard)
    out = []
    for i in range(12):
        loss, n = ts(rays, pixels, render_bkgd=bk, refresh_occupancy=False)
        field.eval()
        with torch.no_grad():
            a = float(torch.nn.functional.mse_loss(render_image_with_occgrid(field, est, rays, render_step_size=5e-3, render_bkgd=bk)[0], pixels))
            mb = field.mlp_base
            for e in (mb.encoding_xyz, mb.encoding_xy, mb.encoding_xz, mb.encoding_yz): e.invalidate()
            field._blob_key = None
            b = float(torch.nn.functional.mse_loss(render_image_with_occgrid(field, est, rays, render_step_size=5e-3, render_bkgd=bk)[0], pixels))
        field.train()
        out.append((round(float(loss), 4), round(a, 4), round(b, 4), float(mb.encoding_xyz.params.abs().max()), float(mb.network[0].weight.abs().max())))
    print("shard", shard)
    for o in out: print("  ", o)
