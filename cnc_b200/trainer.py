"""One training iteration of the CNC scripts as a reusable object (train_CNC_nerf_synthetic.py:302-366):
occupancy refresh -> occupancy/visibility sampling -> differentiable render -> photometric loss (+ lambda * rate)
-> backward -> [data parallel: bucketed gradient all-reduce] -> Adam.  It is the caller of the hot path, kept small:
no datasets, schedulers or logging."""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F

from .dp import GradAllReducer, allreduce_scalar
from .render import Rays, render_image_with_occgrid


class TrainStep:
    def __init__(self, radiance_field, estimator, context_model=None, lmbda: float = 0.0, lr: float = 6e-3,
                 render_step_size: float = 5e-3, target_sample_batch_size: int = 1 << 18, bucket_bytes: int = 32 << 20):
        self.field, self.estimator, self.cm, self.lmbda = radiance_field, estimator, context_model, lmbda
        self.render_step_size, self.target = render_step_size, target_sample_batch_size
        params = list(radiance_field.parameters()) + (list(context_model.parameters()) if context_model is not None else [])
        self.optimizer = torch.optim.Adam(params, lr=lr, eps=1e-15, fused=params[0].is_cuda)   # train...:254-266; one fused pass over the 40 M latents
        self.reducer = GradAllReducer(params, bucket_bytes=bucket_bytes)
        self.step_id = 0

    def __call__(self, rays: Rays, pixels: torch.Tensor, render_bkgd: Optional[torch.Tensor] = None, refresh_occupancy: bool = True):
        """returns (loss value tensor, number of rendered samples on this rank)"""
        self.field.train()
        self.estimator.train()
        if refresh_occupancy:   # train...:314-321
            self.estimator.update_every_n_steps(step=self.step_id, occ_thre=1e-2,
                                                occ_eval_fn=lambda x: self.field.query_density(x) * self.render_step_size)
        rgb, acc, depth, n_samples = render_image_with_occgrid(self.field, self.estimator, rays, render_step_size=self.render_step_size,
                                                               render_bkgd=render_bkgd)
        if n_samples == 0:      # train...:337-338
            self.step_id += 1
            self.reducer.reduce()
            return torch.zeros((), device=pixels.device), 0
        loss = F.mse_loss(rgb, pixels)   # train...:346
        if self.cm is not None and self.lmbda > 0:
            mb = self.field.mlp_base
            bpp, _ = self.cm.forward_binary_vxl_mixPg_3D2D(mb.encoding_xyz, mb.encoding_xy, mb.encoding_xz, mb.encoding_yz,
                                                           self.estimator.binaries, step=self.step_id)
            loss = loss + self.lmbda * bpp
        self.optimizer.zero_grad(set_to_none=True)
        loss.backward()
        self.reducer.reduce()            # the one exchange of the data-parallel step
        self.optimizer.step()
        self.step_id += 1
        return loss.detach(), n_samples

    def adapt_num_rays(self, num_rays: int, n_samples_local: int, device) -> int:
        """train...:340-344 with the global sample count (all ranks take the same decision)"""
        total = allreduce_scalar(float(n_samples_local), device)
        world = torch.distributed.get_world_size() if torch.distributed.is_initialized() else 1
        return int(num_rays * (self.target * world / max(total, 1.0)))
