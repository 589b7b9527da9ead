"""One training iteration of the CNC scripts as a reusable object (train_CNC_nerf_synthetic.py:302-366):
occupancy refresh -> occupancy/visibility sampling -> differentiable render -> photometric loss (+ lambda * rate)
-> backward -> [data parallel: gradient exchange] -> Adam.  It is the caller of the hot path, kept small: no datasets,
schedulers or logging (`lr` is an attribute: a scheduler sets it between steps).

Optimiser layout (train_CNC_nerf_synthetic.py:254-266): Adam(lr, eps=1e-15, weight_decay) over the field, Adam(lr,
eps=1e-15) over the context models.  Here the four latent tables (99.7 % of the parameters) go through
`dp.ShardedTableAdam` -- one fused pass that also refreshes the 1-bit planes the next forward reads; under data parallelism
a reduce-scatter in, two bit planes out -- and everything else (MLPs, context models, < 1 MB) through a bucketed
all-reduce and torch's fused Adam.

Data parallelism: every rank renders its own ray shard.  The decision to skip a step (the reference `continue`s when a
batch produced no sample, :337-338) is taken by all ranks together, and the occupancy grid is refreshed on rank 0's random
numbers only (broadcast), so replicas never diverge.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist
import torch.nn.functional as F

from .dp import GradAllReducer, ShardedTableAdam, allreduce_scalar, broadcast_module_buffers
from .render import Rays, render_image_with_occgrid


class TrainStep:
    def __init__(self, radiance_field, estimator, context_model=None, lmbda: float = 0.0, lr: float = 6e-3,
                 render_step_size: float = 5e-3, target_sample_batch_size: int = 1 << 18, bucket_bytes: int = 32 << 20,
                 weight_decay: float = 2e-6, occ_refresh_every: int = 16, shard_tables: bool = True, exchange: str = "auto",
                 data_parallel: bool = True, sync_initial_state: bool = True, overlap_rate_term: bool = False):
        self.field, self.estimator, self.cm, self.lmbda = radiance_field, estimator, context_model, lmbda
        if dist.is_initialized() and data_parallel and dist.get_world_size() > 1 and sync_initial_state:
            # replicas start from rank 0's parameters, occupancy and context models (whatever the ranks' RNG did before)
            with torch.no_grad():
                for m in (radiance_field, context_model):
                    if m is not None:
                        for p in m.parameters():
                            dist.broadcast(p.data, src=0)
            broadcast_module_buffers(estimator, ["occs", "binaries"], src=0)
            if hasattr(radiance_field, "invalidate_caches"):
                radiance_field.invalidate_caches()
        self.render_step_size, self.target, self.occ_every = render_step_size, target_sample_batch_size, occ_refresh_every
        self.lr = lr
        mb = radiance_field.mlp_base
        encs = [mb.encoding_xyz, mb.encoding_xy, mb.encoding_xz, mb.encoding_yz]
        self.sharded = shard_tables and all(e.params.numel() % 32 == 0 and e.ste_binary for e in encs)
        table_ids = {id(e.params) for e in encs} if self.sharded else set()
        field_rest = [p for p in radiance_field.parameters() if id(p) not in table_ids and p.requires_grad]
        ctx = [p for p in context_model.parameters() if p.requires_grad] if context_model is not None else []
        # the STE window of the render path's table gradient is applied inside the table optimizer's pass (idempotent for the
        # contributions of the rate term, which arrive already masked): four full-table elementwise kernels less per step
        # data_parallel=False: a single-process step even under torchrun (every rank trains its own replica; bench.py times it
        # beside the data-parallel step to state the scaling efficiency from one run)
        self.table_opt = ShardedTableAdam(encs, lr=lr, eps=1e-15, weight_decay=weight_decay, ste_window=True, exchange=exchange,
                                          group=None if data_parallel else False) if self.sharded else None
        groups = [{"params": field_rest, "weight_decay": weight_decay}]
        if ctx:
            groups.append({"params": ctx, "weight_decay": 0.0})
        self.optimizer = torch.optim.Adam(groups, lr=lr, eps=1e-15, fused=field_rest[0].is_cuda)
        self.reducer = GradAllReducer(field_rest + ctx, bucket_bytes=bucket_bytes)
        self.world = dist.get_world_size() if (dist.is_initialized() and data_parallel) else 1
        self.step_id = 0
        self._premarch = None
        self._rate_stream = None
        self.overlap_rate_term = overlap_rate_term
        if context_model is not None and hasattr(context_model, "set_data_parallel"):
            # the rate term is shared among the ranks: each samples 1/N of the entries and evaluates its share of the plane terms
            context_model.set_data_parallel(dist.get_rank() if self.world > 1 else 0, self.world)
        # the tables' exchange runs over NVLink peer memory: so do the small gradients and the skip vote -- no NCCL kernel
        # (a spinning all-reduce holds SM slots beside the persistent GEMM kernels of backward) is left in the step
        self._peer_sig = self.table_opt.peer["sig"] if (self.table_opt is not None and self.table_opt.peer is not None) else None
        if self._peer_sig is not None and self.reducer.params:
            self.reducer.use_peer_memory(self._peer_sig, slot=4)

    # ---- what the exchange moves (for the bench line / DESIGN.md)
    def comm_bytes_per_step(self) -> int:
        return (self.table_opt.comm_bytes_per_step() if self.table_opt is not None else 0) + \
            (self.reducer.bytes_per_step() if self.world > 1 else 0)

    def comm_description(self) -> str:
        if self.world == 1:
            return "single process: no exchange"
        if self.table_opt is None:
            return "bucketed all-reduce of every gradient"
        if self.table_opt.peer is not None:
            return ("latent tables over NVLink peer memory: device-side barrier, every rank loads and averages its 1/N of the rows "
                    "from all N gradient buffers (cnc_peer_reduce), Adam on them, stores its words of the sign and STE-window bit "
                    "planes into every peer (cnc_peer_push); MLP / context-model gradients: packed into a mapped buffer, summed by every "
                    "rank from all N (cnc_peer_reduce); skip vote: minimum of the sample counts through the signal pads (cnc_peer_min); "
                    "no NCCL kernel in the step" + ("; rate term: sampled entries and plane terms split over the ranks" if self.cm is not None else ""))
        return ("latent tables: reduce-scatter(avg) of the gradient by rows, Adam on the owned 1/N, all-gather of the sign and "
                "STE-window bit planes; MLP / context-model gradients: one bucketed all-reduce; sample count: 8-byte all-reduce")

    def _everyone_has_samples(self, n_samples: int, device) -> bool:
        if self.world == 1:
            return n_samples > 0
        return allreduce_scalar(float(n_samples), device, op=dist.ReduceOp.MIN) > 0

    def _vote_async(self, n_samples: int, device):
        """start the all-ranks minimum of the sample count; `_vote_result` reads it later without stalling the launch
        queue: the result is copied to pinned memory on a side stream, and only that copy's event is waited for"""
        if self.world == 1 or device.type != "cuda":
            return n_samples
        if getattr(self, "_vote_stream", None) is None:
            self._vote_stream = torch.cuda.Stream(device)
            self._vote_host = torch.zeros(1, dtype=torch.int64).pin_memory()
            self._vote_in = torch.zeros(1, dtype=torch.int64).pin_memory()
            self._vote_dev = torch.zeros(2, dtype=torch.int32, device=device)       # (one int64 for the copy below)
        if self._peer_sig is not None:
            # one warp: the count goes into every rank's signal pad, the minimum comes out of the own one
            with torch.cuda.stream(self._vote_stream):
                self._peer_sig.minimum(10, n_samples, self._vote_dev)
                self._vote_host.copy_(self._vote_dev.view(torch.int64), non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self._vote_stream)
            return ev
        self._vote_in[0] = n_samples            # (the copy of the previous step left this buffer long ago: its result was read)
        flag = self._vote_in.to(device, non_blocking=True)
        self._vote_stream.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(self._vote_stream):
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            self._vote_host.copy_(flag, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._vote_stream)
        flag.record_stream(self._vote_stream)
        return ev

    def _vote_result(self, vote) -> bool:
        if isinstance(vote, int):
            return vote > 0
        vote.synchronize()
        return int(self._vote_host[0]) > 0

    def _rate_term(self):
        mb = self.field.mlp_base
        bpp, _ = self.cm.forward_binary_vxl_mixPg_3D2D(mb.encoding_xyz, mb.encoding_xy, mb.encoding_xz, mb.encoding_yz,
                                                       self.estimator.binaries, step=self.step_id, mb_as_tensor=True)
        return bpp

    def _pick_next(self, next_rays, n_samples: int, refresh_occupancy: bool):
        """the batch whose occupancy march may run ahead (nerfacc.Premarch), and the point of the stream from which it may.
        Not before a step that refreshes the grid: that step marches in line.  The march itself is issued at the END of the
        step (`_march_ahead`), i.e. after the rate term has drawn its random windows, so the order of random draws -- jitter
        of batch t, windows of step t, jitter of batch t + 1 -- and hence every sample stays what it is without the
        look-ahead (with `overlap_rate_term` the rate term runs first and the order is its own)."""
        if next_rays is None:
            return None
        if refresh_occupancy and (self.step_id + 1) % self.occ_every == 0:
            return None
        if callable(next_rays):
            next_rays = next_rays(n_samples)
        if next_rays is None or not next_rays.origins.is_cuda:
            return None
        ready = torch.cuda.Event()
        ready.record()
        return next_rays, ready

    def _march_ahead(self, picked) -> None:
        """issue the march of the next batch on the side stream.  Called once this step's backward and update are queued: the
        host is then far ahead of the device, and the march (which depends on the rays and the grid only) runs beside the
        weight-gradient and optimizer kernels instead of in front of the next step."""
        if picked is None:
            return
        next_rays, ready = picked
        if self._premarch is None:
            from .nerfacc import Premarch

            self._premarch = Premarch(next_rays.origins.device)
        self._premarch.issue(self.estimator, next_rays.origins.reshape(-1, 3), next_rays.viewdirs.reshape(-1, 3),
                             render_step_size=self.render_step_size, stratified=True, after=ready)

    def __call__(self, rays: Rays, pixels: torch.Tensor, render_bkgd: Optional[torch.Tensor] = None, refresh_occupancy: bool = True,
                 next_rays=None):
        """returns (loss value tensor, number of rendered samples on this rank).  `next_rays` (optional): the batch of the
        NEXT call, or a function `n_samples -> Rays` that draws it once this step's sample count is known (the reference sizes
        the next batch from it, train...:340-344); its occupancy march then overlaps this step (see `_march_ahead`)."""
        self.field.train()
        self.estimator.train()
        if refresh_occupancy:   # train...:314-321
            self.estimator.update_every_n_steps(step=self.step_id, occ_thre=1e-2, n=self.occ_every,
                                                occ_eval_fn=lambda x: self.field.query_density(x) * self.render_step_size)
            if self.world > 1 and self.step_id % self.occ_every == 0:
                # each rank drew its own random cell samples: rank 0's grid is everybody's grid
                broadcast_module_buffers(self.estimator, ["occs", "binaries"], src=0)
        # overlap_rate_term (off by default): the rate term does not depend on the rays, so its forward can be issued first, on
        # a side stream, and its small kernels (and, in backward, the nodes autograd runs on that same stream) share the device
        # with the render path's few large ones.  Measured: 13.8 instead of 14.0 ms per step -- the step is bound by the rate
        # term's own chain of kernels, not by the 3.7 ms beside it -- at the price of cross-stream gradient accumulation.
        bpp = None
        with_rate = self.cm is not None and self.lmbda > 0
        if with_rate and self.overlap_rate_term and pixels.is_cuda:
            cur = torch.cuda.current_stream(pixels.device)
            if self._rate_stream is None:
                self._rate_stream = torch.cuda.Stream(pixels.device)
            self._rate_stream.wait_stream(cur)
            with torch.cuda.stream(self._rate_stream):
                bpp = self._rate_term()
        rgb, acc, depth, n_samples = render_image_with_occgrid(self.field, self.estimator, rays, render_step_size=self.render_step_size,
                                                               render_bkgd=render_bkgd, premarch=self._premarch)
        if self._premarch is not None:
            self._premarch.drop()                 # (a march issued for other rays / another grid is not kept)
        picked = self._pick_next(next_rays, n_samples, refresh_occupancy)
        # train...:337-338: a batch without samples skips the step.  Under data parallelism all ranks decide together; the
        # vote travels while this rank goes on (a rank without samples contributes zero gradients to the collectives, which
        # every rank issues in the same order either way) and is read just before the update is applied.
        vote = self._vote_async(n_samples, pixels.device)
        if bpp is not None:
            torch.cuda.current_stream(pixels.device).wait_stream(self._rate_stream)
            bpp.record_stream(torch.cuda.current_stream(pixels.device))
        if self.world == 1 and n_samples == 0:
            self._march_ahead(picked)
            self.step_id += 1
            return torch.zeros((), device=pixels.device), n_samples
        loss = F.mse_loss(rgb, pixels)   # train...:346
        if with_rate:
            if bpp is None:
                bpp = self._rate_term()
            loss = loss + self.lmbda * bpp
        self.optimizer.zero_grad(set_to_none=True)
        if self.table_opt is not None:
            for t in self.table_opt.tables:
                t["p"].grad = None
        # without a rate term the render path is the only source of table gradients: their exchange starts inside backward
        early = self.table_opt is not None and self.world > 1 and not (self.cm is not None and self.lmbda > 0)
        self.field._table_grad_sink = self.table_opt.contribute if early else None
        self.field._defer_ste = self.table_opt is not None
        # (peer exchange: K2 accumulates straight into the buffer the other ranks will read)
        self.field._table_grad_buffer = self.table_opt.grad_buffer if early else None
        if loss.requires_grad:           # (a data-parallel rank whose batch produced no sample still joins the collectives)
            loss.backward()
        self.field._table_grad_sink = None
        self.field._table_grad_buffer = None
        self.field._defer_ste = False
        for g in self.optimizer.param_groups:
            g["lr"] = self.lr
        exchanged = None
        if self.table_opt is not None:
            self.table_opt.lr = self.lr
            exchanged = self.table_opt.exchange()   # tables: reduce-scatter of the rows (asynchronous)
        if self.world > 1:
            self.reducer.reduce()                   # MLPs + context models: one small bucketed average over the ranks
        if self._vote_result(vote):
            if exchanged is not None:
                self.table_opt.apply(exchanged)     # Adam on the owned rows -> bit planes all-gather -> stand-ins
            self.optimizer.step()
        elif exchanged is not None:
            for w in exchanged[0]:
                w.wait()
        self._march_ahead(picked)
        self.step_id += 1
        return loss.detach(), n_samples

    def adapt_num_rays(self, num_rays: int, n_samples_local: int, device) -> int:
        """train...:340-344 with the global sample count (all ranks take the same decision)"""
        total = allreduce_scalar(float(n_samples_local), device)
        return int(num_rays * (self.target * self.world / max(total, 1.0)))

    def sync_params(self) -> None:
        """true fp32 latents on every rank (a data-parallel rank otherwise keeps stand-ins with the right sign and STE
        window for the rows it does not own)"""
        if self.table_opt is not None:
            self.table_opt.sync_params()
