"""Data parallelism over rays for the CNC path: one process per GPU (torchrun), replicas of the tables / MLPs /
occupancy grid, each rank its own ray shard, ONE exchange per step -- a bucketed all-reduce of the gradients
(dominated by the 161 MB hash-table gradient at the product layout) -- plus an 8-byte all-reduce of the sample
count that drives the adaptive ray budget (train_CNC_nerf_synthetic.py:340-344).  The reference has no distributed
code; this is the B200 equivalent described in SURVEY 8(e).  Works on NCCL (GPU) and gloo (CPU tests).
"""
from __future__ import annotations

from typing import Iterable, List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """contiguous, balanced [lo, hi) of n items for `rank` (sizes differ by at most one)"""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_rays(rays, rank: int, world: int):
    """the rank's slice of every field of a Rays namedtuple (or of a plain tensor)"""
    if isinstance(rays, torch.Tensor):
        lo, hi = shard_range(rays.shape[0], rank, world)
        return rays[lo:hi]
    lo, hi = shard_range(rays[0].shape[0], rank, world)
    return type(rays)(*(None if r is None else r[lo:hi] for r in rays))


class GradAllReducer:
    """Bucketed average of `.grad` over the process group.

    Parameters are packed into flat buckets of at most `bucket_bytes` (a parameter larger than that gets its own
    bucket and is reduced in place, without a copy: the hash tables).  `reduce()` launches every bucket's
    all-reduce asynchronously, then waits and scatters the averages back; parameters without a gradient
    contribute zeros so that all ranks issue identical collectives.
    """

    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_bytes: int = 32 << 20, group=None):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.group = group
        self.buckets: List[List[int]] = []
        cur, cur_bytes = [], 0
        for i, p in enumerate(self.params):
            nb = p.numel() * p.element_size()
            if nb >= bucket_bytes:
                self.buckets.append([i])
                continue
            if cur and cur_bytes + nb > bucket_bytes:
                self.buckets.append(cur)
                cur, cur_bytes = [], 0
            cur.append(i)
            cur_bytes += nb
        if cur:
            self.buckets.append(cur)

    def bytes_per_step(self) -> int:
        return sum(p.numel() * p.element_size() for p in self.params)

    @torch.no_grad()
    def reduce(self) -> None:
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return
        world = dist.get_world_size(self.group)
        pending = []
        for b in self.buckets:
            ps = [self.params[i] for i in b]
            for p in ps:
                if p.grad is None:
                    p.grad = torch.zeros_like(p)
            if len(ps) == 1:
                flat = ps[0].grad.view(-1) if ps[0].grad.is_contiguous() else None
                if flat is None:
                    ps[0].grad = ps[0].grad.contiguous()
                    flat = ps[0].grad.view(-1)
                pending.append((dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True), flat, None))
            else:
                flat = torch.cat([p.grad.reshape(-1) for p in ps])
                pending.append((dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True), flat, ps))
        for work, flat, ps in pending:
            work.wait()
            flat.div_(world)
            if ps is not None:
                off = 0
                for p in ps:
                    n = p.numel()
                    p.grad.copy_(flat[off:off + n].view_as(p.grad))
                    off += n


def allreduce_scalar(value: float, device, op=None, group=None) -> float:
    """sum (default) of a python number over the ranks, e.g. the number of rendered samples of the step"""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=op or dist.ReduceOp.SUM, group=group)
    return float(t.item())


def broadcast_module_buffers(module: torch.nn.Module, names: Sequence[str], src: int = 0, group=None) -> None:
    """keep replicated state coherent (e.g. OccGridEstimator.occs / .binaries after a refresh on rank 0)"""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for n in names:
        t = getattr(module, n)
        if t.dtype == torch.bool:
            u = t.to(torch.uint8)
            dist.broadcast(u, src=src, group=group)
            t.copy_(u.bool())
        else:
            dist.broadcast(t, src=src, group=group)
