"""Data parallelism over rays for the CNC path: one process per GPU (torchrun), replicas of the tables / MLPs /
occupancy grid, each rank its own ray shard, ONE exchange per step -- a bucketed all-reduce of the gradients
(dominated by the 161 MB hash-table gradient at the product layout) -- plus an 8-byte all-reduce of the sample
count that drives the adaptive ray budget (train_CNC_nerf_synthetic.py:340-344).  The reference has no distributed
code; this is the B200 equivalent described in SURVEY 8(e).  Works on NCCL (GPU) and gloo (CPU tests).

`ShardedTableAdam` is the exchange of the latent tables: the rows of every table are split over the ranks, the gradient is
REDUCE-SCATTERED (each rank receives the average of its rows only: half the traffic of an all-reduce), Adam runs on the
owned rows (1/N of the 40 M latents, one pass that also emits the rows' two STE bit planes), and what comes back is an
ALL-GATHER of the bit planes -- 2 bits per latent, 10 MB instead of 161 MB -- because the field reads a latent only through
sign(p) (forward) and [|p| <= 1] (backward).  Rows a rank does not own hold a stand-in with the same two bits.
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
        self.peer = None
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

    def use_peer_memory(self, signals, slot: int, control_group=None) -> None:
        """average over NVLink peer memory instead of NCCL (csrc/peer.cu): the gradients are packed into a buffer every rank
        maps, the ranks meet at `signals.barrier(slot)`, and each rank sums all N buffers itself (rank order: every rank gets
        the same bits).  For the small parameters (MLPs, context models: < 1 MB) the cost of an all-reduce is its latency;
        this one has no ring and no spinning NCCL kernel beside the weight-gradient GEMMs.  Collective call."""
        from . import peer as P

        n = sum(p.numel() for p in self.params)
        n = (n + 3) // 4 * 4
        dev = self.params[0].device
        mem = P.PeerMemory(2 * 4 * n, group=self.group, device=dev, control_group=control_group)   # two copies: see reduce()
        self.peer = {"P": P, "sig": signals, "slot": slot, "mem": mem, "n": n, "turn": 0,
                     "mine": [mem.tensor(4 * n * k, n) for k in (0, 1)], "src": [mem.pointer_array(4 * n * k) for k in (0, 1)],
                     "out": torch.empty(n, device=dev)}
        off, views = 0, [[], []]
        for p in self.params:
            for k in (0, 1):
                views[k].append(self.peer["mine"][k][off:off + p.numel()].view_as(p))
            off += p.numel()
        self.peer["views"] = views
        self.peer["out_views"] = []
        off = 0
        for p in self.params:
            self.peer["out_views"].append(self.peer["out"][off:off + p.numel()].view_as(p))
            off += p.numel()

    @torch.no_grad()
    def _reduce_peer(self) -> None:
        pr = self.peer
        k = pr["turn"]
        pr["turn"] ^= 1        # (a rank that skips its update may be a whole step ahead of the slowest reader of this buffer)
        world = dist.get_world_size(self.group)
        have = [p.grad is not None for p in self.params]
        if not all(have):
            pr["mine"][k].zero_()
        if any(have):
            torch._foreach_copy_([v for v, h in zip(pr["views"][k], have) if h], [p.grad for p, h in zip(self.params, have) if h])
        pr["sig"].barrier(pr["slot"])
        pr["P"].reduce_rows(pr["src"][k], world, 0, pr["n"], 1.0 / world, pr["out"])
        for p, h in zip(self.params, have):
            if not h:
                p.grad = torch.empty_like(p)
        torch._foreach_copy_([p.grad for p in self.params], pr["out_views"])

    @torch.no_grad()
    def reduce(self) -> None:
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return
        if self.peer is not None:
            return self._reduce_peer()
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
                pending.append((dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.group, async_op=True), flat, None))
            else:
                flat = torch.cat([p.grad.reshape(-1) for p in ps])
                pending.append((dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.group, async_op=True), flat, ps))
        for work, flat, ps in pending:
            work.wait()
            if ps is not None:
                off = 0
                for p in ps:
                    n = p.numel()
                    p.grad.copy_(flat[off:off + n].view_as(p.grad))
                    off += n


class _EventWork:
    """a CUDA event with the `.wait()` of a torch.distributed work handle: the current stream waits for it"""

    def __init__(self, event):
        self.event = event

    def wait(self):
        torch.cuda.current_stream().wait_event(self.event)


class ShardedTableAdam:
    """Adam over the latent hash tables of `encoders` (GridEncoder modules, F = 8 product tables), rows split over the
    ranks of `group`.  One `step()` per training iteration, after `backward()`:

        reduce-scatter(AVG) of every table gradient        -> this rank's rows          [161 MB in, 161/N MB out]
        Adam + bit-plane refresh on the owned rows          (train_ops.adam_planes, one pass)
        all-gather of the sign and window planes            [2 x 5 MB]
        stand-in latents for the rows owned elsewhere       (train_ops.surrogate_fill, one write pass)

    `exchange` picks who moves the bytes.  "nccl": the two collectives above.  "peer" (what "auto" takes when all ranks sit
    on one NVLink box): the gradients are written into buffers every rank maps (peer.PeerMemory); after a device-side
    barrier each rank LOADS its rows straight from all N buffers and averages them (`cnc_peer_reduce`: every byte crosses
    NVLink once, no ring, no NCCL kernel holding SMs), and after Adam STORES its plane words into every peer's planes
    (`cnc_peer_push`) -- the same result up to the order of the N-term sum, which here is the rank order.

    A table of n elements is cut at multiples of 32*world elements (whole words of the bit planes per rank); the < 32*world
    elements behind the last cut are replicated: their gradient is all-reduced and every rank updates them identically.
    The sign plane is handed to the encoder's sign cache, so the forward of the next step starts without a repack pass.
    world == 1 is the same code without the collectives.  `sync_params()` all-gathers the true fp32 rows (checkpoints)."""

    def __init__(self, encoders, lr: float = 6e-3, betas=(0.9, 0.999), eps: float = 1e-15, weight_decay: float = 0.0, group=None,
                 ste_window: bool = False, exchange: str = "auto", peer_control_group=None):
        from .train_ops import planes_pack

        self.encoders = list(encoders)
        self.lr, self.betas, self.eps, self.weight_decay, self.group = lr, betas, eps, weight_decay, group
        self.ste_window = ste_window    # the table gradients arrive without the STE window mask: apply it in the Adam pass
        on = dist.is_initialized() and group is not False         # group=False: this process alone, whatever is initialised
        self.world = dist.get_world_size(group) if on else 1
        self.rank = dist.get_rank(group) if on else 0
        self.step_id = 0
        self.tables = []
        for enc in self.encoders:
            p = enc.params
            n = p.numel()
            if n % 32:
                raise ValueError("ShardedTableAdam needs tables with numel % 32 == 0 (rows are multiples of 8: any F >= 4)")
            gran = 32 * self.world
            n_main = n // gran * gran
            S = n_main // self.world
            lo, hi = self.rank * S, (self.rank + 1) * S
            t = {"p": p, "n": n, "n_main": n_main, "S": S, "lo": lo, "hi": hi,
                 "m": torch.zeros(S, device=p.device), "v": torch.zeros(S, device=p.device),
                 "tm": torch.zeros(n - n_main, device=p.device), "tv": torch.zeros(n - n_main, device=p.device)}
            t["sign"], t["mask"] = planes_pack(p.detach().contiguous().view(-1))
            self.tables.append(t)
        if exchange not in ("auto", "nccl", "peer"):
            raise ValueError("exchange must be 'auto', 'nccl' or 'peer'")
        self.peer = None
        cuda = all(t["p"].is_cuda for t in self.tables)
        if self.world > 1 and cuda and exchange != "nccl":
            from . import peer as P

            if exchange == "peer" or P.peer_capable(group):
                self._setup_peer(P, peer_control_group)
        elif exchange == "peer" and self.world > 1:
            raise RuntimeError("exchange='peer' needs CUDA tables")

    # ---- peer-memory exchange (csrc/peer.cu) -------------------------------------------------------------------------
    def _setup_peer(self, P, control_group) -> None:
        dev = self.tables[0]["p"].device
        sig = P.PeerSignals(group=self.group, device=dev, control_group=control_group)
        # gradient arena: two copies of every table (a step writes one while a slow peer may still read the other)
        g_off, off = [], 0
        for t in self.tables:
            g_off.append(off)
            off += (4 * t["n"] + 255) // 256 * 256
        g_bytes = off
        gmem = P.PeerMemory(2 * g_bytes, group=self.group, device=dev, control_group=control_group)
        # plane arena: sign and window plane of every table
        p_off, off = [], 0
        for t in self.tables:
            nb = (t["n"] // 8 + 255) // 256 * 256
            p_off.append((off, off + nb))
            off += 2 * nb
        pmem = P.PeerMemory(off, group=self.group, device=dev, control_group=control_group)
        for k, t in enumerate(self.tables):
            for name, o in zip(("sign", "mask"), p_off[k]):
                plane = pmem.tensor(o, t["n"] // 8, torch.uint8)
                plane.copy_(t[name])
                t[name] = plane
            t["gbuf"] = [gmem.tensor(par * g_bytes + g_off[k], t["n"]).view_as(t["p"]) for par in (0, 1)]
            t["gsrc"] = [gmem.pointer_array(par * g_bytes + g_off[k]) for par in (0, 1)]
            t["gs"] = torch.empty(t["S"], device=dev)
            t["gt"] = torch.empty(t["n"] - t["n_main"], device=dev)
            t["push"] = [(o // 4 + t["lo"] // 32, t["S"] // 32) for o in p_off[k]]
        self.peer = {"P": P, "sig": sig, "gmem": gmem, "pmem": pmem, "stream": torch.cuda.Stream(dev), "parity": 0}
        torch.cuda.synchronize(dev)
        sig.barrier(15)                    # nobody starts reading before everybody's arenas hold their initial contents
        torch.cuda.synchronize(dev)

    def grad_buffer(self, k: int):
        """where this step's gradient of table k should be accumulated (peer exchange: the buffer the other ranks read;
        otherwise None = allocate as usual).  The caller zeroes it."""
        return None if self.peer is None else self.tables[k]["gbuf"][self.peer["parity"]]

    def _peer_reduce(self, k: int, grad: torch.Tensor):
        """table k's gradient of this step is complete in `grad`: meet the other ranks, average the owned rows"""
        pr, t = self.peer, self.tables[k]
        P, par = pr["P"], pr["parity"]
        buf = t["gbuf"][par]
        if grad is None:
            buf.zero_()
        elif grad.data_ptr() != buf.data_ptr():
            buf.copy_(grad.view_as(buf))
        ready = torch.cuda.Event()
        ready.record()
        with torch.cuda.stream(pr["stream"]):
            pr["stream"].wait_event(ready)
            pr["sig"].barrier(k)
            P.reduce_rows(t["gsrc"][par], self.world, t["lo"], t["S"], 1.0 / self.world, t["gs"])
            if t["n"] > t["n_main"]:
                P.reduce_rows(t["gsrc"][par], self.world, t["n_main"], t["n"] - t["n_main"], 1.0 / self.world, t["gt"])
            done = torch.cuda.Event()
            done.record()
        return buf.view(-1), t["gs"], t["gt"], _EventWork(done)

    def comm_bytes_per_step(self) -> int:
        """bytes this rank hands to the collectives per step (reduce-scatter input + its share of the all-gathers)"""
        if self.world == 1:
            return 0
        return sum(4 * t["n_main"] + 4 * (t["n"] - t["n_main"]) + 2 * t["S"] // 8 for t in self.tables)

    def link_bytes_per_step(self) -> int:
        """bytes that cross this rank's NVLink port per step in the peer exchange: rows loaded from the N - 1 peers plus the
        plane words stored into them"""
        W = self.world
        return sum((W - 1) * 4 * (t["S"] + t["n"] - t["n_main"]) + (W - 1) * 2 * t["S"] // 8 for t in self.tables) if W > 1 else 0

    def _batched(self):
        """context that turns the collectives issued inside it into ONE NCCL group launch (gloo: issued one by one)"""
        import contextlib

        if self.world > 1 and dist.get_backend(self.group) == "nccl":
            return dist._coalescing_manager(group=self.group, async_ops=True)
        return contextlib.nullcontext()

    @torch.no_grad()
    def contribute(self, k: int, grad: torch.Tensor) -> bool:
        """hand over the COMPLETE gradient of table k as soon as it exists (called from inside backward): its
        reduce-scatter starts at once and overlaps whatever backward still has to do.  Only valid when nothing else
        contributes to that table's gradient in this step (no rate term); returns True = consumed (leave `.grad` empty)."""
        if self.world == 1:
            return False
        t = self.tables[k]
        self._early = getattr(self, "_early", {})
        if self.peer is not None:
            self._early[k] = self._peer_reduce(k, grad)
            return True
        g = grad.contiguous().view(-1)
        gs = torch.empty(t["S"], device=g.device)
        w = dist.reduce_scatter_tensor(gs, g[:t["n_main"]], op=dist.ReduceOp.AVG, group=self.group, async_op=True)
        self._early[k] = (g, gs, None, w)
        return True

    @torch.no_grad()
    def exchange(self):
        """launch the gradient exchange (asynchronous): reduce-scatter of every table's rows (those not handed over early
        by `contribute`), one small all-reduce of the replicated tails.  Returns what `apply` needs."""
        W, works, parts = self.world, [], []
        early, self._early = getattr(self, "_early", {}), {}
        if self.peer is not None:
            for k, t in enumerate(self.tables):
                g, gs, gt, w = early[k] if k in early else self._peer_reduce(k, t["p"].grad)
                works.append(w)
                parts.append((g, gs, gt))
            self.peer["parity"] ^= 1
            return works, parts
        tails, late = [], []
        for k, t in enumerate(self.tables):
            if k in early:
                g, gs, _, w = early[k]
                works.append(w)
            else:
                p = t["p"]
                g = (p.grad if p.grad is not None else torch.zeros_like(p)).contiguous().view(-1)
                gs = torch.empty(t["S"], device=g.device) if W > 1 else g[:t["n_main"]]
                late.append((t, g, gs))
            parts.append([g, gs, None])
            tails.append(g[t["n_main"]:])
        tail = torch.cat(tails) if sum(x.numel() for x in tails) else None
        if W > 1:
            if late:
                with self._batched() as cm:
                    for t, g, gs in late:
                        w = dist.reduce_scatter_tensor(gs, g[:t["n_main"]], op=dist.ReduceOp.AVG, group=self.group, async_op=True)
                        if cm is None:
                            works.append(w)
                if cm is not None:
                    works.append(cm)
            if tail is not None:
                works.append(dist.all_reduce(tail, op=dist.ReduceOp.AVG, group=self.group, async_op=True))
        off = 0
        for part, t in zip(parts, self.tables):     # every table's replicated tail as a view of the reduced buffer
            nt = t["n"] - t["n_main"]
            part[2] = tail[off:off + nt] if nt else None
            off += nt
        return works, parts

    @torch.no_grad()
    def apply(self, exchanged) -> None:
        """Adam on the owned rows (+ the replicated tails), bit planes to everybody, stand-ins for the rows owned elsewhere"""
        from .train_ops import adam_planes, surrogate_fill

        works, parts = exchanged[0], exchanged[1]
        for w in works:
            w.wait()
        self.step_id += 1
        W = self.world
        kw = dict(step=self.step_id, lr=self.lr, betas=self.betas, eps=self.eps, weight_decay=self.weight_decay,
                  ste_window=self.ste_window)
        sends = []
        for t, (g, gs, gt) in zip(self.tables, parts):
            flat = t["p"].detach().view(-1)
            lo, hi, nm = t["lo"], t["hi"], t["n_main"]
            adam_planes(flat[lo:hi], gs, t["m"], t["v"], sign=t["sign"][lo // 8:hi // 8], mask=t["mask"][lo // 8:hi // 8], **kw)
            if t["n"] > nm:
                adam_planes(flat[nm:], gt.contiguous(), t["tm"], t["tv"], sign=t["sign"][nm // 8:], mask=t["mask"][nm // 8:], **kw)
            if W > 1 and self.peer is None:
                for plane in (t["sign"], t["mask"]):
                    sends.append((plane[:nm // 8], plane[lo // 8:hi // 8].clone()))
        if self.peer is not None:
            pr = self.peer
            pr["P"].push_words(pr["pmem"], [seg for t in self.tables for seg in t["push"]])
            pr["sig"].barrier(8)               # everybody's words have arrived here, and mine everywhere
        elif W > 1:
            works = []
            with self._batched() as cm:
                for out, mine in sends:
                    w = dist.all_gather_into_tensor(out, mine, group=self.group, async_op=True)
                    if cm is None:
                        works.append(w)
            if cm is not None:
                works.append(cm)
            for w in works:
                w.wait()
        for t, enc in zip(self.tables, self.encoders):
            p = t["p"]
            if W > 1:
                keep = p.detach().view(-1)[t["n_main"]:].clone()
                surrogate_fill(p.detach().view(-1), t["sign"], t["mask"], t["lo"], t["hi"])
                p.detach().view(-1)[t["n_main"]:] = keep
            torch.autograd.graph.increment_version(p)         # the kernels wrote through raw pointers
            cache = getattr(enc, "_sign_cache", None)
            if cache is not None and p.is_cuda:               # the next forward gathers from this plane: no repack pass
                cache.publish(p, t["sign"])

    def step(self) -> None:
        self.apply(self.exchange())

    @torch.no_grad()
    def sync_params(self) -> None:
        """make `.params` the true fp32 latents on every rank (before a checkpoint; not needed for training or encoding,
        which read signs only)"""
        if self.world == 1:
            return
        for t in self.tables:
            flat = t["p"].detach().view(-1)
            mine = flat[t["lo"]:t["hi"]].clone()
            dist.all_gather_into_tensor(flat[:t["n_main"]], mine, group=self.group)
            torch.autograd.graph.increment_version(t["p"])


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
