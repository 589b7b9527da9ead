"""CPU suite (gloo, world_size 2): the data-parallel host logic of cnc_b200.dp -- ray sharding, the bucketed
gradient all-reduce, scalar reduction and buffer broadcast.  No CUDA involved."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cnc_b200.dp import GradAllReducer, allreduce_scalar, broadcast_module_buffers, shard_rays
    from cnc_b200.render import Rays

    torch.manual_seed(0)  # identical replicas
    model = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.Linear(5, 3))
    table = torch.nn.Parameter(torch.zeros(1000, 8))          # "hash table": larger than the bucket -> own bucket
    frozen = torch.nn.Parameter(torch.zeros(4), requires_grad=False)
    unused = torch.nn.Parameter(torch.zeros(6))                # never receives a gradient on any rank
    params = list(model.parameters()) + [table, frozen, unused]
    red = GradAllReducer(params, bucket_bytes=256)
    assert [len(b) for b in red.buckets if 4 in [i for i in b]] or True
    # each rank: loss on its own shard of the "rays"
    g = torch.Generator().manual_seed(123)
    rays = Rays(torch.randn(11, 7, generator=g), torch.randn(11, 3, generator=g))
    mine = shard_rays(rays, rank, world)
    rows = torch.arange(11)[slice(*__import__("cnc_b200.dp", fromlist=["shard_range"]).shard_range(11, rank, world))]
    loss = (model(mine.origins) * mine.viewdirs).sum() + (table[rows * 3] * (rank + 1.0)).sum()
    loss.backward()
    red.reduce()
    n = allreduce_scalar(float(mine.origins.shape[0]), "cpu")
    buf = torch.nn.Module()
    buf.register_buffer("binaries", torch.full((4,), rank == 0))
    buf.register_buffer("occs", torch.full((4,), float(rank + 1)))
    broadcast_module_buffers(buf, ["binaries", "occs"], src=0)
    q.put((rank, [p.grad.numpy().copy() if p.grad is not None else None for p in params], n, mine.origins.shape[0],
           buf.binaries.numpy().copy(), buf.occs.numpy().copy(), [list(b) for b in red.buckets]))
    dist.barrier()
    dist.destroy_process_group()


def test_bucketed_grad_allreduce_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    T = lambda l: [None if t is None else torch.from_numpy(t) for t in l]
    (r0, g0, n0, c0, b0, o0, buckets), (r1, g1, n1, c1, b1, o1, _) = out
    g0, g1, b0, b1, o1 = T(g0), T(g1), torch.from_numpy(b0), torch.from_numpy(b1), torch.from_numpy(o1)
    assert c0 + c1 == 11 and abs(c0 - c1) <= 1 and n0 == n1 == 11.0
    # single-process reference: mean over ranks of the per-rank gradients
    from cnc_b200.dp import shard_range
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.Linear(5, 3))
    table = torch.nn.Parameter(torch.zeros(1000, 8))
    g = torch.Generator().manual_seed(123)
    ro, rd = torch.randn(11, 7, generator=g), torch.randn(11, 3, generator=g)
    total = 0
    for rank in range(2):
        lo, hi = shard_range(11, rank, 2)
        total = total + ((model(ro[lo:hi]) * rd[lo:hi]).sum() + (table[torch.arange(lo, hi) * 3] * (rank + 1.0)).sum()) / 2
    total.backward()
    want = [p.grad for p in model.parameters()] + [table.grad]
    for a, b, w in zip(g0[:5], g1[:5], want):
        torch.testing.assert_close(a, b, rtol=0, atol=0)          # replicas end up identical
        torch.testing.assert_close(a, w, rtol=1e-6, atol=1e-7)
    assert g0[5] is None                                           # frozen parameter untouched
    assert g0[6] is not None and (g0[6] == 0).all()                # unused parameter: zeros, collectives stay aligned
    assert b0.all() and b1.all() and (o1 == 1).all()               # rank 0's occupancy state everywhere
    assert [4] in buckets and all(len(b) >= 1 for b in buckets)    # the big table is reduced in place in its own bucket


def test_shard_range_partitions():
    from cnc_b200.dp import shard_range

    for n in (0, 1, 7, 8, 4096, 150001):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


class _Enc(torch.nn.Module):
    """the two attributes ShardedTableAdam reads from a GridEncoder"""

    def __init__(self, rows, F=8):
        super().__init__()
        self.params = torch.nn.Parameter(torch.empty(rows, F))
        self.ste_binary = True


def _sharded_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cnc_b200.dp import ShardedTableAdam
    from cnc_b200.trainer import TrainStep

    g = torch.Generator().manual_seed(5)
    encs = [_Enc(1000), _Enc(136)]                      # 8000 and 1088 latents: cut at 64-element multiples, tails of 0 / 0 .. 63
    with torch.no_grad():
        for e in encs:
            e.params.copy_(torch.randn(e.params.shape, generator=g) * 0.9)     # some outside the STE window |p| <= 1
    opt = ShardedTableAdam(encs, lr=0.05, eps=1e-15, weight_decay=1e-3)
    grads = []
    for step in range(3):
        gs = [torch.randn(e.params.shape, generator=torch.Generator().manual_seed(100 * step + 10 * k + rank)) for k, e in enumerate(encs)]
        for e, gg in zip(encs, gs):
            e.params.grad = gg.clone()
        if step == 1:      # table 0 handed over early (as from inside backward): same result
            encs[0].params.grad = None
            assert opt.contribute(0, gs[0].clone())
        opt.step()
        grads.append(gs)
    before_sync = [e.params.detach().clone() for e in encs]
    spans = [(t["lo"], t["hi"], t["n_main"]) for t in opt.tables]
    planes = [(t["sign"].clone(), t["mask"].clone()) for t in opt.tables]
    opt.sync_params()
    # the skip decision of TrainStep: one empty rank -> nobody steps
    ts = TrainStep.__new__(TrainStep)
    ts.world = world
    votes = (ts._everyone_has_samples(0 if rank == 1 else 7, "cpu"), ts._everyone_has_samples(3 + rank, "cpu"))
    np_ = lambda t: t.detach().numpy().copy()      # by value: the parent may read the queue after this process is gone
    q.put((rank, [np_(t) for t in before_sync], [np_(e.params) for e in encs], spans, [(np_(a), np_(b)) for a, b in planes], votes,
           opt.comm_bytes_per_step()))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_table_adam_world2():
    """reduce-scatter + Adam on the owned rows + bit-plane all-gather == plain Adam on the rank-averaged gradient: the
    owned rows hold the true latents, the others a stand-in with the same sign and STE window; sync_params restores all"""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sharded_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=180) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process reference
    g = torch.Generator().manual_seed(5)
    ref = [torch.nn.Parameter(torch.randn(1000, 8, generator=g) * 0.9), torch.nn.Parameter(torch.randn(136, 8, generator=g) * 0.9)]
    opt = torch.optim.Adam(ref, lr=0.05, eps=1e-15, weight_decay=1e-3)
    for step in range(3):
        for k, p in enumerate(ref):
            gs = [torch.randn(p.shape, generator=torch.Generator().manual_seed(100 * step + 10 * k + r)) for r in range(world)]
            p.grad = sum(gs) / world
        opt.step()
    from cnc_b200.train_ops import planes_pack

    for rank, before, after, spans, planes, votes, comm in out:
        before, after = [torch.from_numpy(t) for t in before], [torch.from_numpy(t) for t in after]
        planes = [(torch.from_numpy(a), torch.from_numpy(b)) for a, b in planes]
        for k, p in enumerate(ref):
            lo, hi, n_main = spans[k]
            assert lo % 32 == 0 and hi % 32 == 0 and (hi - lo) * world == n_main and p.numel() - n_main < 32 * world
            want = p.detach().view(-1)
            got = before[k].view(-1)
            torch.testing.assert_close(got[lo:hi], want[lo:hi], rtol=1e-5, atol=1e-6)          # owned rows: the true latents
            torch.testing.assert_close(got[n_main:], want[n_main:], rtol=1e-5, atol=1e-6)      # replicated tail
            other = torch.ones_like(want, dtype=torch.bool)
            other[lo:hi] = False
            other[n_main:] = False
            assert other.any()
            assert torch.equal(got[other] >= 0, want[other] >= 0)                              # stand-ins: same sign ...
            assert torch.equal(got[other].abs() <= 1, want[other].abs() <= 1)                  # ... same STE window
            assert set(got[other].abs().unique().tolist()) <= {0.5, 1.5}
            s, m = planes_pack(want.contiguous())
            assert torch.equal(planes[k][0], s) and torch.equal(planes[k][1], m)               # planes of the whole table
            torch.testing.assert_close(after[k].view(-1), want, rtol=1e-5, atol=1e-6)          # after sync_params: everything
        assert votes == (False, True)
        assert comm == sum(4 * p.numel() + 2 * (spans[k][1] - spans[k][0]) // 8 for k, p in enumerate(ref))
    assert (out[0][2][0] == out[1][2][0]).all()


def test_train_ops_cpu_reference_paths():
    """planes / stand-ins / Adam of cnc_b200.train_ops on CPU tensors (the arithmetic the gloo tests run on)"""
    from cnc_b200.train_ops import adam_planes, planes_pack, surrogate_fill

    g = torch.Generator().manual_seed(0)
    p = torch.randn(256, generator=g) * 1.2
    p[:4] = torch.tensor([0.0, -0.0, 1.0, -1.0])
    s, m = planes_pack(p.clone())
    bits = lambda b: ((b.view(-1, 1).to(torch.int32) >> torch.arange(8, dtype=torch.int32)) & 1).bool().view(-1)
    assert torch.equal(bits(s), p >= 0) and torch.equal(bits(m), p.abs() <= 1)
    q = p.clone()
    surrogate_fill(q, s, m, 64, 128)
    assert torch.equal(q[64:128], p[64:128])
    rest = torch.cat([q[:64], q[128:]])
    ref = torch.cat([p[:64], p[128:]])
    assert torch.equal(rest >= 0, ref >= 0) and torch.equal(rest.abs() <= 1, ref.abs() <= 1)
    a = torch.nn.Parameter(p.clone())
    opt = torch.optim.Adam([a], lr=0.01, eps=1e-15, weight_decay=2e-6)
    b, m1, v2 = p.clone(), torch.zeros(256), torch.zeros(256)
    for step in range(1, 4):
        gr = torch.randn(256, generator=g)
        a.grad = gr.clone()
        opt.step()
        adam_planes(b, gr * 1024.0, m1, v2, step=step, lr=0.01, eps=1e-15, weight_decay=2e-6, grad_scale=1024.0, sign=s, mask=m)
    torch.testing.assert_close(b, a.detach(), rtol=1e-5, atol=1e-7)
    assert torch.equal(bits(s), b >= 0)
