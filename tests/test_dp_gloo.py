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
    q.put((rank, [p.grad.clone() if p.grad is not None else None for p in params], n, mine.origins.shape[0],
           buf.binaries.clone(), buf.occs.clone(), [list(b) for b in red.buckets]))
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
    (r0, g0, n0, c0, b0, o0, buckets), (r1, g1, n1, c1, b1, o1, _) = out
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
