"""GPU suite: the peer-memory exchange of the data-parallel step (csrc/peer.cu, cnc_b200/peer.py, dp.ShardedTableAdam with
exchange="peer").  Two processes; each takes its own GPU when the box has two, otherwise both share cuda:0 (CUDA IPC maps
the other process's allocation either way; the barrier kernels then meet through time slicing).  The control plane is a gloo
group, so nothing here depends on NCCL."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _Enc(torch.nn.Module):
    """the two attributes ShardedTableAdam reads from a GridEncoder"""

    def __init__(self, rows, F=8, device="cpu"):
        super().__init__()
        self.params = torch.nn.Parameter(torch.empty(rows, F, device=device))
        self.ste_binary = True


def _init(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dev = torch.device("cuda", rank % torch.cuda.device_count())
    torch.cuda.set_device(dev)
    return dev


def _primitives_worker(rank, world, port, q):
    dev = _init(rank, world, port)
    from cnc_b200 import peer as P

    n = 4096 * 8 + 64
    mem = P.PeerMemory(4 * n)
    sig = P.PeerSignals(timeout_ms=60000)
    mine = mem.tensor(0, n)
    vals = torch.randn(n, generator=torch.Generator().manual_seed(40 + rank))
    mine.copy_(vals)
    torch.cuda.synchronize()
    sig.barrier(0)
    # every rank averages "its" half (plus a ragged second call for the tail) out of both buffers
    S = 4096 * 4
    out = torch.empty(S, device=dev)
    tail = torch.empty(64, device=dev)
    P.reduce_rows(mem.pointer_array(), world, rank * S, S, 1.0 / world, out)
    P.reduce_rows(mem.pointer_array(), world, 2 * S, 64, 1.0 / world, tail, blocks=3)
    sig.barrier(0)                                   # everybody has read: the buffers may change
    # push: each rank owns words [rank * 100, rank * 100 + 100) and [5000 + 7 * rank, ... + 7) of a uint32 arena
    arena = P.PeerMemory(4 * 8192)
    words = arena.tensor(0, 8192, torch.int32)
    words[rank * 100:rank * 100 + 100] = torch.arange(100, dtype=torch.int32, device=dev) + 1000 * (rank + 1)
    words[5000 + 7 * rank:5007 + 7 * rank] = -(rank + 1)
    torch.cuda.synchronize()
    sig.barrier(1)
    P.push_words(arena, [(rank * 100, 100), (5000 + 7 * rank, 7)])
    sig.barrier(1)
    # minimum over the ranks of one number each (the skip vote), twice in a row (value words alternate by epoch parity)
    mins = torch.zeros(2, dtype=torch.int32, device=dev)
    got_min = []
    for step, value in enumerate((100 + 7 * rank, 0 if rank == 1 else 5)):
        sig.minimum(10, value, mins)
        got_min.append(int(mins.view(torch.int64)[0]))
    # bucketed gradient average through a mapped buffer (dp.GradAllReducer.use_peer_memory), twice, second time with a
    # parameter that has no gradient on rank 1
    from cnc_b200.dp import GradAllReducer

    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(33, 7, device=dev)), torch.nn.Parameter(torch.zeros(5, device=dev)),
              torch.nn.Parameter(torch.zeros(160, 160, device=dev))]
    red = GradAllReducer(params)
    red.use_peer_memory(sig, slot=4)
    avg = []
    for step in range(2):
        for i, p in enumerate(params):
            p.grad = torch.randn(p.shape, generator=torch.Generator().manual_seed(1000 * step + 10 * i + rank)).to(dev)
        if step == 1 and rank == 1:
            params[1].grad = None
        red.reduce()
        avg.append([p.grad.cpu().numpy().copy() for p in params])
    torch.cuda.synchronize()
    q.put((rank, out.cpu().numpy().copy(), tail.cpu().numpy().copy(), words.cpu().numpy().copy(), got_min, avg))
    dist.barrier()
    del mine, words
    del red, params
    for m in (arena, mem):
        m.close()
    dist.destroy_process_group()


def _run(worker, world=2, timeout=240):
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        out = sorted([q.get(timeout=timeout) for _ in range(world)], key=lambda t: t[0])
    finally:
        for p in procs:
            p.join(timeout=60)
            if p.is_alive():
                p.kill()
    assert all(p.exitcode == 0 for p in procs)
    return out


@pytest.mark.timeout(300)
def test_peer_barrier_reduce_push_world2():
    world = 2
    out = _run(_primitives_worker, world)
    n, S = 4096 * 8 + 64, 4096 * 4
    vals = [torch.randn(n, generator=torch.Generator().manual_seed(40 + r)) for r in range(world)]
    want = (vals[0] + vals[1]) * 0.5                 # rank order, then the scale: what the kernel does
    for rank, got, tail, words, got_min, avg in out:
        assert got_min == [100, 0]
        for step in range(2):
            for i, a in enumerate(avg[step]):
                gs = [torch.randn(a.shape, generator=torch.Generator().manual_seed(1000 * step + 10 * i + r)) for r in range(world)]
                if step == 1 and i == 1:
                    gs[1] = torch.zeros_like(gs[1])
                assert torch.equal(torch.from_numpy(a), (gs[0] + gs[1]) * 0.5)
        assert torch.equal(torch.from_numpy(got), want[rank * S:(rank + 1) * S])
        assert torch.equal(torch.from_numpy(tail), want[2 * S:2 * S + 64])
        w = torch.from_numpy(words)
        for r in range(world):
            assert torch.equal(w[r * 100:r * 100 + 100], torch.arange(100, dtype=torch.int32) + 1000 * (r + 1))
            assert (w[5000 + 7 * r:5007 + 7 * r] == -(r + 1)).all()
        untouched = torch.ones(8192, dtype=torch.bool)
        for r in range(world):
            untouched[r * 100:r * 100 + 100] = False
            untouched[5000 + 7 * r:5007 + 7 * r] = False
        assert (w[untouched] == 0).all()


def _sharded_worker(rank, world, port, q):
    dev = _init(rank, world, port)
    from cnc_b200.dp import ShardedTableAdam

    g = torch.Generator().manual_seed(5)
    encs = [_Enc(1000, device=dev), _Enc(136, device=dev)]     # 8000 and 1088 latents: the second has a replicated tail
    with torch.no_grad():
        for e in encs:
            e.params.copy_(torch.randn(e.params.shape, generator=g) * 0.9)
    opt = ShardedTableAdam(encs, lr=0.05, eps=1e-15, weight_decay=1e-3, exchange="peer")
    assert opt.peer is not None
    for step in range(4):
        gs = [torch.randn(e.params.shape, generator=torch.Generator().manual_seed(100 * step + 10 * k + rank)).to(dev)
              for k, e in enumerate(encs)]
        for e, gg in zip(encs, gs):
            e.params.grad = gg.clone()
        if step == 1:      # table 0 accumulated in the shared buffer and handed over early, as from inside backward
            encs[0].params.grad = None
            buf = opt.grad_buffer(0)
            buf.zero_().add_(gs[0])
            assert opt.contribute(0, buf)
        if step == 2 and rank == 1:      # a rank whose batch produced nothing: no gradient at all
            for e in encs:
                e.params.grad = None
        opt.step()
    torch.cuda.synchronize()
    np_ = lambda t: t.detach().cpu().numpy().copy()
    spans = [(t["lo"], t["hi"], t["n_main"]) for t in opt.tables]
    q.put((rank, [np_(e.params) for e in encs], spans, [(np_(t["sign"]), np_(t["mask"])) for t in opt.tables],
           opt.link_bytes_per_step()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_table_adam_peer_world2():
    """loads from the peers' gradient buffers + Adam on the owned rows + plane words stored into the peers == plain Adam on
    the rank-averaged gradient (the same statement tests/test_dp_gloo.py makes about the collective version)"""
    world = 2
    out = _run(_sharded_worker, world)
    g = torch.Generator().manual_seed(5)
    ref = [torch.nn.Parameter(torch.randn(1000, 8, generator=g) * 0.9), torch.nn.Parameter(torch.randn(136, 8, generator=g) * 0.9)]
    ropt = torch.optim.Adam(ref, lr=0.05, eps=1e-15, weight_decay=1e-3)
    for step in range(4):
        for k, p in enumerate(ref):
            gs = [torch.randn(p.shape, generator=torch.Generator().manual_seed(100 * step + 10 * k + r)) for r in range(world)]
            if step == 2:
                gs[1] = torch.zeros_like(gs[1])
            p.grad = sum(gs) / world
        ropt.step()
    for rank, params, spans, planes, link in out:
        for k, p in enumerate(ref):
            lo, hi, n_main = spans[k]
            want, got = p.detach().view(-1), torch.from_numpy(params[k]).view(-1)
            torch.testing.assert_close(got[lo:hi], want[lo:hi], rtol=1e-5, atol=1e-6)          # owned rows: the true latents
            torch.testing.assert_close(got[n_main:], want[n_main:], rtol=1e-5, atol=1e-6)      # replicated tail
            other = torch.ones_like(want, dtype=torch.bool)
            other[lo:hi] = False
            other[n_main:] = False
            firm = (want.abs() > 1e-5) & ((want.abs() - 1).abs() > 1e-5)                       # away from the two thresholds
            sel = other & firm
            assert sel.any()
            assert torch.equal(got[sel] >= 0, want[sel] >= 0)                                  # stand-ins: same sign ...
            assert torch.equal(got[sel].abs() <= 1, want[sel].abs() <= 1)                      # ... same STE window
            assert set(got[other].abs().unique().tolist()) <= {0.5, 1.5}
            sign = torch.from_numpy(planes[k][0])
            bits = ((sign.view(-1, 1).to(torch.int32) >> torch.arange(8, dtype=torch.int32)) & 1).bool().view(-1)
            assert torch.equal(bits[firm], (want >= 0)[firm])                                  # plane of the whole table
        assert link == sum(4 * (spans[k][1] - spans[k][0] + p.numel() - spans[k][2]) + 2 * (spans[k][1] - spans[k][0]) // 8
                           for k, p in enumerate(ref))
    # both ranks hold the same planes
    for k in range(2):
        assert (out[0][3][k][0] == out[1][3][k][0]).all() and (out[0][3][k][1] == out[1][3][k][1]).all()
