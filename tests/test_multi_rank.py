"""World-size-2 gloo test (CPU) of the multi-rank host logic: pair-sharding of a batch across ranks and the gradient
combine produce the single-rank result.  The per-rank 'kernel' here is the oracle's autograd (CPU); the GPU path
uses the same sharding helper and NCCL instead of gloo."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, os.path.join(ROOT, 'matrix-manifolds_b200'))
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import manifolds_oracle as O
    from graphembed.parallel import shard_range, allreduce_step_buffers
    torch.manual_seed(0)
    orc = O.SpdOracle(3)
    x = orc.rand(20, ir=1.0)
    P = 101
    I = torch.randint(20, (P,))
    J = (I + 1 + torch.randint(19, (P,))) % 20
    t = torch.rand(P, dtype=torch.float64) + 0.2
    lo, hi = shard_range(P, rank, world)
    xr = x.clone().requires_grad_()
    loss = O.quotient_loss(t[lo:hi], orc.dist2(xr[I[lo:hi]], xr[J[lo:hi]]), 1.0, 1)
    loss.backward()
    grad, acc = xr.grad.clone(), torch.tensor([loss.item(), 0.0], dtype=torch.float64)
    allreduce_step_buffers(grad, acc, dist.group.WORLD)
    if rank == 0:
        xs = x.clone().requires_grad_()
        full = O.quotient_loss(t, orc.dist2(xs[I], xs[J]), 1.0, 1)
        full.backward()
        out.put((float((grad - xs.grad).abs().max()), abs(acc[0].item() - full.item()), [shard_range(7, r, 3) for r in range(3)]))
    dist.destroy_process_group()


def test_pair_sharding_and_gradient_combine_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    gerr, lerr, ranges = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert gerr < 1e-12 and lerr < 1e-10
    assert ranges == [(0, 3), (3, 5), (5, 7)]  # contiguous, balanced, covering


def _shard_worker(rank, world, port, out):
    sys.path.insert(0, os.path.join(ROOT, 'matrix-manifolds_b200'))
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from graphembed.parallel import RowShards
    torch.manual_seed(rank)
    n = 12
    grad = torch.randn(n, 3, 3, dtype=torch.float64)
    every = [None] * world
    dist.all_gather_object(every, grad.clone())
    total = sum(every)
    sh = RowShards(n, dist.group.WORLD)
    mine = sh.reduce_scatter(grad.clone())
    ok_rs = torch.allclose(mine, total[sh.lo:sh.hi], atol=1e-14)
    # owner update on the owned rows, then publish
    x = torch.zeros(n, 3, 3, dtype=torch.float64)
    sh.own(x).copy_(-0.1 * mine)
    sh.all_gather(x)
    ok_ag = torch.allclose(x, -0.1 * total, atol=1e-14)
    bad = False
    try:
        RowShards(n + 1, dist.group.WORLD)
    except ValueError:
        bad = True
    out.put((rank, ok_rs, ok_ag, bad, (sh.lo, sh.hi)))
    dist.barrier()
    dist.destroy_process_group()


def test_row_shards_reduce_scatter_all_gather_world2():
    """Owner-update plumbing (gloo, CPU): reduce-scatter hands every rank the summed rows it owns, all-gather
    publishes the updated rows; uneven row counts are rejected."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert [g[4] for g in got] == [(0, 6), (6, 12)]
    assert all(g[1] and g[2] and g[3] for g in got)


def _phases_worker(rank, world, port, out):
    sys.path.insert(0, os.path.join(ROOT, 'matrix-manifolds_b200'))
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from graphembed.parallel import run_phases
    log = []

    def fail_on(r, tag):
        def fn():
            log.append(tag)
            if rank == r:
                raise RuntimeError(f'{tag} failed on rank {rank}')
        return fn

    def coll(tag):
        def fn():
            t = torch.tensor([rank])
            dist.all_reduce(t)  # a real collective: hangs or mismatches if only some ranks get here
            log.append(tag)
        return fn

    res = []
    # (a) everybody succeeds; (b) rank 1 fails in phase 2 AFTER the first collective; (c) rank 0 fails in phase 1
    for fail_rank, fail_phase in ((-1, 0), (1, 2), (0, 1)):
        log.clear()
        ok, err = run_phases(dist.group.WORLD, torch.device('cpu'), [
            (fail_on(fail_rank if fail_phase == 1 else -1, 'p1'), coll('c1')),
            (fail_on(fail_rank if fail_phase == 2 else -1, 'p2'), coll('c2')),
        ])
        res.append((ok, None if err is None else str(err), list(log)))
    t = torch.tensor([1.0])
    dist.all_reduce(t)  # the group is still in step afterwards
    out.put((rank, res, float(t.item())))
    dist.destroy_process_group()


def test_phase_vote_keeps_ranks_in_lock_step_when_one_fails():
    """parallel.run_phases (used by try_peer_arena): a failure on one rank after a collective must not leave the other
    ranks inside a barrier -- every rank votes after every phase and all of them stop at the same point."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() + 977) % 2000
    procs = [ctx.Process(target=_phases_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict()
    for _ in range(2):
        rank, res, total = q.get(timeout=120)
        got[rank] = (res, total)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank in (0, 1):
        res, total = got[rank]
        assert total == 2.0
        assert res[0] == (True, None, ['p1', 'c1', 'p2', 'c2'])
        assert res[1][0] is False and res[1][2] == ['p1', 'c1', 'p2']       # both ranks stop before c2
        assert (res[1][1] is not None) == (rank == 1)
        assert res[2][0] is False and res[2][2] == ['p1']                    # both ranks stop before c1
        assert (res[2][1] is not None) == (rank == 0)


def _row_sharded_worker(rank, world, port, out):
    """CPU analogue of engine.ShardedPairTrainer (the GPU path gathers / reduces the remote rows inside the pair kernel
    over NVLink; here the same data movement is spelled out with gloo collectives and the oracle's autograd): every rank
    keeps only its cyclic shard of the points and of the optimizer state, evaluates ITS slice of the pair batch, the
    gradient rows travel to their owners, the optimizer update is local."""
    sys.path.insert(0, os.path.join(ROOT, 'matrix-manifolds_b200'))
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import manifolds_oracle as O
    from graphembed.parallel import cyclic_shard, cyclic_unshard, gather_cyclic, shard_range
    torch.manual_seed(0)
    orc = O.SpdOracle(3)
    n, P = 24, 157
    x0 = orc.rand(n, ir=1.0)
    I = torch.randint(n, (P,))
    J = (I + 1 + torch.randint(n - 1, (P,))) % n
    t = torch.rand(P, dtype=torch.float64) + 0.2
    shard = cyclic_shard(x0, rank, world)          # the only copy of the points this rank keeps
    state = {}
    lo, hi = shard_range(P, rank, world)
    for step in range(3):
        full = gather_cyclic(shard)                # (GPU: row gathers over NVLink inside the kernel)
        xr = full.clone().requires_grad_()
        O.quotient_loss(t[lo:hi], orc.dist2(xr[I[lo:hi]], xr[J[lo:hi]]), 1.0, step + 1).backward()
        g = xr.grad.clone()
        dist.all_reduce(g)                         # (GPU: red.global.add into the owning shard)
        shard = O.radam_step(orc, shard, cyclic_shard(g, rank, world), state, lr=0.01, max_grad_norm=100,
                             exact=True).detach()
    result = gather_cyclic(shard)
    if rank == 0:
        x, st = x0.clone(), {}
        for step in range(3):
            xs = x.clone().requires_grad_()
            O.quotient_loss(t, orc.dist2(xs[I], xs[J]), 1.0, step + 1).backward()
            x = O.radam_step(orc, x, xs.grad, st, lr=0.01, max_grad_norm=100, exact=True).detach()
        shards = [cyclic_shard(x, r, world) for r in range(world)]
        out.put((float((result - x).abs().max() / x.abs().max()), bool(torch.equal(cyclic_unshard(shards), x)),
                 tuple(shard.shape)))
    dist.destroy_process_group()


def test_row_sharded_step_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_row_sharded_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    err, round_trip, shape = q.get(timeout=180)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert err < 1e-12 and round_trip and shape == (12, 3, 3)
