"""GPU tests of row-sharded tables (gm_pairs_loss_fused_sharded, include/gm_kernels.h gm_row_shards_t) on ONE GPU: the
shards of the point / gradient table are separate allocations of the same device (the kernel does not care whether a
shard pointer is local or a peer mapping), world = 1, 2, 4, 8.  The launch is the same sum over the same pairs as the
unsharded fused kernel: per-pair distances bit for bit, loss and the re-assembled gradient to summation order -- for
explicit lists (packed and unpacked hop counts) and pairs drawn on the device (GM_PAIRS_SAMPLED).  The multi-process
path (CUDA IPC arenas, barriers, local optimizer) is tests/multi_gpu_worker.py."""
import pytest
import torch

from helpers import rel_err

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda', 0)


def _shards(t, world):
    from graphembed.parallel import cyclic_shard
    return [cyclic_shard(t, r, world) for r in range(world)]


@pytest.mark.parametrize('n,dtype,world', [(4, torch.float32, 1), (4, torch.float32, 2), (4, torch.float32, 8),
                                           (3, torch.float64, 4), (5, torch.float32, 4), (4, torch.float64, 2)])
def test_sharded_launch_is_the_unsharded_sum(n, dtype, world):
    from graphembed import _ops, _lib as L
    from graphembed.engine import pack_hops
    from graphembed.manifolds import SymmetricPositiveDefinite
    from graphembed.parallel import cyclic_unshard
    torch.manual_seed(5)
    N, G, per = 4096, 64, 700
    man = SymmetricPositiveDefinite(n)
    x = man.rand(N, out=torch.empty(0, device=DEV, dtype=dtype), ir=0.7).contiguous()
    g = torch.Generator().manual_seed(9)
    src = torch.randperm(N, generator=g)[:G].int()
    I = src.repeat_interleave(per)
    J = torch.randint(N - 1, (G * per,), generator=g, dtype=torch.int32)
    J = torch.where(J >= I, J + 1, J)
    I[-5:] = torch.randint(N, (5,), generator=g, dtype=torch.int32)  # a ragged tail: lanes of a warp with different sources
    J[-5:] = (I[-5:] + 1) % N
    hops = torch.randint(1, 10, (G * per,), generator=g, dtype=torch.uint8)
    spec = _ops.LossSpec(L.GM_LOSS_QUOTIENT, True, True, alpha=1.0, eps=0.5)
    xs = _shards(x, world)
    for packed in (True, False):
        if packed:
            pairs = _ops.PairSet.from_lists(I.to(DEV), pack_hops(J, hops).to(DEV), DEV)
            tg = _ops.TargetSpec.hops_packed(81.0)
        else:
            pairs = _ops.PairSet.from_lists(I.to(DEV), J.to(DEV), DEV)
            tg = _ops.TargetSpec.hops(hops.to(DEV), 81.0)
        grad = torch.zeros_like(x)
        acc, d2 = _ops.pairs_loss_fused(man.spec, x, pairs, tg, spec, 0.93, grad, want_d2=True)
        gs = [torch.zeros_like(s) for s in xs]
        acc_s = torch.zeros(2, dtype=torch.float64, device=DEV)
        _, d2_s = _ops.pairs_loss_fused_sharded(man.spec, [s.data_ptr() for s in xs], [t.data_ptr() for t in gs], dtype,
                                                DEV, pairs, tg, spec, 0.93, acc_s, want_d2=True)
        assert torch.equal(d2, d2_s)
        assert rel_err(acc_s, acc) < 1e-9
        assert rel_err(cyclic_unshard(gs), grad) < (2e-5 if dtype == torch.float32 else 1e-11)
        for r in range(world):  # a shard only ever receives the rows it owns: what it holds is the strided slice
            assert rel_err(gs[r], grad[r::world]) < (2e-5 if dtype == torch.float32 else 1e-11)


def test_sharded_launch_with_device_drawn_pairs():
    from graphembed import _ops, _lib as L
    from graphembed.manifolds import SymmetricPositiveDefinite
    from graphembed.parallel import cyclic_unshard
    torch.manual_seed(6)
    N, G, per, world = 8192, 50, 512, 4
    man = SymmetricPositiveDefinite(4)
    x = man.rand(N, out=torch.empty(0, device=DEV, dtype=torch.float32), ir=0.7).contiguous()
    levels = torch.randint(1, 12, (G, N), dtype=torch.uint8, device=DEV)
    src = torch.randperm(N)[:G].int().to(DEV)
    pairs = _ops.PairSet.sampled(src, levels, per, 0xABCDEF)
    tg = _ops.TargetSpec.hops_packed(121.0)
    spec = _ops.LossSpec(L.GM_LOSS_STRESS, True, True, alpha=1.0, eps=0.5)
    grad = torch.zeros_like(x)
    acc, d2 = _ops.pairs_loss_fused(man.spec, x, pairs, tg, spec, 1.1, grad, want_d2=True)
    xs = _shards(x, world)
    gs = [torch.zeros_like(s) for s in xs]
    acc_s = torch.zeros(2, dtype=torch.float64, device=DEV)
    _, d2_s = _ops.pairs_loss_fused_sharded(man.spec, [s.data_ptr() for s in xs], [t.data_ptr() for t in gs],
                                            torch.float32, DEV, pairs, tg, spec, 1.1, acc_s, want_d2=True)
    assert torch.equal(d2, d2_s)
    assert rel_err(acc_s, acc) < 1e-9
    assert rel_err(cyclic_unshard(gs), grad) < 2e-5


def test_sharded_launch_validation():
    from graphembed import _ops, _lib as L
    from graphembed.manifolds import Lorentz, SymmetricPositiveDefinite
    man = SymmetricPositiveDefinite(4)
    x = man.rand(64, out=torch.empty(0, device=DEV, dtype=torch.float32)).contiguous()
    g = torch.zeros_like(x)
    i = torch.zeros(8, dtype=torch.int32, device=DEV)
    j = torch.ones(8, dtype=torch.int32, device=DEV) | (1 << 24)
    pairs, tg = _ops.PairSet.from_lists(i, j, DEV), _ops.TargetSpec.hops_packed(9.0)
    spec = _ops.LossSpec(L.GM_LOSS_QUOTIENT, True, True, alpha=1.0, eps=0.5)
    acc = torch.zeros(2, dtype=torch.float64, device=DEV)
    with pytest.raises(ValueError):  # three shards: not a power of two
        _ops.pairs_loss_fused_sharded(man.spec, [x.data_ptr()] * 3, [g.data_ptr()] * 3, torch.float32, DEV, pairs, tg, spec,
                                      1.0, acc)
    lor = Lorentz(5)
    xl = lor.rand(64, out=torch.empty(0, device=DEV, dtype=torch.float32)).contiguous()
    with pytest.raises(RuntimeError, match='UNSUPPORTED'):  # vector manifolds: not built
        _ops.pairs_loss_fused_sharded(lor.spec, [xl.data_ptr()], [torch.zeros_like(xl).data_ptr()], torch.float32, DEV,
                                      pairs, tg, spec, 1.0, acc)
    big = SymmetricPositiveDefinite(6)  # rows too large for the streaming kernel's staging: declined, not mis-run
    xb = big.rand(64, out=torch.empty(0, device=DEV, dtype=torch.float32)).contiguous()
    with pytest.raises(RuntimeError, match='UNSUPPORTED'):
        _ops.pairs_loss_fused_sharded(big.spec, [xb.data_ptr()], [torch.zeros_like(xb).data_ptr()], torch.float32, DEV,
                                      pairs, tg, spec, 1.0, acc)
    with pytest.raises(RuntimeError, match='EINVAL'):  # node-batch enumeration is not a sharded mode
        _ops.pairs_loss_fused_sharded(man.spec, [x.data_ptr()], [g.data_ptr()], torch.float32, DEV, _ops.PairSet.triu(16, device=DEV),
                                      _ops.TargetSpec.hops_packed(9.0), spec, 1.0, acc)
