"""GPU tests of gm_pairs_t.segments (a LIST cut into consecutive parts that the streaming pair kernels walk one after
another -- batches in engine.window_order): a launch with segments = W over the reordered batch is the same sum over the
same pairs as the plain launch over the original batch -- per-pair distances bit for bit (after undoing the
permutation), loss and gradient to floating-point reduction order -- for the fused (K_FUSED) and the backward (K_BWD)
kernels, ragged list lengths, lists too short to be cut, and through PairTrainer.step / step_host_grouped."""
import pytest
import torch

from helpers import rel_err

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda', 0)


def _batch(N, G, per, seed, ragged=0):
    g = torch.Generator().manual_seed(seed)
    src = torch.randperm(N, generator=g)[:G].int()
    I = src.repeat_interleave(per)
    J = torch.randint(N - 1, (G * per,), generator=g, dtype=torch.int32)
    J = torch.where(J >= I, J + 1, J)
    hops = torch.randint(1, 10, (G * per,), generator=g, dtype=torch.uint8)
    P = G * per - ragged
    return I[:P].contiguous(), J[:P].contiguous(), hops[:P].contiguous()


@pytest.mark.parametrize('n,dtype,W,ragged', [(4, torch.float32, 4, 0), (4, torch.float32, 8, 77), (3, torch.float64, 3, 5),
                                              (4, torch.float32, 64, 1), (2, torch.float32, 5, 0)])
def test_segmented_walk_is_the_same_sum(n, dtype, W, ragged):
    from graphembed import _ops, _lib as L
    from graphembed.engine import pack_hops, window_order
    from graphembed.manifolds import SymmetricPositiveDefinite
    torch.manual_seed(1)
    N, G, per = 30000, 512, 8192  # 2^22 pairs: enough for 64 parts of >= 32 pairs per warp of the persistent grid
    man = SymmetricPositiveDefinite(n)
    x = man.rand(N, out=torch.empty(0, device=DEV, dtype=dtype), ir=0.7).contiguous()
    I, J, hops = _batch(N, G, per, 7, ragged)
    order = window_order(J, N, W)
    assert torch.equal(torch.sort(order).values, torch.arange(I.numel()))
    win = (J[order].long() * W) // N
    assert bool((win[1:] >= win[:-1]).all())  # windows ascending, original (source-grouped) order inside each
    spec = _ops.LossSpec(L.GM_LOSS_QUOTIENT, True, True, alpha=1.0, eps=0.5)
    tg = _ops.TargetSpec.hops_packed(81.0)
    Id, Jp = I.to(DEV), pack_hops(J, hops).to(DEV)
    od = order.to(DEV)
    res = []
    for seg in (0, W):
        i, jp = (Id, Jp) if seg == 0 else (Id[od].contiguous(), Jp[od].contiguous())
        pairs = _ops.PairSet.from_lists(i, jp, DEV, segments=seg)
        assert pairs.c_struct().segments == seg
        grad = torch.zeros_like(x)
        acc, d2 = _ops.pairs_loss_fused(man.spec, x, pairs, tg, spec, 0.93, grad, want_d2=True)
        # backward-only kernel over the same (possibly cut) list, upstream gradient = position-independent weight
        gout = (hops.to(DEV).to(dtype) / 9.0) if seg == 0 else (hops.to(DEV).to(dtype) / 9.0)[od].contiguous()
        plain = _ops.PairSet.from_lists(i, (jp & 0x00ffffff), DEV, segments=seg)
        gb = torch.zeros_like(x)
        _ops.pairs_grad(man.spec, x, x, plain, gout, gb, gb, coef=1.0)
        res.append((acc.clone(), d2 if seg == 0 else None, d2, grad, gb))
    d2_plain, d2_cut = res[0][2], res[1][2]
    assert torch.equal(d2_plain[od], d2_cut)  # every pair computed once, with the same arithmetic
    rt = 2e-5 if dtype == torch.float32 else 1e-11
    assert rel_err(res[1][0], res[0][0]) < 1e-9
    assert rel_err(res[1][3], res[0][3]) < rt
    assert rel_err(res[1][4], res[0][4]) < rt
    assert torch.isfinite(res[1][3]).all()


def test_short_lists_ignore_the_hint():
    """A list too short to give every warp 32 pairs per part is walked in one piece: same results, nothing skipped."""
    from graphembed import _ops, _lib as L
    from graphembed.engine import pack_hops
    from graphembed.manifolds import SymmetricPositiveDefinite
    torch.manual_seed(2)
    man = SymmetricPositiveDefinite(4)
    N = 500
    x = man.rand(N, out=torch.empty(0, device=DEV, dtype=torch.float32), ir=0.7).contiguous()
    spec = _ops.LossSpec(L.GM_LOSS_STRESS, True, True, alpha=1.0, eps=0.5)
    tg = _ops.TargetSpec.hops_packed(81.0)
    for G, per, ragged in ((3, 17, 0), (40, 100, 13), (64, 1024, 1)):
        I, J, hops = _batch(N, G, per, 3, ragged)
        Id, Jp = I.to(DEV), pack_hops(J, hops).to(DEV)
        out = []
        for seg in (0, 16):
            grad = torch.zeros_like(x)
            acc, d2 = _ops.pairs_loss_fused(man.spec, x, _ops.PairSet.from_lists(Id, Jp, DEV, segments=seg), tg, spec,
                                            1.0, grad, want_d2=True)
            out.append((acc.clone(), d2, grad))
        assert torch.equal(out[0][1], out[1][1])
        assert rel_err(out[1][0], out[0][0]) < 1e-9
        assert rel_err(out[1][2], out[0][2]) < 2e-5


def test_segments_argument_is_validated():
    from graphembed import _ops
    i = torch.zeros(8, dtype=torch.int32, device=DEV)
    with pytest.raises(ValueError):
        _ops.PairSet.from_lists(i, i, DEV, segments=65)
    with pytest.raises(ValueError):
        _ops.PairSet.from_lists(i, i, DEV, segments=-1)


def test_trainer_steps_on_window_ordered_batches():
    """PairTrainer.step(segments=W) on the reordered batch == PairTrainer.step on the original batch (three RAdam
    steps: loss to 1e-6, points to fp32 rounding); step_host_grouped(segments=W) with (window, source) groups likewise."""
    from graphembed.engine import PairTrainer, pack_hops, window_order
    from graphembed.manifolds import SymmetricPositiveDefinite
    from graphembed.modules import ManifoldEmbedding
    from graphembed.objectives import StressLoss
    from graphembed.optim import RiemannianAdam
    N, G, per, W = 20000, 256, 4096, 4
    I, J, hops = _batch(N, G, per, 11)
    order = window_order(J, N, W)
    Iw, Jw, hw = I[order].contiguous(), J[order].contiguous(), hops[order].contiguous()
    # (window, source) groups of the reordered batch: run-length encode its first endpoints
    change = torch.ones(Iw.numel(), dtype=torch.bool)
    change[1:] = Iw[1:] != Iw[:-1]
    starts = torch.nonzero(change).flatten()
    sources = Iw[starts].contiguous()
    offsets = torch.cat([starts, torch.tensor([Iw.numel()])]).to(torch.int64).contiguous()
    results = []
    for mode in ('plain', 'windows', 'windows_grouped_host'):
        torch.manual_seed(3)
        emb = ManifoldEmbedding(N, [SymmetricPositiveDefinite(4)], device=DEV, dtype=torch.float32)
        opt = RiemannianAdam(emb.xs, lr=0.01, max_grad_norm=100, exact=True)
        # (the smooth loss: QuotientLoss has kinks at m == t where a different summation order flips a pair's gradient
        # sign -- in the reference just the same -- see tests/multi_gpu_worker.py)
        tr = PairTrainer(emb, opt, StressLoss(), max_hops_sq=81.0)
        losses = []
        for _ in range(3):
            if mode == 'plain':
                losses.append(tr.step(I.to(DEV), pack_hops(J, hops).to(DEV), None, epoch=1).item())
            elif mode == 'windows':
                losses.append(tr.step(Iw.to(DEV), pack_hops(Jw, hw).to(DEV), None, epoch=1, segments=W).item())
            else:
                losses.append(tr.step_host_grouped(sources.pin_memory(), offsets.pin_memory(),
                                                   pack_hops(Jw, hw).pin_memory(), None, epoch=1, segments=W))
        results.append((losses, emb.xs[0].detach().cpu().clone()))
    for r in results[1:]:
        assert max(abs(a - b) / abs(b) for a, b in zip(r[0], results[0][0])) < 1e-6
        assert rel_err(r[1], results[0][1]) < 2e-5
