"""GPU tests of the deterministic gradient accumulation option (gm_segment_sum, PairTrainer(deterministic=True); SURVEY 8a
A11): the fixed-order scatter-add equals a host loop that adds in the stated order BIT FOR BIT, two runs of a training
trajectory give identical bits, and the trajectory agrees with the default (atomics) path to fp rounding."""
import numpy as np
import pytest
import torch

from helpers import rel_err

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda', 0)


def _host_fixed_order(rows, index, n_out, chunk):
    """The order gm_segment_sum is specified to add in: per destination row, entries in ascending position, chunks of
    `chunk` summed left to right, then the chunk sums left to right -- in the dtype of the rows."""
    out = np.zeros((n_out,) + rows.shape[1:], dtype=rows.dtype)
    order = np.argsort(index, kind='stable')
    sorted_idx = index[order]
    start = 0
    while start < len(order):
        end = start
        while end < len(order) and sorted_idx[end] == sorted_idx[start]:
            end += 1
        total = np.zeros(rows.shape[1:], dtype=rows.dtype)
        for c0 in range(start, end, chunk):
            part = np.zeros(rows.shape[1:], dtype=rows.dtype)
            for k in range(c0, min(c0 + chunk, end)):
                part = part + rows[order[k]]
            total = total + part
        out[sorted_idx[start]] = total
        start = end
    return out


@pytest.mark.parametrize('dtype,shape', [(torch.float32, (4, 4)), (torch.float64, (3, 3)), (torch.float32, (11,))])
def test_fixed_order_scatter_add_bit_exact(dtype, shape):
    from graphembed import _ops
    g = torch.Generator().manual_seed(0)
    n_out, M, chunk = 50, 3000, 16
    rows = torch.randn((M,) + shape, generator=g, dtype=dtype) * torch.logspace(-3, 3, M, dtype=dtype).reshape((M,) + (1,) * len(shape))
    index = torch.randint(n_out - 5, (M,), generator=g)  # (rows n_out-5 .. n_out-1 receive nothing)
    index[:700] = 7                                      # one heavy row: many chunks
    out = torch.zeros((n_out,) + shape, dtype=dtype, device=DEV)
    _ops.scatter_add_rows_deterministic(rows.to(DEV), index.to(DEV), out, chunk=chunk)
    want = _host_fixed_order(rows.numpy(), index.numpy(), n_out, chunk)
    assert np.array_equal(out.cpu().numpy(), want)
    ref = torch.zeros((n_out,) + shape, dtype=torch.float64).index_add_(0, index, rows.double())
    assert rel_err(out.cpu().double(), ref) < (1e-5 if dtype == torch.float32 else 1e-13)
    again = torch.zeros_like(out)
    _ops.scatter_add_rows_deterministic(rows.to(DEV), index.to(DEV), again, chunk=chunk)
    assert torch.equal(out, again)


def _run(deterministic, steps=3, dtype=torch.float32, packed=True):
    from graphembed.engine import PairTrainer, pack_hops
    from graphembed.manifolds import SymmetricPositiveDefinite
    from graphembed.modules import ManifoldEmbedding
    from graphembed.objectives import StressLoss
    from graphembed.optim import RiemannianAdam
    torch.manual_seed(3)
    N, G, per = 3000, 64, 1024
    emb = ManifoldEmbedding(N, [SymmetricPositiveDefinite(4)], device=DEV, dtype=dtype)
    opt = RiemannianAdam(emb.xs, lr=0.01, max_grad_norm=100, exact=True)
    tr = PairTrainer(emb, opt, StressLoss(), max_hops_sq=81.0, deterministic=deterministic)
    g = torch.Generator().manual_seed(5)
    src = torch.randperm(N, generator=g)[:G].int()
    I = src.repeat_interleave(per).contiguous()
    J = torch.randint(N - 1, (G * per,), generator=g, dtype=torch.int32)
    J = torch.where(J >= I, J + 1, J).contiguous()
    hops = torch.randint(1, 10, (G * per,), generator=g, dtype=torch.uint8)
    losses = []
    for s in range(steps):
        if packed:
            losses.append(tr.step(I.to(DEV), pack_hops(J, hops).to(DEV), None, epoch=s + 1).item())
        else:
            losses.append(tr.step(I.to(DEV), J.to(DEV), hops.to(DEV), epoch=s + 1).item())
    return losses, emb.xs[0].detach().clone(), tr.grad.clone()


@pytest.mark.parametrize('dtype', [torch.float32, torch.float64])
def test_deterministic_trainer_is_reproducible_and_agrees_with_the_default_path(dtype):
    l1, x1, g1 = _run(True, dtype=dtype)
    l2, x2, g2 = _run(True, dtype=dtype)
    assert torch.equal(x1, x2) and torch.equal(g1, g2)  # same bits, run after run
    l3, x3, _ = _run(True, dtype=dtype, packed=False)   # the hop-count format does not matter
    assert torch.equal(x1, x3)
    l0, x0, _ = _run(False, dtype=dtype)
    tol = 2e-5 if dtype == torch.float32 else 1e-10
    assert rel_err(x1, x0) < tol
    assert max(abs(a - b) / abs(b) for a, b in zip(l1, l0)) < (1e-5 if dtype == torch.float32 else 1e-10)


def test_deterministic_option_declines_what_it_cannot_order():
    from graphembed import _ops
    from graphembed.engine import PairTrainer
    from graphembed.manifolds import SymmetricPositiveDefinite
    from graphembed.modules import ManifoldEmbedding
    from graphembed.objectives import StressLoss
    from graphembed.optim import RiemannianAdam
    emb = ManifoldEmbedding(256, [SymmetricPositiveDefinite(4)], device=DEV, dtype=torch.float32)
    tr = PairTrainer(emb, RiemannianAdam(emb.xs, lr=0.01), StressLoss(), max_hops_sq=9.0, deterministic=True)
    levels = torch.randint(1, 4, (4, 256), dtype=torch.uint8, device=DEV)
    with pytest.raises(ValueError):  # pairs drawn inside the kernel have no list to sort
        tr.step_sampled(torch.arange(4, dtype=torch.int32, device=DEV), levels, 8, seed=1)
    with pytest.raises(RuntimeError):
        _ops.segment_sum(torch.zeros(4, 2, device=DEV), torch.zeros(1, dtype=torch.int32, device=DEV),
                         torch.ones(1, dtype=torch.int64, device=DEV), torch.zeros(1, 2, device=DEV))
