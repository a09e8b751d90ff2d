"""GPU tests of GM_PAIRS_SAMPLED (pairs drawn inside the pair kernels): the device draw equals the host statement
oracle/sampler_oracle.py bit for bit -- indices and hop counts -- for the SPD streaming kernel, the generic SPD
kernel and the vector-manifold kernels, at toy size and at BASELINE config 5's full size (2 M nodes, 2^24 pairs); and
PairTrainer.step_sampled / step_sampled_host is the same step as PairTrainer.step on the host-drawn lists."""
import numpy as np
import pytest
import torch

import sampler_oracle as S
from helpers import rel_err

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda', 0)


def _levels(G, N, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(1, 12, (G, N), generator=g, dtype=torch.uint8)


@pytest.mark.parametrize('per_src', [64, 100])  # a power of two (shift) and a general divisor
def test_device_draw_equals_host_statement(per_src):
    """Euclidean points x[v] = (v, v^2) in fp64: d2 identifies (i, j) exactly; the fused loss identifies the hops."""
    from graphembed import _ops, _lib as L
    from graphembed.manifolds import Euclidean, Lorentz, SymmetricPositiveDefinite
    N, G, seed = 5000, 37, 0xDEADBEEF12345
    levels = _levels(G + 5, N, 1)
    rng = np.random.RandomState(3)
    src = rng.choice(N, G, replace=False).astype(np.int32)
    slots = rng.permutation(G + 5)[:G].astype(np.int32)
    I, J, H = S.sample_pairs(src, levels.numpy(), per_src, seed, slots=slots)
    P = G * per_src
    lv, sd, sl = levels.to(DEV), torch.from_numpy(src).to(DEV), torch.from_numpy(slots).to(DEV)
    sampled = _ops.PairSet.sampled(sd, lv, per_src, seed, slots=sl)
    listed = _ops.PairSet.from_lists(torch.from_numpy(I).to(DEV), torch.from_numpy(J).to(DEV), DEV)
    assert sampled.P == P
    v = torch.arange(N, dtype=torch.float64, device=DEV)
    x = torch.stack([v, v * v], dim=1).contiguous()
    man = Euclidean(2)
    d_s = _ops.pairs_dist2(man.spec, x, x, sampled)
    d_l = _ops.pairs_dist2(man.spec, x, x, listed)
    assert torch.equal(d_s, d_l)
    want = (I.astype(np.float64) - J) ** 2 + (I.astype(np.float64) ** 2 - J.astype(np.float64) ** 2) ** 2
    assert np.array_equal(d_s.cpu().numpy(), want)
    # hop counts: fused loss with packed targets, sampled vs explicit (j | hop << 24) list, every kernel family
    packed = torch.from_numpy((J.astype(np.int64) | (H.astype(np.int64) << 24)).astype(np.int32)).to(DEV)
    listed_p = _ops.PairSet.from_lists(torch.from_numpy(I).to(DEV), packed, DEV)
    tg = _ops.TargetSpec.hops_packed(121.0)
    spec = _ops.LossSpec(L.GM_LOSS_QUOTIENT, True, True, alpha=1.0, eps=0.5)
    torch.manual_seed(0)
    hint32 = torch.empty(0, device=DEV, dtype=torch.float32)
    for man, xs in ((SymmetricPositiveDefinite(4), None), (SymmetricPositiveDefinite(3), None), (Lorentz(6), None)):
        xs = man.rand(N, out=hint32, ir=0.5).contiguous()
        outs = []
        for pairs in (sampled, listed_p):
            grad = torch.zeros_like(xs)
            acc, d2 = _ops.pairs_loss_fused(man.spec, xs, pairs, tg, spec, 0.9, grad, want_d2=True)
            outs.append((acc.clone(), d2, grad))
        assert torch.equal(outs[0][1], outs[1][1]), type(man).__name__
        assert rel_err(outs[0][0], outs[1][0]) < 1e-12
        assert rel_err(outs[0][2], outs[1][2]) < 1e-5  # float atomics: same terms, different arrival order


def test_validation_of_sampled_pairs():
    from graphembed import _ops, _lib as L
    from graphembed.manifolds import SymmetricPositiveDefinite
    lv = _levels(4, 100, 0).to(DEV)
    src = torch.arange(4, dtype=torch.int32, device=DEV)
    pairs = _ops.PairSet.sampled(src, lv, 8, 1)
    man = SymmetricPositiveDefinite(2)
    x = man.rand(100, out=torch.empty(0, device=DEV, dtype=torch.float32)).contiguous()
    spec = _ops.LossSpec(L.GM_LOSS_STRESS)
    with pytest.raises(RuntimeError):  # drawn pairs bring their own hop counts: only packed targets
        _ops.pairs_loss_fused(man.spec, x, pairs, _ops.TargetSpec.vector(torch.ones(32, device=DEV)), spec, 1.0,
                              torch.zeros_like(x))
    with pytest.raises(ValueError):
        _ops.PairSet.sampled(src.long(), lv, 8, 1)


def test_step_sampled_is_the_step_on_the_host_drawn_lists():
    from graphembed.engine import PairTrainer, pack_hops
    from graphembed.manifolds import SymmetricPositiveDefinite
    from graphembed.modules import ManifoldEmbedding
    from graphembed.objectives import QuotientLoss
    from graphembed.optim import RiemannianAdam
    N, G, per = 3000, 50, 256
    levels = _levels(G, N, 5)
    lv = levels.to(DEV)
    res = []
    for mode in ('sampled_host', 'lists'):
        torch.manual_seed(1)
        emb = ManifoldEmbedding(N, [SymmetricPositiveDefinite(4)], device=DEV, dtype=torch.float32)
        opt = RiemannianAdam(emb.xs, lr=0.01, max_grad_norm=100, exact=True)
        tr = PairTrainer(emb, opt, QuotientLoss(), max_hops_sq=121.0)
        losses = []
        for step in range(3):
            src = np.random.RandomState(step).choice(N, G, replace=False).astype(np.int32)
            seed = 77 + (step << 32)
            if mode == 'sampled_host':
                losses.append(tr.step_sampled_host(torch.from_numpy(src).pin_memory(), lv, per, seed, epoch=1))
            else:
                I, J, H = S.sample_pairs(src, levels.numpy(), per, seed)
                jp = pack_hops(torch.from_numpy(J), torch.from_numpy(H))
                losses.append(tr.step(torch.from_numpy(I).to(DEV), jp.to(DEV), None, epoch=1).item())
        res.append((losses, emb.xs[0].detach().cpu()))
    assert max(abs(a - b) / abs(b) for a, b in zip(*[r[0] for r in res])) < 1e-6
    assert rel_err(res[0][1], res[1][1]) < 1e-5


def test_full_size_draw_config5():
    """2 M nodes, 1024 sources x 16384 targets = 2^24 pairs: the streaming kernel's draw against the host statement
    (d2 bitwise equal to the LIST launch on the host-drawn pairs), loss equal, finite gradient."""
    from graphembed import _ops, _lib as L
    from graphembed.manifolds import SymmetricPositiveDefinite
    N, G, per, seed = 2_000_000, 1024, 16384, 0x1234ABCD5678
    g = torch.Generator().manual_seed(9)
    levels = torch.randint(1, 20, (64, N), generator=g, dtype=torch.uint8)  # 64 resident rows, shared through slots
    slots = torch.randint(64, (G,), generator=g, dtype=torch.int32)
    src = torch.randperm(N, generator=g)[:G].int()
    I, J, H = S.sample_pairs(src.numpy(), levels.numpy(), per, seed, slots=slots.numpy())
    man = SymmetricPositiveDefinite(4)
    torch.manual_seed(0)
    x = man.rand(N, out=torch.empty(0, device=DEV, dtype=torch.float32)).contiguous()
    lv = levels.to(DEV)
    sampled = _ops.PairSet.sampled(src.to(DEV), lv, per, seed, slots=slots.to(DEV))
    packed = torch.from_numpy((J.astype(np.int64) | (H.astype(np.int64) << 24)).astype(np.int32)).to(DEV)
    listed = _ops.PairSet.from_lists(torch.from_numpy(I).to(DEV), packed, DEV)
    tg = _ops.TargetSpec.hops_packed(361.0)
    spec = _ops.LossSpec(L.GM_LOSS_QUOTIENT, True, True, alpha=1.0, eps=0.5)
    outs = []
    for pairs in (sampled, listed):
        grad = torch.zeros_like(x)
        acc, d2 = _ops.pairs_loss_fused(man.spec, x, pairs, tg, spec, 0.97, grad, want_d2=True)
        outs.append((acc.clone(), d2, grad))
    assert torch.equal(outs[0][1], outs[1][1])
    assert rel_err(outs[0][0], outs[1][0]) < 1e-12
    assert rel_err(outs[0][2], outs[1][2]) < 1e-5 and torch.isfinite(outs[0][2]).all()
