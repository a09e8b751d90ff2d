"""GPU tests of the 2-byte-per-pair upload format (engine.pack_hops2 / gm_unpack_pairs2): the device expansion equals
the packed 4-byte words of the sorted batch bit for bit -- ragged and empty groups, groups longer than one scan tile,
gaps at the 13-bit limit -- and PairTrainer.step_host_grouped on the 2-byte words is the step on the explicit lists."""
import pytest
import torch

from helpers import rel_err

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda', 0)


def test_unpack_pairs2_bit_exact():
    from graphembed import _ops
    from graphembed.engine import pack_hops, pack_hops2
    g = torch.Generator().manual_seed(1)
    N = 2_000_000
    counts = torch.tensor([0, 5, 16384, 1, 0, 2048, 2049, 300, 7000, 0], dtype=torch.int64)
    offsets = torch.cat([torch.zeros(1, dtype=torch.int64), counts.cumsum(0)])
    P = int(offsets[-1])
    J = torch.randint(N, (P,), generator=g, dtype=torch.int32)
    J[5:5 + 16384] = torch.randint(N, (16384,), generator=g, dtype=torch.int32)
    lo = int(offsets[7])  # a group with duplicates and gaps of exactly 8191
    J[lo:lo + 300] = (torch.arange(300, dtype=torch.int32) // 2) * 8191 + 12345
    H = torch.randint(1, 9, (P,), generator=g, dtype=torch.uint8)
    # sparse groups of a 2 M-row table have gaps far beyond 13 bits: thin them so that the batch fits
    for gi in (1, 3, 5, 6, 8):
        a, b = int(offsets[gi]), int(offsets[gi + 1])
        J[a:b] = (torch.randint(4000, (b - a,), generator=g, dtype=torch.int32).cumsum(0) % N).int()
    packed = pack_hops2(offsets, J, H)
    if packed is None:  # (a random gap of the dense group above 8191: practically impossible, but say so)
        pytest.skip('random batch does not fit the format')
    words, bases, order = packed
    out = torch.full((P + 3,), -1, dtype=torch.int32, device=DEV)
    _ops.unpack_pairs2(words.to(DEV), bases.to(DEV), offsets.to(DEV), out)
    want = pack_hops(J[order].contiguous(), H[order].contiguous())
    assert torch.equal(out[:P].cpu(), want)
    assert bool((out[P:] == -1).all())  # nothing written past the batch
    # the same pass can write the first-endpoint vector (gm_expand_groups folded in)
    rows = torch.arange(100, 100 + counts.numel(), dtype=torch.int32)
    out2 = torch.full((P,), -1, dtype=torch.int32, device=DEV)
    out_i = torch.full((P + 2,), -7, dtype=torch.int32, device=DEV)
    _ops.unpack_pairs2(words.to(DEV), bases.to(DEV), offsets.to(DEV), out2, group_rows=rows.to(DEV), out_i=out_i)
    assert torch.equal(out2.cpu(), want)
    assert torch.equal(out_i[:P].cpu(), torch.repeat_interleave(rows, counts))
    assert bool((out_i[P:] == -7).all())
    # tables that do not start on a 16-byte boundary take the scalar path: same words
    w_off = torch.cat([torch.zeros(1, dtype=torch.int16), words]).to(DEV)[1:]
    out3 = torch.full((P + 1,), -1, dtype=torch.int32, device=DEV)[1:]
    _ops.unpack_pairs2(w_off, bases.to(DEV), offsets.to(DEV), out3)
    assert torch.equal(out3.cpu(), want)
    # every group keeps its pairs (a permutation inside the group), sorted by row
    for gi in range(counts.numel()):
        a, b = int(offsets[gi]), int(offsets[gi + 1])
        assert sorted(order[a:b].tolist()) == list(range(a, b))
        jj = J[order][a:b]
        assert bool((jj[1:] >= jj[:-1]).all())


def test_pack_hops2_declines_what_does_not_fit():
    from graphembed.engine import pack_hops2
    offsets = torch.tensor([0, 3], dtype=torch.int64)
    H = torch.tensor([1, 2, 3], dtype=torch.uint8)
    assert pack_hops2(offsets, torch.tensor([0, 8191, 8192], dtype=torch.int32), H) is not None
    assert pack_hops2(offsets, torch.tensor([0, 8192, 8193], dtype=torch.int32), H) is None       # gap of 8192
    assert pack_hops2(offsets, torch.tensor([0, 1, 2], dtype=torch.int32), torch.tensor([1, 9, 3], dtype=torch.uint8)) is None
    assert pack_hops2(offsets, torch.tensor([0, 1, 2], dtype=torch.int32), torch.tensor([0, 1, 3], dtype=torch.uint8)) is None


def test_grouped_host_step_on_two_byte_words_matches_list_step():
    from graphembed.engine import PairTrainer, pack_hops, pack_hops2
    from graphembed.manifolds import SymmetricPositiveDefinite
    from graphembed.modules import ManifoldEmbedding
    from graphembed.objectives import StressLoss
    from graphembed.optim import RiemannianAdam
    n, G, per = 3000, 37, 1500
    g = torch.Generator().manual_seed(5)
    src = torch.randperm(n, generator=g)[:G].int()
    offsets = torch.arange(G + 1, dtype=torch.int64) * per
    I = src.repeat_interleave(per).contiguous()
    J = torch.randint(n - 1, (G * per,), generator=g, dtype=torch.int32)
    J = torch.where(J >= I, J + 1, J).contiguous()
    hops = torch.randint(1, 9, (G * per,), generator=g, dtype=torch.uint8)
    words, bases, order = pack_hops2(offsets, J, hops)
    results = []
    for mode in ('lists', 'two_byte', 'two_byte_pipelined'):
        torch.manual_seed(3)
        emb = ManifoldEmbedding(n, [SymmetricPositiveDefinite(4)], device=DEV, dtype=torch.float32)
        opt = RiemannianAdam(emb.xs, lr=0.01, max_grad_norm=100, exact=True)
        tr = PairTrainer(emb, opt, StressLoss(), max_hops_sq=64.0)
        losses = []
        pin = lambda t: t.pin_memory()  # noqa: E731
        batch = (pin(src), pin(offsets), pin(words), None, pin(bases))
        for step in range(3):
            if mode == 'lists':
                losses.append(tr.step(I.to(DEV), pack_hops(J, hops).to(DEV), None, epoch=1).item())
            elif mode == 'two_byte':
                losses.append(tr.step_host_grouped(*batch[:4], epoch=1, bases=batch[4]))
            else:  # next batch uploaded on the copy stream while this one runs
                losses.append(tr.step_host_grouped(*batch[:4], epoch=1, bases=batch[4], next_batch=batch))
        results.append((losses, emb.xs[0].detach().cpu().clone()))
    for r in results[1:]:
        assert max(abs(a - b) / abs(b) for a, b in zip(r[0], results[0][0])) < 1e-6
        assert rel_err(r[1], results[0][1]) < 2e-5
    with pytest.raises(ValueError):
        tr.step_host_grouped(*batch[:4], epoch=1)  # the words without their bases
