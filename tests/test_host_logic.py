"""CPU tests of host-side pieces that carry semantics of their own: the lazy scalar log of TrainingEngine, the cached
softplus(scale), the optimizer argument builders that gm_train_epoch relies on, and the reference-facing structure of
products.Embedding / graphembed.pyx."""
import numpy as np
import pytest
import torch


def test_scalar_log_is_lazy_for_device_values_only():
    from graphembed.train import ScalarLog
    log = ScalarLog(None, tensorboard=False)
    log.add_scalar('a', 1.5, 1)
    log.add_scalar('a', torch.tensor(2.5), 2)  # CPU tensor: converted at once
    assert log._pending == [] and log.history['a'] == [(1, 1.5), (2, 2.5)]

    class FakeCuda(torch.Tensor):  # a tensor that claims to live on the device: must be queued, not synchronised on
        @property
        def is_cuda(self):
            return True

    v = torch.tensor(3.25).as_subclass(FakeCuda)
    log.add_scalar('a', v, 3)
    assert len(log._pending) == 1 and len(log._history['a']) == 2
    assert log.history['a'][-1] == (3, 3.25) and log._pending == []  # looking at it flushes, in logging order
    log.add_scalar('b', torch.tensor(7.0).as_subclass(FakeCuda), 1)
    log.close()
    assert log._history['b'] == [(1, 7.0)]


def test_softplus_cache_follows_in_place_updates_and_storage_swaps():
    from graphembed.modules import _softplus_value
    p = torch.nn.Parameter(torch.tensor(0.5, dtype=torch.float64))
    sp = lambda v: float(torch.nn.functional.softplus(torch.tensor(v, dtype=torch.float64)))  # noqa: E731
    assert _softplus_value(p) == pytest.approx(sp(0.5), rel=1e-15)
    with torch.no_grad():
        p.add_(1.0)
    assert _softplus_value(p) == pytest.approx(sp(1.5), rel=1e-15)
    p.data = torch.tensor(0.1, dtype=torch.float64)  # no version bump: the storage pointer is part of the key
    assert _softplus_value(p) == pytest.approx(sp(0.1), rel=1e-15)
    assert _softplus_value(p) is not None and getattr(p, '_gm_softplus')[1] == _softplus_value(p)


def test_optimizer_kernel_args_mirror_reference_hyperparameters():
    """RiemannianAdam / RiemannianSGD._kernel_args: what step() and the one-call epoch hand to the kernels
    (optim/radam.py:43-98, optim/rsgd.py:40-82)."""
    from graphembed import _lib as L
    from graphembed.manifolds import Lorentz
    from graphembed.modules import ManifoldParameter
    from graphembed.optim import RiemannianAdam, RiemannianSGD
    x = ManifoldParameter(torch.zeros(4, 5), manifold=Lorentz(5))
    adam = RiemannianAdam([x], lr=0.02, betas=(0.8, 0.99), max_grad_norm=7.0, exact=True)
    cfg, m, v = adam._kernel_args(adam.param_groups[0], x)
    assert (cfg.kind, cfg.exact, cfg.has_clip, cfg.step, cfg.zero_grad) == (L.GM_OPT_RADAM, 1, 1, 1, 0)
    assert (cfg.lr, cfg.beta1, cfg.beta2, cfg.max_grad_norm, cfg.eps) == (0.02, 0.8, 0.99, 7.0, 1e-8)
    assert m.shape == x.shape and v.shape == x.shape and float(m.abs().sum() + v.abs().sum()) == 0.0
    adam._advance(x, 3)
    assert adam.state[x]['step'] == 4 and adam._kernel_args(adam.param_groups[0], x)[0].step == 4
    nc = RiemannianAdam([x], lr=0.02, nc=True)
    nc._kernel_args(nc.param_groups[0], x)
    nc._advance(x, 4)
    assert nc._kernel_args(nc.param_groups[0], x)[0].beta2 == pytest.approx(1 - 1 / 5)  # AdamNc: 1 - 1/t
    sgd = RiemannianSGD([x], lr=0.1, momentum=0.9, dampening=0.2)
    cfg, buf, none = sgd._kernel_args(sgd.param_groups[0], x)
    assert (cfg.kind, cfg.has_momentum, cfg.first_step, cfg.has_clip, none) == (L.GM_OPT_RSGD, 1, 1, 0, None)
    assert (cfg.momentum, cfg.dampening, cfg.lr) == (0.9, 0.2, 0.1) and buf.shape == x.shape
    assert sgd._kernel_args(sgd.param_groups[0], x)[0].first_step == 0  # the buffer exists from now on
    plain = RiemannianSGD([x], lr=0.1)
    cfg, buf, _ = plain._kernel_args(plain.param_groups[0], x)
    assert cfg.has_momentum == 0 and buf is None
    with pytest.raises(ValueError):
        RiemannianSGD([x], lr=0.1, momentum=-1.0)


def test_products_embedding_structure_without_gpu():
    """products.Embedding (products/embedding.py:8-61): Universal factors as sub-modules, curvature parameters, the
    state-dict keys agg_grid_results.py reads, and the r_max norm constraint of stabilize() -- tensor-only parts."""
    from graphembed.products import Embedding
    emb = Embedding.__new__(Embedding)
    torch.nn.Module.__init__(emb)
    from graphembed.manifolds import Universal
    from graphembed.modules import ManifoldParameter
    emb.n, emb.ds, emb.r_max = 6, [3, 2], 5.0
    emb.manifolds = torch.nn.ModuleList([Universal(d, c_init=c) for d, c in zip(emb.ds, (0.4, -0.6))])
    emb.xs = torch.nn.ParameterList([ManifoldParameter(torch.randn(6, d), manifold=m)
                                     for d, m in zip(emb.ds, emb.manifolds)])
    assert sorted(emb.state_dict().keys()) == ['manifolds.0.c', 'manifolds.1.c', 'xs.0', 'xs.1']
    assert [p.item() for p in emb.curvature_params] == pytest.approx([0.4, -0.6])
    emb.burnin(True)
    assert not any(p.requires_grad for p in emb.curvature_params)
    emb.burnin(False)
    assert all(p.requires_grad for p in emb.curvature_params)
    assert len(emb) == 6 and emb.fused_pair_kernels and not hasattr(emb, 'scales')
    assert emb.manifolds[0].get_c().item() == pytest.approx(0.401) and emb.manifolds[1].get_c().item() == pytest.approx(-0.601)


def test_fast_precision_surface_and_csr_construction():
    """graphembed.pyx exposes the reference's class under both of its names; edges_to_csr builds the symmetric CSR
    the BFS and rank kernels expect (sorted neighbours, no duplicates or self loops)."""
    import graphembed.pyx as pyx
    from graphembed.data import edges_to_csr
    assert pyx.PyFastPrecision is pyx.FastPrecision
    for name in ('mean_average_precision', 'layer_mean_f1_scores', 'layer_mean_average_f1_scores', 'nodes_per_layer'):
        assert callable(getattr(pyx.FastPrecision, name))
    rowptr, colidx = edges_to_csr(5, np.array([[0, 1], [1, 0], [1, 2], [3, 3], [4, 1], [1, 2]]))
    assert rowptr.tolist() == [0, 1, 4, 5, 5, 6] and colidx.tolist() == [1, 0, 2, 4, 1, 1]
    rowptr, colidx = edges_to_csr(3, np.array([[0, 1], [1, 2]]), directed=True)  # in-neighbours of every node
    assert rowptr.tolist() == [0, 0, 1, 2] and colidx.tolist() == [0, 1]


def test_softplus_cache_dropped_when_a_riemannian_optimizer_steps_the_scale(monkeypatch):
    """RiemannianAdam / RiemannianSGD update parameters through a raw-pointer kernel write: neither `_version` nor
    `data_ptr()` changes, so fused_step itself must invalidate the cached softplus(scale) (round-1 advisor finding:
    with run_grid.py's `RiemannianAdam([{params: xs}, {params: emb.curvature_params}])` the kernels kept using
    softplus(0.5) while the scale drifted)."""
    from graphembed import _ops
    from graphembed.modules import _softplus_value
    from graphembed.optim import RiemannianSGD

    def fake_optim_step(spec, cfg, x, grad, buf1, buf2):  # what the kernel does: write through the pointer
        ptr_before = x.data_ptr()
        x.view(-1).numpy()[...] -= cfg.lr * grad.view(-1).numpy()
        assert x.data_ptr() == ptr_before

    monkeypatch.setattr(_ops, 'optim_step', fake_optim_step)
    from graphembed import _torch_ops
    monkeypatch.setattr(_torch_ops, 'optim_step', fake_optim_step)  # (the registered op has no CPU kernel)
    s = torch.nn.Parameter(torch.tensor(0.5, dtype=torch.float64))
    v0 = _softplus_value(s)
    assert _softplus_value(s) == v0 and s._gm_softplus is not None
    opt = RiemannianSGD([s], lr=0.1)
    s.grad = torch.tensor(2.0, dtype=torch.float64)
    version = s._version
    opt.step()
    assert s._version == version  # the update really is invisible to torch's version counter
    assert abs(float(s) - 0.3) < 1e-12
    assert abs(_softplus_value(s) - float(torch.nn.functional.softplus(torch.tensor(0.3, dtype=torch.float64)))) < 1e-15


def test_sampler_oracle_is_counter_based_and_uniform():
    """oracle/sampler_oracle.py: the draw depends only on (seed, k, i, N); j != i; roughly uniform; known answers of
    the splitmix64 output function (Vigna's reference: seed 0 -> first outputs 0xE220A8397B1DCDAF, 0x6E789E6AA1B965F4)."""
    import sampler_oracle as S
    z = S.hash32(0, np.arange(2))
    assert [int(v) for v in z] == [0xE220A8397B1DCDAF >> 32, 0x6E789E6AA1B965F4 >> 32]
    N, G, per = 1000, 7, 4096
    rng = np.random.RandomState(0)
    levels = rng.randint(1, 9, size=(G, N)).astype(np.uint8)
    src = rng.choice(N, G, replace=False)
    I, J, H = S.sample_pairs(src, levels, per, seed=1234)
    assert I.shape == (G * per,) and (I.reshape(G, per) == src[:, None]).all()
    assert (J != I).all() and J.min() >= 0 and J.max() < N
    assert (H == levels[np.repeat(np.arange(G), per), J]).all()
    # a prefix / another grouping of the same stream gives the same pairs
    I2, J2, H2 = S.sample_pairs(src, levels, per, seed=1234, P=5000)
    assert (J2 == J[:5000]).all() and (H2 == H[:5000]).all()
    counts = np.bincount(J, minlength=N)
    assert counts.max() < 3 * counts.mean() and (counts > 0).mean() > 0.99
    assert (S.sample_pairs(src, levels, per, seed=1235)[1] != J).mean() > 0.99


def test_custom_ops_are_registered_and_have_no_cpu_kernel():
    """torch.ops.graphembed_b200.*: schema-checked custom ops over the C-ABI, CUDA only (no CPU fallback)."""
    import graphembed  # noqa: F401
    from graphembed import _torch_ops  # noqa: F401
    ns = torch.ops.graphembed_b200
    for name in ('pair_dist2', 'pairs_loss_fused', 'optim_step', 'bfs_levels'):
        assert hasattr(ns, name)
    schema = str(ns.pairs_loss_fused.default._schema)
    assert 'Tensor(a!) grad' in schema and 'Tensor(b!) acc' in schema
    with pytest.raises(NotImplementedError):
        ns.pair_dist2(torch.zeros(3, 2, 2), torch.zeros(1, dtype=torch.int64), torch.zeros(1, dtype=torch.int64), 0, 2,
                      0, 0, 1e-8, 1e8)
    with pytest.raises(NotImplementedError):
        ns.bfs_levels(torch.zeros(3, dtype=torch.int32), torch.zeros(2, dtype=torch.int32),
                      torch.zeros(1, dtype=torch.int32))


def test_cyclic_row_shards_round_trip():
    """graphembed.parallel.cyclic_shard / cyclic_unshard: rank r of `world` holds rows r, r + world, ... (the ownership
    rule of gm_row_shards_t: global row v = row v // world of shard v % world)."""
    import torch
    from graphembed.parallel import cyclic_shard, cyclic_unshard
    x = torch.arange(22 * 3, dtype=torch.float64).reshape(22, 3)
    for world in (1, 2, 4, 8):
        shards = [cyclic_shard(x, r, world) for r in range(world)]
        for r, s in enumerate(shards):
            assert s.is_contiguous()
            for k in range(s.shape[0]):
                v = k * world + r
                assert torch.equal(s[k], x[v]) and v % world == r and v // world == k
        assert torch.equal(cyclic_unshard(shards), x)


def test_window_order_groups_by_target_window():
    import torch
    from graphembed.engine import window_order
    g = torch.Generator().manual_seed(0)
    n, P, W = 1000, 5000, 4
    j = torch.randint(n, (P,), generator=g, dtype=torch.int32)
    packed = j | (torch.randint(1, 9, (P,), generator=g, dtype=torch.int32) << 24)
    order = window_order(packed, n, W)
    assert torch.equal(order, window_order(j, n, W))  # a packed hop count is ignored
    w = (j[order].long() * W) // n
    assert bool((w[1:] >= w[:-1]).all())
    for k in range(W):  # stable: original order inside a window
        pos = order[w == k]
        assert bool((pos[1:] > pos[:-1]).all())


def test_pack_hops2_round_trip_on_the_host():
    """engine.pack_hops2: every group's targets sorted by row and stored as 13-bit gaps + 3 bits of hop count; decoding
    the words as include/gm_kernels.h states the format gives back the sorted batch."""
    import torch
    from graphembed.engine import pack_hops2
    g = torch.Generator().manual_seed(0)
    offs = torch.tensor([0, 400, 400, 900, 1300, 2000], dtype=torch.int64)
    J = torch.randint(20000, (2000,), generator=g, dtype=torch.int32)
    H = torch.randint(1, 9, (2000,), generator=g, dtype=torch.uint8)
    words, bases, order = pack_hops2(offs, J, H)
    assert words.dtype == torch.int16 and bases.dtype == torch.int32 and words.numel() == 2000
    wu = words.long() & 0xFFFF
    for gi in range(5):
        j = int(bases[gi])
        for k in range(int(offs[gi]), int(offs[gi + 1])):
            j += int(wu[k] & 0x1FFF)
            assert j == int(J[order[k]]) and (int(wu[k]) >> 13) + 1 == int(H[order[k]])
            assert int(offs[gi]) <= int(order[k]) < int(offs[gi + 1])
    J2 = J.clone()
    J2[0], J2[1:400] = 0, 19999
    assert pack_hops2(offs, J2, H) is None  # a gap that does not fit 13 bits


def test_window_groups_are_the_runs_of_the_reordered_batch():
    """engine.window_groups: after window_order the batch consists of (window, source) runs; the returned group rows and
    offsets describe exactly those runs (empty groups included), so expand_groups / pack_hops2 can take them as is."""
    import torch
    from graphembed.engine import window_groups
    g = torch.Generator().manual_seed(0)
    n_points, G, W = 1000, 7, 4
    counts = torch.tensor([50, 0, 13, 200, 1, 77, 64])
    offsets = torch.cat([torch.zeros(1, dtype=torch.int64), counts.cumsum(0)])
    src = torch.randperm(n_points, generator=g)[:G].int()
    P = int(offsets[-1])
    I = torch.repeat_interleave(src, counts)
    J = torch.randint(n_points, (P,), generator=g, dtype=torch.int32)
    order, rows, offs = window_groups(src, offsets, J, n_points, W)
    assert rows.numel() == W * G and offs.numel() == W * G + 1 and int(offs[-1]) == P
    assert torch.equal(torch.repeat_interleave(rows, offs[1:] - offs[:-1]), I[order])
    win = (J[order].long() * W) // n_points
    for k in range(W * G):
        a, b = int(offs[k]), int(offs[k + 1])
        assert bool((win[a:b] == k // G).all())
    packed = J | (torch.randint(1, 9, (P,), generator=g, dtype=torch.int32) << 24)  # a packed hop count is ignored
    o2, r2, f2 = window_groups(src, offsets, packed, n_points, W)
    assert torch.equal(o2, order) and torch.equal(f2, offs)
