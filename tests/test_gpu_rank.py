"""GPU parity tests (-m gpu) for the ranking metrics: graphembed.pyx.FastPrecision on the GPU
(gm_bfs_multi_source + gm_rank_metrics) against the pinned C oracle (oracle/precision_oracle.c), the reference's
Python mAP fixtures and the reference's known-answer test.  The counts are integers (exact); the F1 sums are fp64 sums
whose order differs between the two (atomics), hence 1e-11."""
import numpy as np
import pytest
import torch

from helpers_precision import TAGS, csr_of, load_precision_golden, nx_graph
from precision_oracle import FastPrecisionOracle

pytestmark = pytest.mark.gpu


def _pair(tag):
    from graphembed.pyx import FastPrecision
    g = load_precision_golden()
    n = int(g[f'{tag}_n'])
    return g, FastPrecision(nx_graph(n, g[f'{tag}_edges'])), FastPrecisionOracle(*csr_of(n, g[f'{tag}_edges']))


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
@pytest.mark.parametrize('tag', TAGS)
def test_f1_and_map_vs_oracle_and_reference_fixture(tag, dtype):
    g, fp, orc = _pair(tag)
    pd = g[f'{tag}_pdists'].astype(dtype)
    assert abs(fp.mean_average_precision(pd) - float(g[f'{tag}_map'])) < 1e-12  # the reference's own Python mAP
    for got, want in ((fp.layer_mean_f1_scores(pd), orc.layer_mean_f1_scores(pd)),
                      (fp.layer_mean_average_f1_scores(pd), orc.layer_mean_average_f1_scores(pd)),
                      (fp.layer_mean_f1_scores(pd, min_degree=3, max_degree=8),
                       orc.layer_mean_f1_scores(pd, min_degree=3, max_degree=8))):
        assert len(got[0]) == len(want[0])
        # a layer nobody passes the degree filter on is 0/0 = NaN in the reference too (precision.cpp:415-418)
        assert np.allclose(got[0], want[0], rtol=1e-11, atol=0, equal_nan=True)
        assert np.allclose(got[1], want[1], atol=1e-11, equal_nan=True)
        assert np.array_equal(np.isnan(got[0]), np.isnan(want[0]))
    assert np.array_equal(fp.nodes_per_layer(), orc.nodes_per_layer())


@pytest.mark.parametrize('tag', TAGS)
def test_f1_trivial_known_answer(tag):
    """tests/test_metrics.py:26-35 of the reference: the graph's own distances score F1 == 1 on every layer."""
    g, fp, _ = _pair(tag)
    means, stds = fp.layer_mean_f1_scores(g[f'{tag}_hops'])
    assert np.allclose(means, 1.0, atol=1e-6) and np.allclose(stds, 0.0, atol=1e-9)
    from graphembed.metrics import area_under_curve
    assert np.allclose(area_under_curve(means), 1, atol=1e-6)


def test_multiple_distance_sets_and_root_shards():
    g, fp, orc = _pair('b')
    n = fp.n
    rng = np.random.RandomState(3)
    pd2 = rng.rand(2 * fp.n_pdists)
    means, stds = fp.layer_mean_f1_scores(pd2, num_pdists_sets=2)
    # the reference pools both sets into the same per-layer moments (precision.cpp:400-412)
    m_a, s_a = orc.layer_mean_f1_scores(pd2[:fp.n_pdists])
    m_b, s_b = orc.layer_mean_f1_scores(pd2[fp.n_pdists:])
    assert np.allclose(means, 0.5 * (m_a + m_b), rtol=1e-11)
    assert np.allclose(stds, 0.5 * (s_a + m_a**2 + s_b + m_b**2) - means**2, atol=1e-11)
    # roots sharded over "ranks": the accumulators add up to the single-launch result
    whole, parts = fp._new_acc(), fp._new_acc()
    fp._accumulate(pd2[:fp.n_pdists], whole, 1, 99999)
    for lo, hi in ((0, n // 3), (n // 3, n // 3), (n // 3, n)):
        fp._accumulate(pd2[:fp.n_pdists], parts, 1, 99999, roots=(lo, hi))
    assert torch.equal(whole['f1_cnt'], parts['f1_cnt']) and torch.equal(whole['af_cnt'], parts['af_cnt'])
    assert torch.allclose(whole['f1'], parts['f1'], rtol=1e-12) and torch.allclose(whole['ap'], parts['ap'], rtol=1e-12)


def test_embedding_distances_on_a_larger_graph():
    """2000-node preferential-attachment graph: (a) manifold distances of an fp64 Lorentz embedding, (b) 2M pairwise
    DISTINCT fp32 values (std::sort leaves the order of ties unspecified, so tie-free inputs are what can be compared
    exactly): GPU vs oracle; and the size limit of the shared-memory sort is reported as GM_EUNSUPPORTED, not silently
    mis-sorted."""
    import networkx as nx
    from graphembed import _lib as L
    from graphembed.manifolds import Lorentz
    from graphembed.pyx import FastPrecision
    n = 2000
    g = nx.barabasi_albert_graph(n, 2, seed=1)
    fp = FastPrecision(g)
    orc = FastPrecisionOracle(*csr_of(n, np.array(g.edges())))
    torch.manual_seed(0)
    man = Lorentz(6)
    x = man.rand(n, out=torch.empty(0, device='cuda', dtype=torch.float64), ir=1.0)
    pd64 = man.pdist(x).contiguous()
    P = n * (n - 1) // 2
    pd32 = ((torch.randperm(P, generator=torch.Generator().manual_seed(1)) + 1).double() / P).float().cuda()
    assert pd32.unique().numel() == P
    for pd in (pd64, pd32):
        pd_host = pd.cpu().numpy()
        got, want = fp.layer_mean_f1_scores(pd), orc.layer_mean_f1_scores(pd_host)
        assert np.allclose(got[0], want[0], rtol=1e-10) and np.allclose(got[1], want[1], atol=1e-10)
        got, want = fp.layer_mean_average_f1_scores(pd), orc.layer_mean_average_f1_scores(pd_host)
        assert np.allclose(got[0], want[0], rtol=1e-10) and np.allclose(got[1], want[1], atol=1e-10)
        assert abs(fp.mean_average_precision(pd) - orc.mean_average_precision(pd_host)) < 1e-12
    z = torch.zeros(4, dtype=torch.float64, device='cuda')
    zi = torch.zeros(4, dtype=torch.int64, device='cuda')
    rc = L.lib().gm_rank_metrics(L.GM_F64, L.ptr(z), L.ptr(zi), 20000, 0, 1, 1, 9, 5, L.ptr(z), L.ptr(z), L.ptr(zi),
                                 L.ptr(z), L.ptr(z), L.ptr(zi), L.ptr(z), None)
    assert rc == -2  # 20000 fp64 keys do not fit one SM's shared memory


@pytest.mark.parametrize('name', ['f32', 'f64'])
@pytest.mark.parametrize('tag', TAGS)
def test_f1_vs_reference_native_fixture(tag, name):
    """gm_rank_metrics against the outputs of the reference's own native FastPrecision on random distances
    (tests/golden/precision_f1.npz, from graphembed/pyx/impl/precision.cpp compiled unmodified -- oracle/Makefile)."""
    import os
    from helpers import GOLDEN
    from graphembed.pyx import FastPrecision
    g = load_precision_golden()
    with np.load(os.path.join(GOLDEN, 'precision_f1.npz')) as z:
        f = {k: z[k] for k in z.files}
    n = int(g[f'{tag}_n'])
    fp = FastPrecision(nx_graph(n, g[f'{tag}_edges']))
    pd = g[f'{tag}_pdists'].astype(np.float32) if name == 'f32' else f[f'{tag}_pd64']
    assert np.array_equal(fp.nodes_per_layer(), f[f'{tag}_npl'])
    assert abs(fp.mean_average_precision(pd) - float(f[f'{tag}_{name}_map'])) < 1e-12
    for got, key in ((fp.layer_mean_f1_scores(pd), 'f1'), (fp.layer_mean_f1_scores(pd, min_degree=3, max_degree=8), 'f1w'),
                     (fp.layer_mean_average_f1_scores(pd), 'af')):
        assert np.allclose(got[0], f[f'{tag}_{name}_{key}_mean'], rtol=1e-11, atol=0, equal_nan=True)
        assert np.allclose(got[1], f[f'{tag}_{name}_{key}_std'], atol=1e-11, equal_nan=True)
