"""GPU end-to-end tests (-m gpu) of the experiment driver matrix-manifolds_b200/run.py with the reference's YAML schema
(graphembed/example_config.yaml, graphembed/run.py:20-135): edge list -> BFS targets (+ `.cached_pdists` file) ->
training -> checkpoints and the Layer_Mean_F1 lazy metric files, for (a) the reference's example config in miniature and
(b) a products.Embedding (Universal factors) config with a curvature optimizer.  Written files follow
train.py:331-347 / train.py:278-280 so the reference's aggregation scripts can read a run directory unchanged."""
import gzip
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

EXAMPLE = """
save_dir_root: '{save}/run'
cache_dir: '{cache}'
input_graph: '{graph}'
embedding:
  closure:
    name: graphembed.modules.ManifoldEmbedding
    params:
      manifolds:
        - object:
            name: graphembed.manifolds.SymmetricPositiveDefinite
            params:
              n: 2
              use_stein_div: False
objective_fn:
  closure:
    name: graphembed.objectives.KLDiveregenceLoss
    params:
      inference_model: 'sne'
      inclusive: True
training_params:
  alpha: 10.0
  n_epochs: 6
  batch_size: null
  stabilize_every_epochs: 2
  val_every_epochs: 3
  save_metrics_every_epochs: 3
embedding_optimizer:
  closure:
    name: graphembed.optim.RiemannianAdam
    params:
      lr: 0.01
      max_grad_norm: 100
      exact: False
"""

PRODUCT = """
save_dir_root: '{save}/run'
cache_dir: '{cache}'
input_graph: '{graph}'
embedding:
  closure:
    name: graphembed.products.Embedding
    params:
      ds: [3, 2]
      c_init: 0.3
objective_fn:
  closure:
    name: graphembed.objectives.QuotientLoss
training_params:
  alpha: 1.0
  n_epochs: 6
  val_every_epochs: 3
  save_metrics_every_epochs: 3
embedding_optimizer:
  closure:
    name: graphembed.optim.RiemannianAdam
    params:
      lr: 0.01
      max_grad_norm: 100
      exact: True
curvature_optimizer:
  closure:
    name: torch.optim.SGD
    params:
      lr: 0.0001
"""


def _write_tree(path):
    import networkx as nx
    g = nx.balanced_tree(3, 4)  # 121 nodes
    with gzip.open(path, 'wt') as f:
        for u, v in g.edges():
            f.write(f'{u} {v}\n')
    return g


def _run(tmp_path, template):
    sys.path.insert(0, os.path.join(ROOT, 'matrix-manifolds_b200'))
    import run as driver
    graph = str(tmp_path / 'btree121.edges.gz')
    g = _write_tree(graph)
    cfg = tmp_path / 'config.yaml'
    cfg.write_text(template.format(save=tmp_path, cache=tmp_path / 'cache', graph=graph))
    prev_dtype = torch.get_default_dtype()
    try:
        engine = driver.main(['--config', str(cfg), '--random_seed', '42'])
    finally:
        torch.set_default_dtype(prev_dtype)
        torch.set_default_device('cpu')
    return g, engine, tmp_path / 'run_1' if (tmp_path / 'run_1').exists() else tmp_path / 'run'


def _check_run_dir(tmp_path, g, engine):
    from scipy.sparse.csgraph import shortest_path
    import networkx as nx
    save_dir = engine.save_dir
    files = set(os.listdir(save_dir))
    assert {'config.yaml', 'best_embedding.pth'} <= files
    assert any(f.startswith('best_loss_') for f in files)
    assert {'mean_Layer_Mean_F1_3.npy', 'std_Layer_Mean_F1_3.npy', 'mean_Layer_Mean_F1_6.npy'} <= files
    means = np.load(os.path.join(save_dir, 'mean_Layer_Mean_F1_6.npy'))
    assert means.shape == (8,) and np.all((means > 0) & (means <= 1))  # tree of depth 4: diameter 8
    # the cache holds the condensed BFS distances in the reference's format (float64, squareform order)
    cached = np.load(tmp_path / 'cache' / 'btree121.edges.gz' / 'cached_pdists.npy')
    hops = shortest_path(nx.to_scipy_sparse_array(g, nodelist=sorted(g.nodes())), unweighted=True)
    # run.py relabels nodes in read order (nx.convert_node_labels_to_integers); the multiset of distances is invariant
    assert cached.dtype == np.float64 and cached.shape == (121 * 120 // 2,)
    assert np.array_equal(np.sort(cached), np.sort(hops[np.triu_indices(121, 1)]))
    h = engine.writer.history
    assert len(h['AUC_Layer_Mean_F1']) == 2 and len(h['average_distortion']) == 2
    return h


def test_example_config_in_miniature(tmp_path):
    g, engine, _ = _run(tmp_path, EXAMPLE)
    h = _check_run_dir(tmp_path, g, engine)
    losses = [v for _, v in h['kl_loss']]
    assert len(losses) == 6 and all(np.isfinite(losses)) and losses[-1] < losses[0]
    sd = torch.load(os.path.join(engine.save_dir, 'best_embedding.pth'))
    assert sorted(sd) == ['scales.0', 'xs.0'] and sd['xs.0'].shape == (121, 2, 2)


def test_products_embedding_config_with_curvature_optimizer(tmp_path):
    from graphembed.products import TrainingEngine
    g, engine, _ = _run(tmp_path, PRODUCT)
    assert isinstance(engine, TrainingEngine) and engine.stabilize_every_epochs == 1
    h = _check_run_dir(tmp_path, g, engine)
    losses = [v for _, v in h['quotient_loss']]
    # no monotonicity claim: QuotientLoss' second term divides by m + 1/(epoch+1) (objectives.py:30), so from the default
    # init (m ~ 1e-4) it GROWS with the epoch until the points have spread out -- in the reference just the same
    assert len(losses) == 6 and all(np.isfinite(losses))
    assert all(np.isfinite(v) for _, v in h['average_distortion'])
    cs = [m.c.item() for m in engine.embedding.manifolds]
    assert all(abs(c - 0.3) > 1e-9 for c in cs)  # the curvature optimizer moved both curvatures
    assert [v for _, v in h['curv0']][-1] == pytest.approx(-engine.embedding.manifolds[0].get_c().item())
    sd = torch.load(os.path.join(engine.save_dir, 'best_embedding.pth'))
    assert sorted(sd) == ['manifolds.0.c', 'manifolds.1.c', 'xs.0', 'xs.1']
