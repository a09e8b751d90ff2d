"""`graphembed.pyx.FastPrecision` -- same constructor and methods as the reference's Cython class
(graphembed/pyx/precision.pyx:48-127 over graphembed/pyx/impl/precision.cpp), computed on the GPU:
the shortest-path-tree layers come from the multi-source BFS kernel (gm_bfs_multi_source) and the ranking statistics
from gm_rank_metrics (csrc/gm_rank.cu): one CTA per root sorts that root's manifold distances in shared memory and
turns the reference's ordered-multiset walk into prefix counts.  Unweighted graphs only (the hot path's targets are
hop counts); there is no CPU fallback."""
import numpy as np
import torch

from . import _impl  # noqa: F401  (keeps `graphembed.pyx._impl` importable for symmetry with the reference layout)
from .. import _lib as L
from ..data.graph import bfs_levels, edges_to_csr


class FastPrecision:

    def __init__(self, g, device='cuda'):
        self.n = g.number_of_nodes()
        self.n_pdists = self.n * (self.n - 1) // 2
        if g.number_of_edges() and 'weight' in list(g.edges(data=True))[0][2]:
            raise NotImplementedError('weighted graphs (Dijkstra layers) are outside the GPU path')
        rowptr, colidx = edges_to_csr(self.n, np.array(g.edges(), dtype=np.int64), directed=g.is_directed())
        self.device = torch.device(device)
        self.levels = bfs_levels(rowptr, colidx, device=self.device, level_bytes=1)  # (n, n) uint8 hop counts
        deepest = int(self.levels.max().item())
        if deepest == 255:
            raise NotImplementedError('disconnected graphs / more than 254 layers are outside the GPU path')
        self.n_layers = deepest + 1  # max_num_layers_ (precision.cpp:191-199)

    # ---- one launch: every accumulator of precision.cpp for one distance set ------------------------------------
    def _accumulate(self, mpdists, acc, min_degree, max_degree, roots=None):
        mp = torch.as_tensor(mpdists)
        if mp.dtype not in (torch.float32, torch.float64):
            mp = mp.double()
        mp = mp.to(self.device).contiguous()
        lo, hi = (0, self.n) if roots is None else roots
        k = self.n_layers - 1
        with torch.cuda.device(self.device):
            rc = L.lib().gm_rank_metrics(L.dtype_code(mp.dtype), L.ptr(mp), L.ptr(self.levels), self.n, lo, hi,
                                         int(min(min_degree, 2**31 - 1)), int(min(max_degree, 2**31 - 1)),
                                         self.n_layers, L.ptr(acc['f1'][0]), L.ptr(acc['f1'][1]), L.ptr(acc['f1_cnt']),
                                         L.ptr(acc['af'][0]), L.ptr(acc['af'][1]), L.ptr(acc['af_cnt']),
                                         L.ptr(acc['ap']), L.stream_ptr(self.device))
        L.check(rc, 'gm_rank_metrics')
        return k

    def _new_acc(self):
        k = max(self.n_layers - 1, 1)
        z = lambda *s, dt=torch.float64: torch.zeros(*s, dtype=dt, device=self.device)  # noqa: E731
        return dict(f1=z(2, k), f1_cnt=z(k, dt=torch.int64), af=z(2, k), af_cnt=z(k, dt=torch.int64), ap=z(1))

    def _run(self, mpdists, num_pdists_sets, min_degree=1, max_degree=99999):
        mp = torch.as_tensor(mpdists)
        assert mp.numel() == self.n_pdists * num_pdists_sets
        acc = self._new_acc()
        for s in range(num_pdists_sets):
            self._accumulate(mp.reshape(num_pdists_sets, self.n_pdists)[s], acc, min_degree, max_degree)
        return acc

    @staticmethod
    def _stats(m, cnt):
        cnt = cnt.double()
        means = m[0] / cnt
        stds = m[1] / cnt - means * means  # (sic) a variance, as the reference returns it (precision.cpp:415-418)
        return means.cpu().numpy(), stds.cpu().numpy()

    # ---- the reference's API (pyx/precision.pyx) ---------------------------------------------------------------------
    def mean_average_precision(self, mpdists):
        return float(self._run(mpdists, 1)['ap'].item()) / self.n

    def layer_mean_f1_scores(self, mpdists, num_pdists_sets=1, min_degree=1, max_degree=99999):
        acc = self._run(mpdists, num_pdists_sets, min_degree, max_degree)
        return self._stats(acc['f1'], acc['f1_cnt'])

    def layer_mean_average_f1_scores(self, mpdists, num_pdists_sets=1):
        acc = self._run(mpdists, num_pdists_sets)
        return self._stats(acc['af'], acc['af_cnt'])

    def nodes_per_layer(self):
        return torch.bincount(self.levels.reshape(-1).long(), minlength=self.n_layers).to(torch.int32).cpu().numpy()


PyFastPrecision = FastPrecision
