"""Shared pieces of the ranking-metric tests: fixtures from tests/golden/precision_map.npz (the reference's
own Python mAP on the reference test's inputs) and CSR adjacency construction."""
import os

import numpy as np

from helpers import GOLDEN

TAGS = ('a', 'b', 'c', 'd')


def load_precision_golden():
    with np.load(os.path.join(GOLDEN, 'precision_map.npz')) as z:
        return {k: z[k] for k in z.files}


def csr_of(n, edges):
    """Symmetric CSR (rowptr, colidx) of an undirected edge list -- numpy only (the product's edges_to_csr is checked
    against this in the GPU tests)."""
    e = np.asarray(edges, dtype=np.int64)
    src = np.concatenate([e[:, 0], e[:, 1]])
    dst = np.concatenate([e[:, 1], e[:, 0]])
    order = np.lexsort((dst, src))
    src, dst = src[order], dst[order]
    rowptr = np.zeros(n + 1, dtype=np.int32)
    rowptr[1:] = np.cumsum(np.bincount(src, minlength=n))
    return rowptr, dst.astype(np.int32)


def nx_graph(n, edges):
    import networkx as nx
    g = nx.Graph()
    g.add_nodes_from(range(n))
    g.add_edges_from(np.asarray(edges).tolist())
    return g
