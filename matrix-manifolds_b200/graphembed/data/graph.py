"""Graph loading and all-pairs hop-count targets.

`load_graph_pdists` keeps the contract of the reference's
graphembed/data/graph.py:15-63 (edge-list / .npy input, `.cached_pdists` cache,
returns (condensed distance tensor, networkx graph)); the distances themselves
come from the bit-parallel multi-source BFS kernel (csrc/gm_graph.cu) instead of
networkit's APSP plus an O(N^2) Python loop (graph.py:66-87)."""
import logging
import os

import numpy as np
import torch

from .. import _lib as L
from ..utils import Timer

CACHED_PDISTS_FILE = 'cached_pdists.npy'


def edges_to_csr(n, edges, directed=False):
    """CSR (rowptr, colidx) int32 numpy arrays of the *in*-neighbours of every node (for an undirected graph both
    directions of every edge), neighbours sorted, duplicates and self loops removed."""
    e = np.asarray(edges, dtype=np.int64).reshape(-1, 2)
    src, dst = e[:, 0], e[:, 1]
    if not directed:
        src, dst = np.concatenate([src, dst]), np.concatenate([dst, src])
    keep = src != dst
    src, dst = src[keep], dst[keep]
    key = np.sort(dst * n + src)  # row = dst (who can be reached), col = src
    if key.size:
        key = key[np.concatenate(([True], key[1:] != key[:-1]))]  # drop duplicate edges
    rows, cols = key // n, key % n
    rowptr = np.zeros(n + 1, dtype=np.int64)
    rowptr[1:] = np.cumsum(np.bincount(rows, minlength=n))
    return rowptr.astype(np.int32), cols.astype(np.int32)


_LEVEL_DTYPES = {1: torch.uint8, 2: torch.int16, 4: torch.int32}


def bfs_levels(rowptr, colidx, sources=None, device='cuda', level_bytes=None):
    """(S, N) hop counts from each source (all nodes if None) on the GPU.  Unreachable = all-ones of the level
    type.  Starts with uint8 levels and retries with a wider type if the graph is deeper than 254 hops."""
    rowptr = torch.as_tensor(rowptr, dtype=torch.int32, device=device).contiguous()
    colidx = torch.as_tensor(colidx, dtype=torch.int32, device=device).contiguous()
    n = rowptr.numel() - 1
    if sources is None:
        sources = torch.arange(n, dtype=torch.int32, device=device)
    sources = torch.as_tensor(sources, dtype=torch.int32, device=device).contiguous()
    s = sources.numel()
    L.require_cuda(rowptr)
    ws = torch.empty(L.lib().gm_bfs_workspace_bytes(n, s), dtype=torch.uint8, device=device)
    for nbytes in ((level_bytes,) if level_bytes else (1, 2, 4)):
        levels = torch.empty((s, n), dtype=_LEVEL_DTYPES[nbytes], device=device)
        with torch.cuda.device(rowptr.device):
            rc = L.lib().gm_bfs_multi_source(L.ptr(rowptr), L.ptr(colidx), n, L.ptr(sources), s, nbytes,
                                             L.ptr(levels), L.ptr(ws), ws.numel(), L.stream_ptr(rowptr.device))
        if rc == -2 and not level_bytes:  # GM_EUNSUPPORTED: deeper than this level type can hold
            continue
        L.check(rc, 'gm_bfs_multi_source')
        return levels
    raise RuntimeError('BFS did not terminate within 65535 levels')


def levels_to_condensed(levels, dtype=torch.float32):
    """scipy.squareform-order vector of the strict upper triangle of a full (N, N) level matrix."""
    n = levels.shape[0]
    assert levels.shape == (n, n)
    out = torch.empty(n * (n - 1) // 2, dtype=dtype, device=levels.device)
    with torch.cuda.device(levels.device):
        rc = L.lib().gm_levels_to_condensed(levels.element_size(), L.ptr(levels), n, L.dtype_code(dtype), L.ptr(out),
                                            L.stream_ptr(levels.device))
    L.check(rc, 'gm_levels_to_condensed')
    return out


def compute_graph_pdists(g, cache_dir=None, device='cuda'):
    """Condensed hop-count distances of a networkx graph with nodes 0..n-1 (float64 numpy array)."""
    n = g.number_of_nodes()
    if n and 'weight' in (list(g.edges(data=True))[0][2] if g.number_of_edges() else {}):
        raise NotImplementedError('weighted graphs need Dijkstra; only unweighted BFS targets are on the hot path')
    rowptr, colidx = edges_to_csr(n, np.array(g.edges(), dtype=np.int64), directed=g.is_directed())
    levels = bfs_levels(rowptr, colidx, device=device)
    pd = levels_to_condensed(levels, torch.float64).cpu().numpy()
    if cache_dir and os.path.isdir(cache_dir):
        np.save(os.path.join(cache_dir, CACHED_PDISTS_FILE), pd)
    return pd


def load_graph_pdists(f, cache_dir=None, flip_probability=None, device='cuda'):
    import networkx as nx
    if flip_probability is not None:
        raise NotImplementedError('noisy-graph generation is a data-prep experiment outside the hot path')
    g = None
    f = os.path.abspath(os.path.realpath(f))
    if cache_dir is not None:
        cache_dir = os.path.join(cache_dir, os.path.basename(f))
    exts = ('.edges', '.dir-edges')
    if any(f.endswith(e) or f.endswith(e + '.gz') for e in exts):
        with Timer('graph loading', loglevel=logging.INFO):
            directed = '.dir-edges' in os.path.basename(f)
            g = nx.read_edgelist(f, create_using=nx.DiGraph if directed else nx.Graph)
        g = nx.convert_node_labels_to_integers(g)
        assert g.is_directed() or nx.number_connected_components(g) == 1
        if cache_dir and os.path.isdir(cache_dir):
            f = os.path.join(cache_dir, CACHED_PDISTS_FILE)
            assert os.path.isfile(f)
        else:
            if cache_dir:
                os.makedirs(cache_dir)
            with Timer('computing graph distances', loglevel=logging.INFO):
                pd = compute_graph_pdists(g, cache_dir, device=device)
            return torch.tensor(pd, dtype=torch.get_default_dtype()), g
    if f.endswith('npy'):
        with Timer('loading distances', loglevel=logging.INFO):
            return torch.tensor(np.load(f), dtype=torch.get_default_dtype()), g
    raise ValueError('Unrecognized input graph file: {}'.format(f))
