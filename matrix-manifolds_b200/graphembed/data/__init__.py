from .dataset import GraphDataset
from .graph import load_graph_pdists, compute_graph_pdists, bfs_levels, edges_to_csr

__all__ = ['GraphDataset', 'load_graph_pdists', 'compute_graph_pdists', 'bfs_levels', 'edges_to_csr']
