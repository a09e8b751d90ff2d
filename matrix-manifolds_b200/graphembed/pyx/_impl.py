"""The native side of graphembed.pyx lives in matrix-manifolds_b200/csrc/gm_rank.cu (gm_rank_metrics) and
csrc/gm_graph.cu (gm_bfs_multi_source); the reference keeps it under graphembed/pyx/impl/precision.{hpp,cpp}."""
