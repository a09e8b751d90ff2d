"""graphembed-b200: drop-in for the training hot path of dalab/matrix-manifolds'
`graphembed` package, backed by hand-written sm_100a CUDA kernels
(matrix-manifolds_b200/lib/libgm_b200.so, C-ABI in include/gm_kernels.h).

Same import paths as the reference for everything on the hot path:
graphembed.manifolds, graphembed.optim, graphembed.modules, graphembed.objectives,
graphembed.data, graphembed.train, graphembed.metrics, graphembed.utils.
"""
__version__ = '0.1.0'
