"""ctypes binding of libgm_b200.so -- the C-ABI declared in include/gm_kernels.h.

There is no CPU fallback: if the shared library is missing, or an op is handed a
tensor that is not a contiguous CUDA float32/float64 tensor, a RuntimeError is
raised.  PyTorch is used only for device memory and streams.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# GM_B200_LIB lets an integrator (or an A/B experiment) point at another build of the same C-ABI
LIB_PATH = os.environ.get('GM_B200_LIB') or os.path.join(os.path.dirname(_HERE), 'lib', 'libgm_b200.so')

# ---- enums (include/gm_kernels.h) -------------------------------------------
GM_F32, GM_F64 = 0, 1
GM_SPD_AI, GM_SPD_STEIN, GM_LORENTZ, GM_SPHERE, GM_GRASSMANN, GM_EUCLIDEAN, GM_UNIVERSAL = range(7)
GM_FAST_EIG, GM_FAST_CHOL, GM_FAST_SVD = 1, 2, 4
GM_PAIRS_ELEMENTWISE, GM_PAIRS_LIST, GM_PAIRS_TRIU, GM_PAIRS_SAMPLED = 0, 1, 2, 3
GM_LOSS_QUOTIENT, GM_LOSS_STRESS = 0, 1
GM_TGT_VECTOR, GM_TGT_DENSE, GM_TGT_HOPS_U8, GM_TGT_HOPS_U16, GM_TGT_HOPS_PACKED = 0, 1, 2, 3, 4
GM_OPT_RSGD, GM_OPT_RADAM = 0, 1
(GM_OP_EXP, GM_OP_RETR, GM_OP_LOG, GM_OP_PROJU, GM_OP_PROJX, GM_OP_EGRAD2RGRAD, GM_OP_INNER, GM_OP_NORM2,
 GM_OP_TRANSP, GM_OP_RETR_QR, GM_OP_SPD_SQRTM) = range(11)

_ERRORS = {-1: 'GM_EINVAL (bad argument combination)', -2: 'GM_EUNSUPPORTED (no kernel compiled for this '
           'manifold size / dtype)', -3: 'GM_ENULL (required pointer is NULL)'}


class Manifold(ctypes.Structure):
    _fields_ = [('kind', ctypes.c_int32), ('dtype', ctypes.c_int32), ('n', ctypes.c_int32), ('p', ctypes.c_int32),
                ('flags', ctypes.c_uint32), ('reserved', ctypes.c_int32), ('wmin', ctypes.c_double),
                ('wmax', ctypes.c_double), ('c_dev', ctypes.c_void_p), ('c_grad', ctypes.c_void_p)]


class Pairs(ctypes.Structure):
    _fields_ = [('mode', ctypes.c_int32), ('idx64', ctypes.c_int32), ('P', ctypes.c_int64),
                ('idx_i', ctypes.c_void_p), ('idx_j', ctypes.c_void_p), ('B', ctypes.c_int64),
                ('nodes', ctypes.c_void_p), ('k0', ctypes.c_int64),
                # GM_PAIRS_SAMPLED (zero otherwise)
                ('levels', ctypes.c_void_p), ('slots', ctypes.c_void_p), ('n_nodes', ctypes.c_int64),
                ('per_src', ctypes.c_int64), ('seed', ctypes.c_uint64),
                # GM_PAIRS_LIST locality hint: the list is `segments` consecutive parts walked one after another
                ('segments', ctypes.c_int32), ('reserved', ctypes.c_int32)]


class Loss(ctypes.Structure):
    _fields_ = [('kind', ctypes.c_int32), ('inc_l1', ctypes.c_int32), ('inc_l2', ctypes.c_int32),
                ('reserved', ctypes.c_int32), ('alpha', ctypes.c_double), ('eps', ctypes.c_double)]


class Targets(ctypes.Structure):
    _fields_ = [('mode', ctypes.c_int32), ('reserved', ctypes.c_int32), ('data', ctypes.c_void_p),
                ('ld', ctypes.c_int64), ('max_sq', ctypes.c_double)]


class Optim(ctypes.Structure):
    _fields_ = [('kind', ctypes.c_int32), ('exact', ctypes.c_int32), ('has_clip', ctypes.c_int32),
                ('step', ctypes.c_int32), ('has_momentum', ctypes.c_int32), ('first_step', ctypes.c_int32),
                ('grassmann_retr_qr', ctypes.c_int32), ('zero_grad', ctypes.c_int32), ('lr', ctypes.c_double),
                ('beta1', ctypes.c_double), ('beta2', ctypes.c_double), ('momentum', ctypes.c_double),
                ('dampening', ctypes.c_double), ('max_grad_norm', ctypes.c_double), ('eps', ctypes.c_double)]


GM_MAX_PEERS, GM_PEER_HANDLE_BYTES = 8, 64
GM_PEER_FLAG_BYTES = (2 * GM_MAX_PEERS + 2) * 8


class Peers(ctypes.Structure):
    _fields_ = [('world', ctypes.c_int32), ('rank', ctypes.c_int32), ('row_lo', ctypes.c_int64),
                ('epoch', ctypes.c_uint64), ('x', ctypes.c_void_p * GM_MAX_PEERS),
                ('grad', ctypes.c_void_p * GM_MAX_PEERS), ('flags', ctypes.c_void_p * GM_MAX_PEERS),
                ('acc', ctypes.c_void_p * GM_MAX_PEERS), ('acc_out', ctypes.c_void_p), ('n_acc', ctypes.c_int32),
                ('reserved', ctypes.c_int32), ('gsum', ctypes.c_void_p)]




class RowShards(ctypes.Structure):
    _fields_ = [('world', ctypes.c_int32), ('reserved', ctypes.c_int32), ('x', ctypes.c_void_p * GM_MAX_PEERS),
                ('grad', ctypes.c_void_p * GM_MAX_PEERS)]


_vp, _i32, _i64, _dbl = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_double
_PROTOTYPES = {
    'gm_version': (ctypes.c_char_p, []),
    'gm_launch_count': (_i64, []),
    'gm_supported': (ctypes.c_int, [ctypes.POINTER(Manifold)]),
    'gm_pairs_dist2': (ctypes.c_int, [ctypes.POINTER(Manifold), _vp, _vp, ctypes.POINTER(Pairs), _vp, _vp]),
    'gm_pairs_grad': (ctypes.c_int, [ctypes.POINTER(Manifold), _vp, _vp, ctypes.POINTER(Pairs), _vp, _dbl, _vp, _vp,
                                     _vp]),
    'gm_pairs_loss_fused': (ctypes.c_int, [ctypes.POINTER(Manifold), _vp, ctypes.POINTER(Pairs),
                                           ctypes.POINTER(Targets), ctypes.POINTER(Loss), _dbl, _vp, _vp, _vp, _vp]),
    'gm_pairs_product_fused': (ctypes.c_int, [_i32, ctypes.POINTER(Manifold), ctypes.POINTER(_vp), ctypes.POINTER(Pairs),
                                              ctypes.POINTER(Targets), ctypes.POINTER(Loss), ctypes.POINTER(_dbl), _vp,
                                              ctypes.POINTER(_vp), _vp]),
    'gm_product_loss': (ctypes.c_int, [_i32, _i32, ctypes.POINTER(_vp), ctypes.POINTER(_dbl), ctypes.POINTER(Pairs),
                                       ctypes.POINTER(Targets), ctypes.POINTER(Loss), _i64, _vp, _vp, _vp]),
    'gm_pairs_metrics': (ctypes.c_int, [_i32, _i32, ctypes.POINTER(_vp), ctypes.POINTER(_dbl), ctypes.POINTER(Pairs),
                                        ctypes.POINTER(Targets), _i32, _vp, _vp]),
    'gm_sne_row_stats': (ctypes.c_int, [_i32, _i32, ctypes.POINTER(_vp), ctypes.POINTER(_dbl), _vp, _i64, _dbl, _i32,
                                        _vp, _vp]),
    'gm_sne_pair_terms': (ctypes.c_int, [_i32, _i32, ctypes.POINTER(_vp), ctypes.POINTER(_dbl), _vp, _i64, _dbl, _i32,
                                         _vp, _vp, _vp, _vp]),
    'gm_optim_step': (ctypes.c_int, [ctypes.POINTER(Manifold), ctypes.POINTER(Optim), _vp, _vp, _vp, _vp, _i64, _vp]),
    'gm_train_epoch': (ctypes.c_int, [ctypes.POINTER(Manifold), ctypes.POINTER(Optim), _vp, _vp, _vp, _vp, _i64, _vp, _i32,
                                      _i64, _i64, _i64, ctypes.POINTER(Targets), ctypes.POINTER(Loss), _dbl, _vp, _i64,
                                      ctypes.POINTER(_i64), _vp]),
    'gm_train_epoch_product': (ctypes.c_int, [_i32, ctypes.POINTER(Manifold), ctypes.POINTER(Optim), ctypes.POINTER(_vp),
                                              ctypes.POINTER(_vp), ctypes.POINTER(_vp), ctypes.POINTER(_vp), _i64, _vp, _i32,
                                              _i64, _i64, _i64, ctypes.POINTER(Targets), ctypes.POINTER(Loss),
                                              ctypes.POINTER(_dbl), ctypes.POINTER(_vp), _vp, _vp, _i64,
                                              ctypes.POINTER(_i64), _vp]),
    'gm_optim_step_peer': (ctypes.c_int, [ctypes.POINTER(Manifold), ctypes.POINTER(Optim), ctypes.POINTER(Peers), _vp,
                                          _vp, _i64, _vp]),
    'gm_pairs_loss_fused_sharded': (ctypes.c_int, [ctypes.POINTER(Manifold), ctypes.POINTER(RowShards),
                                                   ctypes.POINTER(Pairs), ctypes.POINTER(Targets), ctypes.POINTER(Loss),
                                                   _dbl, _vp, _vp, _vp]),
    'gm_peer_barrier': (ctypes.c_int, [ctypes.POINTER(Peers), _i32, _vp]),
    'gm_segment_sum': (ctypes.c_int, [_i32, _i32, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp]),
    'gm_peer_alloc': (ctypes.c_int, [ctypes.c_size_t, ctypes.POINTER(_vp)]),
    'gm_peer_free': (ctypes.c_int, [_vp]),
    'gm_peer_export': (ctypes.c_int, [_vp, ctypes.c_char_p]),
    'gm_peer_open': (ctypes.c_int, [ctypes.c_char_p, ctypes.POINTER(_vp)]),
    'gm_peer_close': (ctypes.c_int, [_vp]),
    'gm_point_op': (ctypes.c_int, [ctypes.POINTER(Manifold), _i32, _vp, _vp, _vp, _vp, _i64, _vp]),
    'gm_bfs_workspace_bytes': (ctypes.c_size_t, [_i32, _i32]),
    'gm_bfs_multi_source': (ctypes.c_int, [_vp, _vp, _i32, _vp, _i32, _i32, _vp, _vp, ctypes.c_size_t, _vp]),
    'gm_levels_to_condensed': (ctypes.c_int, [_i32, _vp, _i32, _i32, _vp, _vp]),
    'gm_levels_to_dense_targets': (ctypes.c_int, [_i32, _vp, _i32, _dbl, _i32, _vp, _vp]),
    'gm_gather_levels': (ctypes.c_int, [_i32, _vp, _i32, _vp, _vp, _i64, _vp, _vp]),
    'gm_expand_groups': (ctypes.c_int, [_vp, _vp, _i32, _vp, _i64, _vp]),
    'gm_unpack_pairs3': (ctypes.c_int, [_vp, _i64, _vp, _vp]),
    'gm_unpack_pairs2': (ctypes.c_int, [_vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp]),
    'gm_rank_metrics': (ctypes.c_int, [_i32, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp,
                                       _vp, _vp]),
}
EXPORTED_SYMBOLS = tuple(_PROTOTYPES)

_lib = None


def lib():
    """The loaded shared library (loads on first use; raises if it was not built)."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                f'{LIB_PATH} not found: build the sm_100a kernels first '
                '(`python -c "import __graft_entry__ as g; g.build()"` or `make -C matrix-manifolds_b200`). '
                'graphembed-b200 has no CPU / PyTorch fallback.')
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOTYPES.items():
            fn = getattr(handle, name)  # AttributeError if the .so is stale
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc, what):
    if rc == 0:
        return
    if rc < 0:
        raise RuntimeError(f'{what}: {_ERRORS.get(rc, rc)}')
    raise RuntimeError(f'{what}: CUDA error {rc} ({torch.cuda.get_device_name() if torch.cuda.is_available() else "no device"})')


def dtype_code(dtype):
    if dtype == torch.float32:
        return GM_F32
    if dtype == torch.float64:
        return GM_F64
    raise RuntimeError(f'gm_b200 kernels support float32/float64 only, got {dtype}')


def require_cuda(*tensors):
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError('gm_b200 kernels need CUDA tensors (there is no CPU fallback); got a tensor on '
                               f'{t.device}. Move the embedding to the GPU or set the default device to cuda.')


def ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def launch_count():
    return int(lib().gm_launch_count())
