"""Small helpers shared by the drop-in package (same names as the reference's
graphembed/utils.py where user code imports them)."""
import glob
import logging
import math
import os
import re
import time

import torch

logger = logging.getLogger(__name__)

# the reference uses one epsilon for both precisions (graphembed/utils.py:13)
EPS = {torch.float32: 1e-8, torch.float64: 1e-8}


def nnm1d2_to_n(m):
    """n such that n(n-1)/2 == m."""
    n = (1 + math.isqrt(1 + 8 * m)) // 2
    if n * (n - 1) // 2 != m:
        raise AssertionError(f'{m} is not a triangular number n(n-1)/2')
    return n


def nnp1d2_to_n(m):
    """n such that n(n+1)/2 == m."""
    n = (math.isqrt(1 + 8 * m) - 1) // 2
    if n * (n + 1) // 2 != m:
        raise AssertionError(f'{m} is not a triangular number n(n+1)/2')
    return n


def triu_mask(n, m=None, *, d=0, device=None):
    """Boolean mask of the entries on/above the d-th diagonal of an n x m matrix."""
    m = m or n
    r = torch.arange(n, device=device).unsqueeze(1)
    c = torch.arange(m, device=device).unsqueeze(0)
    return (c - r) >= d


def _squareform(v, diag_offset):
    if v.ndim >= 2 and v.shape[-1] == v.shape[-2]:  # matrix -> vector
        n = v.shape[-1]
        i, j = torch.triu_indices(n, n, diag_offset, device=v.device)
        return v[..., i, j]
    n = nnm1d2_to_n(v.shape[-1]) if diag_offset == 1 else nnp1d2_to_n(v.shape[-1])
    i, j = torch.triu_indices(n, n, diag_offset, device=v.device)
    out = v.new_zeros(v.shape[:-1] + (n, n))
    out[..., i, j] = v
    out[..., j, i] = v
    return out


def squareform1(v):
    """scipy.spatial.distance.squareform for tensors (zero diagonal), both directions."""
    return _squareform(v, 1)


def squareform0(v):
    """Like squareform1 but the vector form includes the diagonal."""
    return _squareform(v, 0)


def basename_numeric_order(path):
    return [int(s) for s in re.findall(r'\d+', os.path.basename(path))][-1]


def latest_path_by_basename_numeric_order(pattern):
    paths = glob.glob(pattern)
    return max(paths, key=basename_numeric_order) if paths else None


def check_mkdir(path, increment=False):
    if not os.path.isdir(path):
        os.makedirs(path)
        return path
    if not increment:
        logger.warning('The given path already exists (%s)', path)
        return path
    k = 0
    while os.path.isdir(path):
        head, base = os.path.split(path)
        parts = base.split('_')
        if parts[-1].isdigit():
            base = '_'.join(parts[:-1])
        path = os.path.join(head, f'{base}_{k}')
        k += 1
    os.makedirs(path)
    logger.info('Created the directory (%s) instead', path)
    return path


class Timer:
    """Context manager logging wall time; synchronises the GPU so the number is real."""

    def __init__(self, msg, precision=4, loglevel=logging.DEBUG):
        self.msg, self.precision, self.loglevel = msg, precision, loglevel

    def __enter__(self):
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        self.start = time.time()
        return self

    def __exit__(self, *exc):
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        self.end = time.time()
        self.interval = self.end - self.start
        logger.log(self.loglevel, 'time(%s): %.*fs', self.msg, self.precision, self.interval)
