"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Host statement (numpy) of the on-device pair draw GM_PAIRS_SAMPLED
(include/gm_kernels.h; device code csrc/gm_launch.cuh::sampled_word).

The reference has no pair sampler (it enumerates all pairs of a node batch, train.py:206-213); sampled pairs are the
capability BASELINE.json's config 5 adds, so there is nothing in /root/reference to pin this against.  What IS pinned:
the definition below is counter based -- pair k of a step depends only on (seed, k, i, N) -- and the GPU tests check
that the kernels' draw equals it bit for bit (indices and hop counts), for any launch geometry.

    z = seed + (k + 1) * 0x9E3779B97F4A7C15          (mod 2^64)
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9;  z = (z ^ (z >> 27)) * 0x94D049BB133111EB;  z ^= z >> 31   (splitmix64)
    r = z >> 32;  j = (r * (N - 1)) >> 32;  j += (j >= i)
"""
import numpy as np

_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def hash32(seed, k):
    """splitmix64 output function of seed + (k + 1) * golden, top 32 bits; k: array of pair numbers."""
    with np.errstate(over='ignore'):
        z = np.uint64(seed) + (np.asarray(k, dtype=np.uint64) + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(32)).astype(np.uint64)


def sample_pairs(sources, levels, per_src, seed, slots=None, P=None):
    """(I int32, J int32, hops uint8) of one step: sources (G,) node ids, levels (S, N) uint8 hop matrix, slots (G,) rows
    of `levels` (default: g)."""
    sources = np.asarray(sources, dtype=np.int64)
    G, N = sources.size, levels.shape[1]
    P = G * per_src if P is None else P
    k = np.arange(P, dtype=np.uint64)
    g = (k // np.uint64(per_src)).astype(np.int64)
    i = sources[g]
    r = hash32(seed, k)
    j = ((r * np.uint64(N - 1)) >> np.uint64(32)).astype(np.int64)
    j = j + (j >= i)
    row = g if slots is None else np.asarray(slots, dtype=np.int64)[g]
    hops = levels[row, j]
    return i.astype(np.int32), j.astype(np.int32), hops.astype(np.uint8)
