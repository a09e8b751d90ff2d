// Deterministic gradient accumulation ("deterministic segmented reduction"; the reference's
// index_put_(accumulate=True) scatter of autograd, base.py:62-63 / SURVEY 3.3, is itself order-dependent on CUDA).
//
// gm_segment_sum adds rows of a (M, E) table in a FIXED order: chunk c is the sum, taken left to right, of
// src[order[k]] for chunk_start[c] <= k < chunk_end[c] and is stored (not accumulated) as row dst[c] of `out`.  Two
// calls give a reproducible scatter-add of per-pair gradient rows into the (N, E) gradient table: level 1 sums chunks
// of at most a fixed number of consecutive entries of every destination row's (sorted) incidence list into partial
// rows, level 2 sums each row's partials in order.  No atomics anywhere: same inputs, same bits, on every run and for
// every launch geometry.  One thread per (chunk, element): the E threads of a chunk read one contiguous row per step.
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/gm_kernels.h"

namespace gm {
void note_launch();
int check_launch();

template <typename T>
__global__ void __launch_bounds__(256)
segment_sum_kernel(int E, const T* __restrict__ src, const long long* __restrict__ order,
                   const long long* __restrict__ chunk_start, const long long* __restrict__ chunk_end,
                   const long long* __restrict__ dst, long long n_chunks, T* __restrict__ out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long c = t / E;
  if (c >= n_chunks) return;
  const int e = (int)(t - c * E);
  const long long k0 = chunk_start[c], k1 = chunk_end[c];
  T s = (T)0;
  for (long long k = k0; k < k1; ++k) {
    const long long r = order ? order[k] : k;
    s += src[r * E + e];
  }
  out[(dst ? dst[c] : c) * E + e] = s;
}
}  // namespace gm

extern "C" {
#pragma GCC visibility push(default)
int gm_segment_sum(int32_t dtype, int32_t E, const void* src, const int64_t* order, const int64_t* chunk_start,
                   const int64_t* chunk_end, const int64_t* dst, int64_t n_chunks, void* out, gm_stream_t stream) {
  if (dtype != GM_F32 && dtype != GM_F64) return GM_EINVAL;
  if (E < 1 || n_chunks < 0) return GM_EINVAL;
  if (n_chunks == 0) return GM_OK;
  if (!src || !chunk_start || !chunk_end || !out) return GM_ENULL;
  const long long threads = (long long)n_chunks * E;
  const long long blocks = (threads + 255) / 256;
  if (blocks > 0x7fffffffLL) return GM_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == GM_F32)
    gm::segment_sum_kernel<float><<<(unsigned)blocks, 256, 0, st>>>(E, (const float*)src, (const long long*)order,
                                                                    (const long long*)chunk_start,
                                                                    (const long long*)chunk_end, (const long long*)dst,
                                                                    n_chunks, (float*)out);
  else
    gm::segment_sum_kernel<double><<<(unsigned)blocks, 256, 0, st>>>(E, (const double*)src, (const long long*)order,
                                                                     (const long long*)chunk_start,
                                                                     (const long long*)chunk_end, (const long long*)dst,
                                                                     n_chunks, (double*)out);
  gm::note_launch();
  return gm::check_launch();
}
#pragma GCC visibility pop
}
