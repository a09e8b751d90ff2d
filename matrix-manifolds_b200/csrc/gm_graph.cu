// Graph-distance targets: bit-parallel multi-source BFS on a CSR graph and the
// conversions GraphDataset / load_graph_pdists need.
//
// Replaces compute_graph_pdists (data/graph.py:66-87: networkit APSP followed by
// an O(N^2) Python copy loop and scipy squareform) and GraphDataset.__init__
// (data/dataset.py:9-13).  For an unweighted graph networkit's APSP is a BFS per
// node; the in-tree statement of that BFS is pyx/impl/precision.cpp:44-63.
// Integer results: bit-exact.
//
// Algorithm: level-synchronous "pull" BFS for all S sources at once.  Every node
// keeps W = ceil(S/64) 64-bit words of frontier / visited bits (node-major, so the
// W words of one neighbour are contiguous and the loads of a warp coalesce).
// Thread t = v*W + w ORs word w of every in-neighbour's frontier, masks the
// visited bits, and writes level L for each newly reached (source, v).
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/gm_kernels.h"

namespace gm {
void note_launch();
int check_launch();

typedef unsigned long long u64;

struct BfsWs {
  u64* frontier;
  u64* next;
  u64* visited;
  int* changed;  // changed[L] != 0 iff level L reached a new node; changed[0] = 1
};

static const int kMaxLevels = 1 << 16;

static size_t ws_bytes(int N, int S) {
  size_t W = ((size_t)S + 63) / 64;
  return 3 * W * (size_t)N * sizeof(u64) + (size_t)(kMaxLevels + 1) * sizeof(int) + 256;
}

static BfsWs carve(void* ws, int N, int S) {
  size_t W = ((size_t)S + 63) / 64;
  BfsWs b;
  char* p = (char*)ws;
  b.frontier = (u64*)p; p += W * N * sizeof(u64);
  b.next = (u64*)p; p += W * N * sizeof(u64);
  b.visited = (u64*)p; p += W * N * sizeof(u64);
  b.changed = (int*)p;
  return b;
}

template <typename L>
__device__ __forceinline__ L unreached();
template <> __device__ __forceinline__ unsigned char unreached<unsigned char>() { return 255; }
template <> __device__ __forceinline__ unsigned short unreached<unsigned short>() { return 65535; }
template <> __device__ __forceinline__ int unreached<int>() { return -1; }

template <typename L>
__global__ void bfs_seed_kernel(const int* __restrict__ sources, int S, int N, int W, u64* __restrict__ frontier,
                                u64* __restrict__ visited, L* __restrict__ levels) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  int v = sources[s];
  u64 bit = 1ull << (s & 63);
  atomicOr(&frontier[(size_t)v * W + (s >> 6)], bit);
  atomicOr(&visited[(size_t)v * W + (s >> 6)], bit);
  levels[(size_t)s * N + v] = (L)0;
}

template <typename L>
__global__ void __launch_bounds__(256)
bfs_level_kernel(const int* __restrict__ rowptr, const int* __restrict__ colidx, int N, int W, int level,
                 const u64* __restrict__ frontier, u64* __restrict__ next, u64* __restrict__ visited,
                 int* __restrict__ changed, L* __restrict__ levels, int S) {
  if (changed[level - 1] == 0) return;  // the search finished at an earlier level
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)N * W) return;
  int v = (int)(t / W);
  int w = (int)(t % W);
  u64 acc = 0;
  int e0 = rowptr[v], e1 = rowptr[v + 1];
  for (int e = e0; e < e1; ++e) acc |= frontier[(size_t)colidx[e] * W + w];
  u64 fresh = acc & ~visited[t];
  next[t] = fresh;
  if (fresh) {
    visited[t] |= fresh;
    changed[level] = 1;
    L lv = (L)level;
    while (fresh) {
      int b = __ffsll((long long)fresh) - 1;
      fresh &= fresh - 1;
      int s = w * 64 + b;
      if (s < S) levels[(size_t)s * N + v] = lv;
    }
  }
}

template <typename L>
static int bfs_run(const int* rowptr, const int* colidx, int N, const int* sources, int S, L* levels, void* ws,
                   cudaStream_t st) {
  const int W = (S + 63) / 64;
  BfsWs b = carve(ws, N, S);
  const size_t words = (size_t)W * N;
  cudaMemsetAsync(b.frontier, 0, 3 * words * sizeof(u64), st);
  cudaMemsetAsync(b.changed, 0, (size_t)(kMaxLevels + 1) * sizeof(int), st);
  cudaMemsetAsync(levels, 0xFF, (size_t)S * N * sizeof(L), st);
  int one = 1;
  cudaMemcpyAsync(b.changed, &one, sizeof(int), cudaMemcpyHostToDevice, st);
  bfs_seed_kernel<L><<<(S + 127) / 128, 128, 0, st>>>(sources, S, N, W, b.frontier, b.visited, levels);
  note_launch();
  const int chunk = 8;
  const long long max_level = sizeof(L) == 1 ? 254 : (sizeof(L) == 2 ? 65534 : (long long)kMaxLevels - 1);
  size_t blocks = (words + 255) / 256;
  if (blocks > 0x7fffffffULL) return GM_EINVAL;
  int level = 1;
  while (true) {
    for (int i = 0; i < chunk; ++i, ++level) {
      if (level > max_level || level >= kMaxLevels) return GM_EUNSUPPORTED;  // needs a wider level type
      bfs_level_kernel<L><<<(unsigned)blocks, 256, 0, st>>>(rowptr, colidx, N, W, level, b.frontier, b.next, b.visited,
                                                            b.changed, levels, S);
      note_launch();
      u64* tmp = b.frontier; b.frontier = b.next; b.next = tmp;
    }
    int flag = 0;
    cudaMemcpyAsync(&flag, b.changed + (level - 1), sizeof(int), cudaMemcpyDeviceToHost, st);
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return (int)e;
    if (!flag) break;
  }
  return check_launch();
}

// k-th (a<b) pair of the condensed (scipy.squareform) order
__device__ __forceinline__ void condensed_decode(long long k, long long B, long long& a, long long& b) {
  double tb = (double)(2 * B - 1);
  long long r = (long long)floor((tb - sqrt(tb * tb - 8.0 * (double)k)) * 0.5);
  if (r < 0) r = 0;
  if (r > B - 2) r = B - 2;
  while (r + 1 <= B - 2 && (r + 1) * (2 * B - r - 2) / 2 <= k) ++r;
  while (r > 0 && r * (2 * B - r - 1) / 2 > k) --r;
  a = r;
  b = k - r * (2 * B - r - 1) / 2 + r + 1;
}

template <typename L, typename T>
__global__ void condensed_kernel(const L* __restrict__ levels, int N, T* __restrict__ out, long long P) {
  long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= P) return;
  long long a, b;
  condensed_decode(k, N, a, b);
  out[k] = (T)levels[a * N + b];
}

template <typename L, typename T>
__global__ void dense_targets_kernel(const L* __restrict__ levels, long long total, T max_sq, T* __restrict__ dense) {
  long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= total) return;
  T h = (T)levels[k];
  dense[k] = (h * h) / max_sq;  // pdists.pow(2) then div_(max): data/dataset.py:11-12
}

// out[k] = row[g] for offsets[g] <= k < offsets[g+1]: 4 consecutive k per thread (one 16-byte store), binary search
// for the first, linear walk for the rest (empty groups are skipped).
__global__ void expand_groups_kernel(const int* __restrict__ row, const long long* __restrict__ offsets, int G,
                                     int* __restrict__ out, long long P) {
  long long k = 4 * ((long long)blockIdx.x * blockDim.x + threadIdx.x);
  if (k >= P) return;
  int lo = 0, hi = G;  // last g with offsets[g] <= k
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (__ldg(offsets + mid) <= k) lo = mid; else hi = mid;
  }
  int g = lo;
  int v[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    while (g + 1 < G && k + e >= __ldg(offsets + g + 1)) ++g;
    v[e] = __ldg(row + g);
  }
  if (k + 3 < P && ((reinterpret_cast<size_t>(out + k) & 15) == 0)) {
    *reinterpret_cast<int4*>(out + k) = make_int4(v[0], v[1], v[2], v[3]);
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e) if (k + e < P) out[k + e] = v[e];
  }
}

template <typename L>
__global__ void gather_levels_kernel(const L* __restrict__ levels, int N, const int* __restrict__ slot,
                                     const int* __restrict__ col, long long P, L* __restrict__ out) {
  long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= P) return;
  out[k] = levels[(size_t)slot[k] * N + col[k]];
}

}  // namespace gm

using namespace gm;

#define GM_LEVEL_SWITCH(bytes, STMT)                                      \
  switch (bytes) {                                                        \
    case 1: { typedef unsigned char L; STMT; } break;                     \
    case 2: { typedef unsigned short L; STMT; } break;                    \
    case 4: { typedef int L; STMT; } break;                               \
    default: return GM_EINVAL;                                            \
  }

extern "C" {
#pragma GCC visibility push(default)

size_t gm_bfs_workspace_bytes(int32_t N, int32_t S) { return (N > 0 && S > 0) ? ws_bytes(N, S) : 0; }

int gm_bfs_multi_source(const int32_t* rowptr, const int32_t* colidx, int32_t N, const int32_t* sources, int32_t S,
                        int32_t level_bytes, void* levels, void* workspace, size_t workspace_bytes,
                        gm_stream_t stream) {
  if (N < 0 || S < 0) return GM_EINVAL;
  if (N == 0 || S == 0) return GM_OK;
  if (!rowptr || !colidx || !sources || !levels || !workspace) return GM_ENULL;
  if (workspace_bytes < ws_bytes(N, S)) return GM_EINVAL;
  int rc = GM_OK;
  GM_LEVEL_SWITCH(level_bytes, rc = bfs_run<L>(rowptr, colidx, N, sources, S, (L*)levels, workspace, (cudaStream_t)stream));
  return rc;
}

int gm_levels_to_condensed(int32_t level_bytes, const void* levels, int32_t N, int32_t dtype, void* out,
                           gm_stream_t stream) {
  if (N < 0) return GM_EINVAL;
  long long P = (long long)N * (N - 1) / 2;
  if (P <= 0) return GM_OK;
  if (!levels || !out) return GM_ENULL;
  long long blocks = (P + 255) / 256;
  if (blocks > 0x7fffffffLL) return GM_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == GM_F32) {
    GM_LEVEL_SWITCH(level_bytes, (condensed_kernel<L, float><<<(unsigned)blocks, 256, 0, st>>>((const L*)levels, N, (float*)out, P)));
  } else if (dtype == GM_F64) {
    GM_LEVEL_SWITCH(level_bytes, (condensed_kernel<L, double><<<(unsigned)blocks, 256, 0, st>>>((const L*)levels, N, (double*)out, P)));
  } else {
    return GM_EINVAL;
  }
  note_launch();
  return check_launch();
}

int gm_levels_to_dense_targets(int32_t level_bytes, const void* levels, int32_t N, double max_sq, int32_t dtype,
                               void* dense, gm_stream_t stream) {
  if (N < 0) return GM_EINVAL;
  long long total = (long long)N * N;
  if (total == 0) return GM_OK;
  if (!levels || !dense) return GM_ENULL;
  long long blocks = (total + 255) / 256;
  if (blocks > 0x7fffffffLL) return GM_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == GM_F32) {
    GM_LEVEL_SWITCH(level_bytes, (dense_targets_kernel<L, float><<<(unsigned)blocks, 256, 0, st>>>((const L*)levels, total, (float)max_sq, (float*)dense)));
  } else if (dtype == GM_F64) {
    GM_LEVEL_SWITCH(level_bytes, (dense_targets_kernel<L, double><<<(unsigned)blocks, 256, 0, st>>>((const L*)levels, total, (double)max_sq, (double*)dense)));
  } else {
    return GM_EINVAL;
  }
  note_launch();
  return check_launch();
}

int gm_gather_levels(int32_t level_bytes, const void* levels, int32_t N, const int32_t* src_slot, const int32_t* col,
                     int64_t P, void* out, gm_stream_t stream) {
  if (N < 0 || P < 0) return GM_EINVAL;
  if (P == 0) return GM_OK;
  if (!levels || !src_slot || !col || !out) return GM_ENULL;
  long long blocks = (P + 255) / 256;
  if (blocks > 0x7fffffffLL) return GM_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  GM_LEVEL_SWITCH(level_bytes, (gather_levels_kernel<L><<<(unsigned)blocks, 256, 0, st>>>((const L*)levels, N, src_slot, col, P, (L*)out)));
  note_launch();
  return check_launch();
}

int gm_expand_groups(const int32_t* group_row, const int64_t* offsets, int32_t G, int32_t* out_i, int64_t P,
                     gm_stream_t stream) {
  if (G < 0 || P < 0) return GM_EINVAL;
  if (P == 0) return GM_OK;
  if (G == 0) return GM_EINVAL;
  if (!group_row || !offsets || !out_i) return GM_ENULL;
  long long blocks = ((P + 3) / 4 + 255) / 256;
  if (blocks > 0x7fffffffLL) return GM_EINVAL;
  expand_groups_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(group_row, (const long long*)offsets, G,
                                                                        out_i, P);
  note_launch();
  return check_launch();
}

#pragma GCC visibility pop
}
