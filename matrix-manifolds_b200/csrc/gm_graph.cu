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
// Thread t = v*W + w ORs word w of every in-neighbour's frontier and masks the
// visited bits.
//
// v2 (uint8 levels, the common case): what made v1 slow was not the search but
// recording it -- one scattered 1-byte store levels[s*N + v] per newly reached
// (source, node), i.e. S*N partial-sector writes with stride N.  v2
//   * records levels in a NODE-major byte matrix lt[v*Sp + s] during the search:
//     the 64 sources of thread (v, w) are 64 contiguous bytes, updated with
//     16-byte read-blend-write, and transposes it to the (S, N) result once at
//     the end with a tiled kernel (coalesced both ways);
//   * skips a (node, word) whose 64 sources have all arrived (saturated), stops
//     pulling as soon as the remaining unvisited bits are covered, and never
//     reads the frontier words of a neighbour that is in no frontier at all
//     (one byte per node, rebuilt every level).
// Levels wider than one byte (graphs deeper than 254 hops) keep the v1 kernels.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
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

// ---------------------------------------------------------------------------------------------------------------------
// v2: uint8 levels
// ---------------------------------------------------------------------------------------------------------------------
// The frontier of level l IS the record of level l: plane[l][v*W + w] has bit s set iff source w*64+s reaches node v in
// exactly l hops.  The planes of the last kPlanes levels are kept (a ring), so the search itself writes 8 bytes per
// (node, word, level) and no level bytes at all.  Levels are materialised afterwards:
//   depth <= kPlanes (every graph of the BASELINE configs but the two trees): ONE pass reads the planes and writes the
//     (S, N) byte matrix through a shared-memory tile -- 16-byte accesses on both sides, every byte written once;
//   deeper graphs: every kPlanes levels the ring is committed into a node-major byte matrix lt[v*Sp + s] (64 contiguous
//     bytes per thread, read-blend-write), which is transposed at the end.
constexpr int kPlanes = 12;
constexpr int kFlagStride = 32;  // ints between the per-level "found something" flags: one flag per 128-byte line
constexpr int kHubDegree = 192;  // nodes with more in-neighbours than this get a whole block per level

struct Bfs2Ws {
  u64* plane[kPlanes + 1];  // plane[0]: the seeds; level l lives in plane[(l - 1) % kPlanes + 1]
  u64* visited;
  unsigned char* any_cur;   // any_cur[u] != 0 iff node u is in the frontier of at least one source
  unsigned char* any_next;
  unsigned char* lt;        // node-major levels, N x (W*64) bytes (deep graphs only)
  int* hubs;                // [0]: count, [1..]: ids of the nodes with degree > kHubDegree
  int* ell;                 // N x 8: the first 8 in-neighbours of every node (-1 padded), one 32-byte sector per node
  unsigned char* cls;       // 0: degree <= 8 (the ELL row is the whole list), 1: longer, 2: hub
  int* changed;
};
static size_t align256(size_t n) { return (n + 255) / 256 * 256; }
static size_t ws2_bytes(int N, int S) {
  size_t W = ((size_t)S + 63) / 64;
  return (kPlanes + 2) * align256(W * (size_t)N * sizeof(u64)) + 2 * align256((size_t)N) +
         align256((size_t)N * W * 64) + align256(((size_t)N + 1) * sizeof(int)) + align256((size_t)N * 32) +
         align256((size_t)N) + (size_t)(kMaxLevels + 1) * sizeof(int) + 256;
}
static Bfs2Ws carve2(void* ws, int N, int S) {
  size_t W = ((size_t)S + 63) / 64;
  Bfs2Ws b;
  char* p = (char*)ws;
  const size_t words = align256(W * (size_t)N * sizeof(u64));
  for (int i = 0; i <= kPlanes; ++i) { b.plane[i] = (u64*)p; p += words; }
  b.visited = (u64*)p; p += words;
  b.any_cur = (unsigned char*)p; p += align256((size_t)N);
  b.any_next = (unsigned char*)p; p += align256((size_t)N);
  b.lt = (unsigned char*)p; p += align256((size_t)N * W * 64);
  b.hubs = (int*)p; p += align256(((size_t)N + 1) * sizeof(int));
  b.ell = (int*)p; p += align256((size_t)N * 32);
  b.cls = (unsigned char*)p; p += align256((size_t)N);
  b.changed = (int*)p;
  return b;
}
static inline int plane_of(int level) { return level == 0 ? 0 : (level - 1) % kPlanes + 1; }

// sources: seed bit in plane 0 and in visited, "in a frontier" byte; padding bits of the last word count as visited
__global__ void bfs2_seed_kernel(const int* __restrict__ sources, int S, int N, int W, u64* __restrict__ frontier,
                                 u64* __restrict__ visited, unsigned char* __restrict__ any_cur) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < S) {
    int v = sources[i];
    u64 bit = 1ull << (i & 63);
    atomicOr(&frontier[(size_t)v * W + (i >> 6)], bit);
    atomicOr(&visited[(size_t)v * W + (i >> 6)], bit);
    any_cur[v] = 1;
  }
  if ((S & 63) != 0 && i < N) atomicOr(&visited[(size_t)i * W + (W - 1)], ~0ull << (S & 63));
}

// Thread (v, w).  The kernel is latency bound -- a handful of dependent round trips per thread with the memory system
// nearly idle -- so everything that does not depend on another load is fetched in the FIRST round trip: the "search
// still running" flag, the visited word, the node's class and its ELL row (first 8 neighbours, one sector).  Second
// round trip: the 8 frontier words.  Only nodes with more than 8 neighbours go on to walk the CSR list (8 per trip,
// filtered by the "in any frontier" byte); hubs are left to bfs2_hub_kernel.
__global__ void __launch_bounds__(256)
bfs2_level_kernel(const int* __restrict__ rowptr, const int* __restrict__ colidx, int N, int W, int level,
                  const u64* __restrict__ frontier, u64* __restrict__ next, u64* __restrict__ visited,
                  const unsigned char* __restrict__ any_cur, unsigned char* __restrict__ any_next,
                  int* __restrict__ changed, const int* __restrict__ ell, const unsigned char* __restrict__ cls,
                  int hubs_apart) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool in_range = t < (size_t)N * W;
  const size_t tc = in_range ? t : 0;
  const int v = (int)(tc / W);
  const int w = (int)(tc - (size_t)v * W);
  const int alive = changed[(level - 1) * kFlagStride];
  const u64 vis = visited[tc];
  const int kind = cls[v];
  const int4 n0 = reinterpret_cast<const int4*>(ell)[2 * (size_t)v];
  const int4 n1 = reinterpret_cast<const int4*>(ell)[2 * (size_t)v + 1];
  if (!alive) return;  // the search finished at an earlier level (block-uniform)
  const bool mine = in_range && !(hubs_apart && kind == 2);
  u64 fresh = 0;
  if (mine && vis != ~0ull) {  // saturated words have nothing left to learn
    u64 acc = 0;
    const int u8[8] = {n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, n1.z, n1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (u8[i] >= 0) acc |= frontier[(size_t)u8[i] * W + w];
    if (kind != 0 && (acc | vis) != ~0ull) {
      const int e1 = rowptr[v + 1];
      for (int e = rowptr[v] + 8; e < e1; e += 8) {
        int u[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) u[i] = (e + i < e1) ? colidx[e + i] : -1;
        unsigned char in[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) in[i] = (u[i] >= 0) ? any_cur[u[i]] : (unsigned char)0;
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (in[i]) acc |= frontier[(size_t)u[i] * W + w];
        if ((acc | vis) == ~0ull) break;  // every missing source has arrived
      }
    }
    fresh = acc & ~vis;
  }
  if (mine) {
    next[t] = fresh;
    if (fresh) {
      visited[t] = vis | fresh;
      any_next[v] = 1;
    }
  }
  // ONE store per block: millions of stores to the same word serialise in its L2 slice -- and the loads of the
  // neighbouring flag at the top of this kernel queue behind them (that was most of v1's and v2.0's level time)
  if (__syncthreads_or(fresh != 0) && threadIdx.x == 0) changed[level * kFlagStride] = 1;
}

// ids of the nodes whose neighbour list is long enough to be the tail of every level
// ... and the per-node adjacency digest the level kernel starts from (ELL row + class)
__global__ void bfs2_find_hubs_kernel(const int* __restrict__ rowptr, const int* __restrict__ colidx, int N,
                                      int* __restrict__ hubs, int* __restrict__ ell, unsigned char* __restrict__ cls,
                                      int want_hubs) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= N) return;
  const int e0 = rowptr[v], deg = rowptr[v + 1] - e0;
  int row[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) row[i] = (i < deg) ? colidx[e0 + i] : -1;
  reinterpret_cast<int4*>(ell)[2 * (size_t)v] = make_int4(row[0], row[1], row[2], row[3]);
  reinterpret_cast<int4*>(ell)[2 * (size_t)v + 1] = make_int4(row[4], row[5], row[6], row[7]);
  const bool hub = want_hubs && deg > kHubDegree;
  cls[v] = hub ? 2 : (deg > 8 ? 1 : 0);
  if (hub) hubs[1 + atomicAdd(hubs, 1)] = v;
}

// One block per hub node: the 256 threads split into `parts` edge slices x Wc words (Wc = min(W, 32) rounded up to a
// power of two); words beyond 32 are looped.  Slice results are OR-ed through shared memory.
__global__ void __launch_bounds__(256)
bfs2_hub_kernel(const int* __restrict__ rowptr, const int* __restrict__ colidx, int W, int level,
                const u64* __restrict__ frontier, u64* __restrict__ next, u64* __restrict__ visited,
                const unsigned char* __restrict__ any_cur, unsigned char* __restrict__ any_next,
                int* __restrict__ changed, const int* __restrict__ hubs, int Wc) {
  extern __shared__ u64 red[];  // [parts][W]
  if (changed[(level - 1) * kFlagStride] == 0) return;
  const int v = hubs[1 + blockIdx.x];
  const int parts = 256 / Wc;
  const int part = threadIdx.x / Wc, lane_w = threadIdx.x % Wc;
  const int e0 = rowptr[v], e1 = rowptr[v + 1];
  for (int w = lane_w; w < W; w += Wc) {
    const u64 vis = visited[(size_t)v * W + w];
    u64 acc = 0;
    if (vis != ~0ull) {
      for (int e = e0 + part * 4; e < e1; e += parts * 4) {
        int u[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) u[i] = (e + i < e1) ? colidx[e + i] : -1;
        unsigned char in[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) in[i] = (u[i] >= 0) ? any_cur[u[i]] : (unsigned char)0;
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (in[i]) acc |= frontier[(size_t)u[i] * W + w];
        if ((acc | vis) == ~0ull) break;
      }
    }
    red[part * W + w] = acc;
  }
  __syncthreads();
  bool any_fresh = false;
  for (int w = threadIdx.x; w < W; w += 256) {
    u64 acc = 0;
    for (int p = 0; p < parts; ++p) acc |= red[p * W + w];
    const size_t t = (size_t)v * W + w;
    const u64 vis = visited[t];
    const u64 fresh = acc & ~vis;
    next[t] = fresh;
    if (fresh) { visited[t] = vis | fresh; any_fresh = true; }
  }
  if (any_fresh) any_next[v] = 1;
  if (__syncthreads_or(any_fresh) && threadIdx.x == 0) changed[level * kFlagStride] = 1;
}

// 4 bits -> 4 byte masks (0xFF where the bit is set)
__device__ __forceinline__ unsigned spread4(unsigned bits) {
  return (((bits & 0xFu) * 0x00204081u) & 0x01010101u) * 0xFFu;
}
// bytes of 16 sources (bits 16c .. 16c+15 of `fresh`) <- level where their bit is set
__device__ __forceinline__ void blend16(uint4& o, u64 fresh, int c, unsigned level) {
  const unsigned bits = (unsigned)(fresh >> (16 * c)) & 0xFFFFu;
  if (!bits) return;
  const unsigned lv4 = level * 0x01010101u;
  unsigned m;
  m = spread4(bits);       o.x = (o.x & ~m) | (lv4 & m);
  m = spread4(bits >> 4);  o.y = (o.y & ~m) | (lv4 & m);
  m = spread4(bits >> 8);  o.z = (o.z & ~m) | (lv4 & m);
  m = spread4(bits >> 12); o.w = (o.w & ~m) | (lv4 & m);
}

struct PlaneSet {
  const u64* p[kPlanes + 1];
  int level[kPlanes + 1];
  int count;
};

// shallow graphs: levels[s*N + v] straight from the planes.  A block owns 32 nodes x 16 words (1024 sources): it stages
// those words of every plane in shared memory with coalesced loads (32 rows of 128 contiguous bytes per plane), then
// emits one 32-node x 64-source byte tile per word, transposed through a double-buffered shared-memory tile: every
// source row receives one full 32-byte sector.
constexpr int kFinWords = 16;
constexpr int kFinNodes = 32;
__device__ __forceinline__ void blend8(uint2& o, u64 fresh, int c, unsigned level) {  // sources 8c .. 8c+7
  const unsigned bits = (unsigned)(fresh >> (8 * c)) & 0xFFu;
  if (!bits) return;
  const unsigned lv4 = level * 0x01010101u;
  unsigned m;
  m = spread4(bits);      o.x = (o.x & ~m) | (lv4 & m);
  m = spread4(bits >> 4); o.y = (o.y & ~m) | (lv4 & m);
}
__global__ void __launch_bounds__(256)
bfs2_finalize_kernel(PlaneSet ps, int N, int W, int S, unsigned char* __restrict__ levels) {
  extern __shared__ __align__(16) unsigned char fin_smem[];
  u64* stage = reinterpret_cast<u64*>(fin_smem);  // [count][kFinNodes][kFinWords]
  typedef unsigned char TileRow[64 + 8];
  TileRow* tile = reinterpret_cast<TileRow*>(fin_smem + (size_t)ps.count * kFinNodes * kFinWords * 8);  // [2][32] rows
  const int v0 = blockIdx.x * kFinNodes, w0 = blockIdx.y * kFinWords;
  const int nw = (W - w0 < kFinWords) ? (W - w0) : kFinWords;
  for (int i = 0; i < ps.count; ++i) {
    for (int q = threadIdx.x; q < kFinNodes * kFinWords; q += 256) {
      const int rr = q / kFinWords, ww = q % kFinWords;
      u64 val = 0;
      if (v0 + rr < N && ww < nw) val = ps.p[i][(size_t)(v0 + rr) * W + w0 + ww];
      stage[((size_t)i * kFinNodes + rr) * kFinWords + ww] = val;
    }
  }
  __syncthreads();
  const int r = threadIdx.x >> 3, c = threadIdx.x & 7;    // blend: node r, sources 8c .. 8c+7 of the word
  const int q = threadIdx.x >> 2, c2 = threadIdx.x & 3;   // emit: source q of the word, nodes 8*c2 .. 8*c2+7
  for (int ww = 0; ww < nw; ++ww) {
    TileRow* tl = tile + (ww & 1) * kFinNodes;
    uint2 val = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);  // unreached
    for (int i = 0; i < ps.count; ++i)
      blend8(val, stage[((size_t)i * kFinNodes + r) * kFinWords + ww], c, (unsigned)ps.level[i]);
    *reinterpret_cast<uint2*>(&tl[r][8 * c]) = val;
    __syncthreads();  // (the other buffer is free: its readers passed this barrier one iteration ago)
    const int s = (w0 + ww) * 64 + q;
    if (s < S) {
      unsigned char out[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) out[j] = tl[8 * c2 + j][q];
      unsigned char* dst = levels + (size_t)s * N + v0 + 8 * c2;
      if (v0 + 8 * c2 + 7 < N && (reinterpret_cast<size_t>(dst) & 7) == 0) {
        *reinterpret_cast<uint2*>(dst) = *reinterpret_cast<const uint2*>(out);
      } else {
        for (int j = 0; j < 8; ++j)
          if (v0 + 8 * c2 + j < N) dst[j] = out[j];
      }
    }
  }
}

// deep graphs: fold the planes of the ring into the node-major byte matrix (64 contiguous bytes per thread)
__global__ void __launch_bounds__(256)
bfs2_commit_kernel(PlaneSet ps, size_t words, unsigned char* __restrict__ lt) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= words) return;
  u64 f[kPlanes + 1];
  u64 any = 0;
  for (int i = 0; i < ps.count; ++i) { f[i] = ps.p[i][t]; any |= f[i]; }
  if (!any) return;
  uint4* row = reinterpret_cast<uint4*>(lt + t * 64);
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    if (((any >> (16 * c)) & 0xFFFFu) == 0) continue;
    uint4 o = row[c];
    for (int i = 0; i < ps.count; ++i) blend16(o, f[i], c, (unsigned)ps.level[i]);
    row[c] = o;
  }
}

// levels[s*N + v] = lt[v*Sp + s]: 64 x 64 byte tiles through shared memory, 16-byte accesses on both sides
__global__ void __launch_bounds__(256)
bfs2_transpose_kernel(const unsigned char* __restrict__ lt, int N, int Sp, int S, unsigned char* __restrict__ levels) {
  __shared__ __align__(16) unsigned char tile[64][64 + 16];
  const int v0 = blockIdx.x * 64, s0 = blockIdx.y * 64;
  const int r = threadIdx.x >> 2, c = threadIdx.x & 3;
  {
    uint4 val = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
    if (v0 + r < N) val = *reinterpret_cast<const uint4*>(lt + (size_t)(v0 + r) * Sp + s0 + 16 * c);
    *reinterpret_cast<uint4*>(&tile[r][16 * c]) = val;
  }
  __syncthreads();
  const int s = s0 + r;
  if (s >= S) return;
  unsigned char out[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) out[j] = tile[16 * c + j][r];
  unsigned char* dst = levels + (size_t)s * N + v0 + 16 * c;
  if (v0 + 16 * c + 15 < N && (reinterpret_cast<size_t>(dst) & 15) == 0) {
    *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(out);
  } else {
    for (int j = 0; j < 16; ++j)
      if (v0 + 16 * c + j < N) dst[j] = out[j];
  }
}

static int bfs2_run(const int* rowptr, const int* colidx, int N, const int* sources, int S, unsigned char* levels,
                    void* ws, cudaStream_t st) {
  const int W = (S + 63) / 64;
  Bfs2Ws b = carve2(ws, N, S);
  const size_t words = (size_t)W * N;
  // plane 0, visited and the two "in a frontier" byte arrays start at zero; the other planes are fully overwritten
  cudaMemsetAsync(b.plane[0], 0, words * sizeof(u64), st);
  cudaMemsetAsync(b.visited, 0, words * sizeof(u64), st);
  cudaMemsetAsync(b.any_cur, 0, 2 * align256((size_t)N), st);
  cudaMemsetAsync(b.changed, 0, (size_t)256 * kFlagStride * sizeof(int), st);  // fits: kMaxLevels + 1 ints reserved
  int one = 1;
  cudaMemcpyAsync(b.changed, &one, sizeof(int), cudaMemcpyHostToDevice, st);
  const int seed_n = S > N ? S : N;
  bfs2_seed_kernel<<<(seed_n + 127) / 128, 128, 0, st>>>(sources, S, N, W, b.plane[0], b.visited, b.any_cur);
  note_launch();
  const int chunk = 4;
  size_t blocks = (words + 255) / 256;
  if (blocks > 0x7fffffffULL) return GM_EINVAL;
  // hub nodes (one host read per BFS): without this every level ends with the few threads that walk the longest lists
  int n_hubs = 0, Wc = 1;
  while (Wc < W && Wc < 32) Wc <<= 1;
  const size_t hub_smem = (size_t)(256 / Wc) * W * sizeof(u64);
  const bool want_hubs = hub_smem <= 96 * 1024;
  cudaMemsetAsync(b.hubs, 0, sizeof(int), st);
  bfs2_find_hubs_kernel<<<(N + 255) / 256, 256, 0, st>>>(rowptr, colidx, N, b.hubs, b.ell, b.cls, want_hubs ? 1 : 0);
  note_launch();
  if (want_hubs) {
    cudaMemcpyAsync(&n_hubs, b.hubs, sizeof(int), cudaMemcpyDeviceToHost, st);
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return (int)e;
    if (hub_smem > 48 * 1024)
      cudaFuncSetAttribute(bfs2_hub_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hub_smem);
  }
  cudaStream_t hub_st = nullptr;  // the hub kernel of a level runs beside its level kernel
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  if (n_hubs > 0) {
    cudaStreamCreateWithFlags(&hub_st, cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming);
  }
  auto release = [&]() {
    if (hub_st) { cudaStreamDestroy(hub_st); cudaEventDestroy(ev_fork); cudaEventDestroy(ev_join); hub_st = nullptr; }
  };
  bool deep = false;      // the ring has wrapped: levels go through lt
  int committed = -1;     // deep: levels <= committed are in lt
  auto commit = [&](int lo, int hi) {  // fold levels lo..hi (all still in the ring) into lt
    PlaneSet ps;
    ps.count = 0;
    for (int l = lo; l <= hi; ++l) { ps.p[ps.count] = b.plane[plane_of(l)]; ps.level[ps.count] = l; ++ps.count; }
    bfs2_commit_kernel<<<(unsigned)blocks, 256, 0, st>>>(ps, words, b.lt);
    note_launch();
  };
  int level = 1;
  int last_level = 0;  // deepest level that reached a node (known after the host sync)
  while (true) {
    for (int i = 0; i < chunk; ++i, ++level) {
      if (level > 254) { release(); return GM_EUNSUPPORTED; }  // needs a wider level type (the v1 kernels)
      if (level > kPlanes && (level - 1) % kPlanes == 0) {  // about to overwrite level (level - kPlanes)
        if (!deep) {
          deep = true;
          cudaMemsetAsync(b.lt, 0xFF, (size_t)N * W * 64, st);
          commit(0, level - 1);
        } else {
          commit(committed + 1, level - 1);
        }
        committed = level - 1;
      }
      cudaMemsetAsync(b.any_next, 0, (size_t)N, st);
      if (n_hubs > 0) { cudaEventRecord(ev_fork, st); cudaStreamWaitEvent(hub_st, ev_fork, 0); }
      bfs2_level_kernel<<<(unsigned)blocks, 256, 0, st>>>(rowptr, colidx, N, W, level, b.plane[plane_of(level - 1)],
                                                        b.plane[plane_of(level)], b.visited, b.any_cur, b.any_next,
                                                        b.changed, b.ell, b.cls, n_hubs > 0);
      note_launch();
      if (n_hubs > 0) {
        bfs2_hub_kernel<<<n_hubs, 256, hub_smem, hub_st>>>(rowptr, colidx, W, level, b.plane[plane_of(level - 1)],
                                                           b.plane[plane_of(level)], b.visited, b.any_cur, b.any_next,
                                                           b.changed, b.hubs, Wc);
        note_launch();
        cudaEventRecord(ev_join, hub_st);
        cudaStreamWaitEvent(st, ev_join, 0);
      }
      unsigned char* ta = b.any_cur; b.any_cur = b.any_next; b.any_next = ta;
    }
    int flags[chunk * kFlagStride];
    cudaMemcpyAsync(flags, b.changed + (size_t)(level - chunk) * kFlagStride, sizeof(flags), cudaMemcpyDeviceToHost, st);
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { release(); return (int)e; }
    bool done = false;
    for (int i = 0; i < chunk; ++i) {
      if (flags[i * kFlagStride]) last_level = level - chunk + i;
      else { done = true; break; }
    }
    if (done) break;
  }
  // levels beyond last_level found nothing; their kernels returned immediately and left their planes untouched
  if (!deep) {
    PlaneSet ps;
    ps.count = 0;
    for (int l = 0; l <= last_level; ++l) { ps.p[ps.count] = b.plane[plane_of(l)]; ps.level[ps.count] = l; ++ps.count; }
    dim3 grid((unsigned)((N + kFinNodes - 1) / kFinNodes), (unsigned)((W + kFinWords - 1) / kFinWords));
    const size_t fin_smem = (size_t)ps.count * kFinNodes * kFinWords * 8 + 2 * kFinNodes * (64 + 8);
    cudaFuncSetAttribute(bfs2_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fin_smem);
    bfs2_finalize_kernel<<<grid, 256, fin_smem, st>>>(ps, N, W, S, levels);
    note_launch();
  } else {
    if (last_level > committed) commit(committed + 1, last_level);
    dim3 grid((unsigned)((N + 63) / 64), (unsigned)W);
    bfs2_transpose_kernel<<<grid, 256, 0, st>>>(b.lt, N, W * 64, S, levels);
    note_launch();
  }
  release();
  return check_launch();
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

// 3-byte pair words -> the 4-byte packed form of GM_TGT_HOPS_PACKED: 4 pairs (12 bytes in, 16 bytes out) per thread
__global__ void unpack_pairs3_kernel(const unsigned* __restrict__ src, long long P, int* __restrict__ out) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // group of 4 pairs
  const long long k = 4 * q;
  if (k >= P) return;
  auto expand = [](unsigned w24) -> int { return (int)((w24 & 0x1FFFFFu) | (((w24 >> 21) + 1u) << 24)); };
  if (k + 3 < P) {
    const unsigned a = src[3 * q], b = src[3 * q + 1], c = src[3 * q + 2];
    const int4 v = make_int4(expand(a & 0xFFFFFFu), expand((a >> 24) | ((b & 0xFFFFu) << 8)),
                             expand((b >> 16) | ((c & 0xFFu) << 16)), expand(c >> 8));
    *reinterpret_cast<int4*>(out + k) = v;
  } else {
    const unsigned char* bytes = reinterpret_cast<const unsigned char*>(src);
    for (long long e = k; e < P; ++e)
      out[e] = expand((unsigned)bytes[3 * e] | ((unsigned)bytes[3 * e + 1] << 8) | ((unsigned)bytes[3 * e + 2] << 16));
  }
}

// 2-byte pair words (gm_unpack_pairs2): group g's targets arrive sorted, word k = (j_k - j_{k-1}) | (hops - 1) << 13 with
// j_{-1} := base[g].  One block per group walks it in tiles of 256 x 8 words that start at a multiple of 8 words, so that
// a thread's 8 words are one 16-byte load and its 8 results two 16-byte stores (words outside the group are masked):
// per-thread sum of 8 gaps, block-wide exclusive scan (shuffles + one shared-memory hop), running carry across tiles;
// out[k] = j_k | hops << 24 and, optionally, out_i[k] = group_row[g] (gm_expand_groups folded in).
__global__ void __launch_bounds__(256)
unpack_pairs2_kernel(const unsigned short* __restrict__ words, const int* __restrict__ base,
                     const long long* __restrict__ offsets, int* __restrict__ out, const int* __restrict__ group_row,
                     int* __restrict__ out_i, int vec_ok) {
  __shared__ unsigned warp_tot[8];
  __shared__ unsigned tile_tot;
  const int g = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const long long lo = offsets[g], hi = offsets[g + 1], P = offsets[gridDim.x];
  unsigned carry = (unsigned)base[g];
  const int row_i = out_i ? group_row[g] : 0;
  for (long long t0 = lo & ~7LL; t0 < hi; t0 += 256 * 8) {
    const long long k0 = t0 + (long long)tid * 8;
    unsigned w[8];
    const bool inside = k0 >= lo && k0 + 8 <= hi;  // all 8 words belong to this group
    if (vec_ok && k0 < hi && k0 + 8 > lo && k0 + 8 <= P) {
      const uint4 v = *reinterpret_cast<const uint4*>(words + k0);
      w[0] = v.x & 0xFFFFu; w[1] = v.x >> 16; w[2] = v.y & 0xFFFFu; w[3] = v.y >> 16;
      w[4] = v.z & 0xFFFFu; w[5] = v.z >> 16; w[6] = v.w & 0xFFFFu; w[7] = v.w >> 16;
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) w[e] = (k0 + e >= lo && k0 + e < hi) ? (unsigned)words[k0 + e] : 0u;
    }
    unsigned sum = 0;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      if (!inside && !(k0 + e >= lo && k0 + e < hi)) w[e] = 0u;  // a neighbouring group's word
      sum += w[e] & 0x1FFFu;
    }
    unsigned incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned o = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += o;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    unsigned before = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      if (q < wid) before += warp_tot[q];
    }
    if (tid == 255) tile_tot = before + incl;
    unsigned j = carry + before + incl - sum;  // row id reached before this thread's first word
    int r[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      j += w[e] & 0x1FFFu;
      r[e] = (int)(j | (((w[e] >> 13) + 1u) << 24));
    }
    if (inside && vec_ok) {
      int4* o4 = reinterpret_cast<int4*>(out + k0);
      o4[0] = make_int4(r[0], r[1], r[2], r[3]);
      o4[1] = make_int4(r[4], r[5], r[6], r[7]);
      if (out_i) {
        int4* i4 = reinterpret_cast<int4*>(out_i + k0);
        i4[0] = make_int4(row_i, row_i, row_i, row_i);
        i4[1] = make_int4(row_i, row_i, row_i, row_i);
      }
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        if (k0 + e >= lo && k0 + e < hi) {
          out[k0 + e] = r[e];
          if (out_i) out_i[k0 + e] = row_i;
        }
      }
    }
    __syncthreads();
    carry += tile_tot;
    __syncthreads();
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

// GM_BFS_V1=1 keeps the round-1 kernels for uint8 levels too (A/B measurements)
static bool bfs_force_v1() {
  static const bool v = [] { const char* e = getenv("GM_BFS_V1"); return e && e[0] == '1'; }();
  return v;
}

#define GM_LEVEL_SWITCH(bytes, STMT)                                      \
  switch (bytes) {                                                        \
    case 1: { typedef unsigned char L; STMT; } break;                     \
    case 2: { typedef unsigned short L; STMT; } break;                    \
    case 4: { typedef int L; STMT; } break;                               \
    default: return GM_EINVAL;                                            \
  }

extern "C" {
#pragma GCC visibility push(default)

size_t gm_bfs_workspace_bytes(int32_t N, int32_t S) {
  if (N <= 0 || S <= 0) return 0;
  const size_t a = ws_bytes(N, S), b = ws2_bytes(N, S);
  return a > b ? a : b;
}

int gm_bfs_multi_source(const int32_t* rowptr, const int32_t* colidx, int32_t N, const int32_t* sources, int32_t S,
                        int32_t level_bytes, void* levels, void* workspace, size_t workspace_bytes,
                        gm_stream_t stream) {
  if (N < 0 || S < 0) return GM_EINVAL;
  if (N == 0 || S == 0) return GM_OK;
  if (!rowptr || !colidx || !sources || !levels || !workspace) return GM_ENULL;
  if (workspace_bytes < gm_bfs_workspace_bytes(N, S)) return GM_EINVAL;
  int rc = GM_OK;
  if (level_bytes == 1 && !bfs_force_v1())
    return bfs2_run(rowptr, colidx, N, sources, S, (unsigned char*)levels, workspace, (cudaStream_t)stream);
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

int gm_unpack_pairs3(const void* src3, int64_t P, int32_t* out, gm_stream_t stream) {
  if (P < 0) return GM_EINVAL;
  if (P == 0) return GM_OK;
  if (!src3 || !out) return GM_ENULL;
  if ((reinterpret_cast<size_t>(src3) & 3) || (reinterpret_cast<size_t>(out) & 15)) return GM_EINVAL;
  long long blocks = ((P + 3) / 4 + 255) / 256;
  if (blocks > 0x7fffffffLL) return GM_EINVAL;
  unpack_pairs3_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const unsigned*)src3, P, out);
  note_launch();
  return check_launch();
}

int gm_unpack_pairs2(const void* words, const int32_t* base, const int64_t* offsets, int32_t G, int32_t* out,
                     const int32_t* group_row, int32_t* out_i, gm_stream_t stream) {
  if (G < 0) return GM_EINVAL;
  if (G == 0) return GM_OK;
  if (!words || !base || !offsets || !out) return GM_ENULL;
  if ((group_row == nullptr) != (out_i == nullptr)) return GM_ENULL;
  if (reinterpret_cast<size_t>(words) & 1) return GM_EINVAL;
  // 16-byte accesses when every table starts on a 16-byte boundary (device allocations do); scalar otherwise
  const int vec_ok = ((reinterpret_cast<size_t>(words) | reinterpret_cast<size_t>(out) |
                       reinterpret_cast<size_t>(out_i)) & 15) == 0;
  unpack_pairs2_kernel<<<(unsigned)G, 256, 0, (cudaStream_t)stream>>>((const unsigned short*)words, base,
                                                                    (const long long*)offsets, out, group_row, out_i,
                                                                    vec_ok);
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
