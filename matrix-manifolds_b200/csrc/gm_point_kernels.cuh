// Kernels shared by gm_point_spd.cu / gm_point_vec.cu: one point per thread,
// fused optimizer update and the single-op Manifold API calls.
#pragma once
#include "gm_launch.cuh"
#include "gm_pointops.cuh"

namespace gm {

// Peer-memory owner update (multi-GPU): every rank holds a full point table and a full table of its own partial
// gradients in memory that every other rank of the NVLink domain has mapped (cudaIpc).  Rank `rank` owns rows
// [row_lo, row_lo + N): it sums those rows over all ranks' gradient tables (the reduce-scatter), applies the optimizer
// update, and stores the new rows into every rank's point table (the all-gather) -- one kernel, no staging copies.
// flags[r] is rank r's flag block: [0, MAXP) "ready" words, [MAXP, 2 MAXP) "done" words, each written by the rank
// of that index, and [2 MAXP] a block counter private to rank r.
constexpr int kMaxPeers = GM_MAX_PEERS;
struct PeerTable {
  int world, rank;
  long long row_lo;
  unsigned long long epoch;  // strictly increasing over calls, identical on every rank
  void* x[kMaxPeers];
  const void* g[kMaxPeers];
  unsigned long long* flags[kMaxPeers];
  const double* acc[kMaxPeers];  // every rank's [loss, scale-grad, ...] accumulator (may be null)
  double* acc_out;               // local: sum over ranks of acc[r][0..n_acc)
  int n_acc;
  void* gsum;                    // local workspace (N_owned rows): non-null selects the pipelined exchange
  unsigned long long timeout_ns; // flag_wait gives up (traps) after this long; 0 = never
  int chunks;                    // pipelined exchange: number of row chunks (1..8)
};

struct PointArgs {
  int kind, dtype, n, p;
  unsigned flags;
  double wmin, wmax;
  const void* c_dev;  // GM_UNIVERSAL: device scalar c
  int op;  // gm_point_op, or -1 for the optimizer step
  OptimCfg oc;
  int grassmann_retr_qr;
  void* x;          // optimizer: in/out.  point op: const input
  const void* u;    // optimizer: grad
  const void* v;
  void* out;        // optimizer: unused
  void* buf1;
  void* buf2;
  long long N;
  cudaStream_t stream;
  const PeerTable* peer;  // non-null: peer-memory owner update (x / u ignored, taken from the table)
};

template <class Man, typename T>
__global__ void __launch_bounds__(128)
optim_kernel(Man man, OptimCfg oc, T* __restrict__ x, T* __restrict__ grad, T* __restrict__ buf1,
             T* __restrict__ buf2, long long N) {
  constexpr int CAP = Man::CAP;
  long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= N) return;
  const int cnt = man.count();
  T xs[CAP], gs[CAP], b1[CAP], b2[CAP];
  const long long base = k * cnt;
  if constexpr (Man::kStatic) {
    load_row<T, CAP>(x, k, xs);
    load_row<T, CAP>(grad, k, gs);
    if (buf1) load_row<T, CAP>(buf1, k, b1);
    if (buf2) load_row<T, CAP>(buf2, k, b2);
    if (!buf1) { GM_UNROLL for (int e = 0; e < CAP; ++e) b1[e] = (T)0; }
    if (!buf2) { GM_UNROLL for (int e = 0; e < CAP; ++e) b2[e] = (T)0; }
  } else {
    for (int e = 0; e < cnt; ++e) {
      xs[e] = x[base + e];
      gs[e] = grad[base + e];
      b1[e] = buf1 ? buf1[base + e] : (T)0;
      b2[e] = buf2 ? buf2[base + e] : (T)0;
    }
  }
  if (oc.zero_grad) {  // the next step's zero_grad(), folded in: the row was just read, hand it back cleared
    if constexpr (Man::kStatic) {
      T z[CAP];
      GM_UNROLL for (int e = 0; e < CAP; ++e) z[e] = (T)0;
      store_row<T, CAP>(grad, k, z);
    } else {
      for (int e = 0; e < cnt; ++e) grad[base + e] = (T)0;
    }
  }
  optim_update<Man, T>(man, oc, xs, gs, b1, b2);
  if constexpr (Man::kStatic) {
    store_row<T, CAP>(x, k, xs);
    if (buf1) store_row<T, CAP>(buf1, k, b1);
    if (buf2) store_row<T, CAP>(buf2, k, b2);
  } else {
    for (int e = 0; e < cnt; ++e) {
      x[base + e] = xs[e];
      if (buf1) buf1[base + e] = b1[e];
      if (buf2) buf2[base + e] = b2[e];
    }
  }
}

// ---- cross-GPU flag words (system scope) ------------------------------------------------------------------------------
__device__ __forceinline__ void flag_store_release(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long flag_load_acquire(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// Spin until *p >= epoch.  A peer that never arrives (crashed rank) must not hang the GPU for ever: trap after
// timeout_ns (PeerTable::timeout_ns: GM_PEER_TIMEOUT_S seconds, default 300 -- rank skew of minutes is legitimate: a
// rank writing a checkpoint, a slow data loader, a debugger; 0 waits for ever).
__device__ __forceinline__ void flag_wait(const unsigned long long* p, unsigned long long epoch,
                                          unsigned long long timeout_ns) {
  unsigned long long t0 = 0;
  asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t0));
  unsigned spins = 0;
  while (flag_load_acquire(p) < epoch) {
    if ((++spins & 0x3ff) == 0 && timeout_ns) {
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t1));
      if (t1 - t0 > timeout_ns) __trap();
    }
    __nanosleep(64);
  }
}

// Fused reduce-scatter + optimizer update + all-gather over peer memory (see PeerTable).
//   entry : block 0 tells every peer "my partial gradients are final" (ready[rank] = epoch); every block waits for
//           the ready word of every rank before it touches a gradient table
//   body  : a block owns a tile of 128 consecutive rows.  (1) all threads pull the tile from every rank's gradient
//           table with warp-contiguous 16-byte loads (512 B per warp instruction -- NVLink packets stay full; a
//           thread-per-row gather would issue 32-byte requests), sum it in rank order and park it in shared memory;
//           (2) thread k updates row k (x, buf1, buf2) and writes the new row back into the tile; (3) all threads
//           push the tile into every rank's point table with the same warp-contiguous stores.
//   exit  : the last block to finish tells every peer "I have read your gradients and written your points"
//           (done[rank] = epoch) and waits for the same from everyone, so that when this kernel completes the local
//           point table is fully updated and the local gradient table may be zeroed again.
// Tiles that do not fit the shared-memory budget (row > kPeerTileRowBytes) take the per-thread path.
constexpr int kPeerTileRowBytes = 256;
constexpr int kPeerTileBytes = 128 * kPeerTileRowBytes;

template <typename T, typename V>
__device__ __forceinline__ V vec_add(V a, V b) {
  constexpr int n = sizeof(V) / sizeof(T);
  T* pa = reinterpret_cast<T*>(&a);
  const T* pb = reinterpret_cast<const T*>(&b);
  GM_UNROLL for (int j = 0; j < n; ++j) pa[j] += pb[j];
  return a;
}

// tile <- sum over ranks of g[r][off .. off + nvec) in units of V.  A remote load takes microseconds over NVLink, so the
// loads of ALL ranks (U vectors each) are issued before the first sum: kMaxPeers * U independent 16-byte requests in
// flight per thread instead of U (the rank loop is fully unrolled and predicated on r < world); the sum itself stays in
// rank order, so the result does not depend on the number of loads in flight.
template <typename T, typename V>
__device__ __forceinline__ void peer_pull_sum(const PeerTable& pt, size_t byte_off, int nvec, V* tile, int tid) {
  constexpr int U = 2;
  for (int c0 = tid; c0 < nvec; c0 += 128 * U) {
    V v[kMaxPeers][U];
    GM_UNROLL for (int r = 0; r < kMaxPeers; ++r) {
      if (r < pt.world) {
        const V* src = reinterpret_cast<const V*>((const char*)pt.g[r] + byte_off);
        GM_UNROLL for (int u = 0; u < U; ++u) {
          int c = c0 + u * 128;
          v[r][u] = (c < nvec) ? src[c] : V{};
        }
      }
    }
    GM_UNROLL for (int u = 0; u < U; ++u) {
      V acc = v[0][u];
      GM_UNROLL for (int r = 1; r < kMaxPeers; ++r)
        if (r < pt.world) acc = vec_add<T, V>(acc, v[r][u]);
      int c = c0 + u * 128;
      if (c < nvec) tile[c] = acc;
    }
  }
}

template <typename V>
__device__ __forceinline__ void peer_push(const PeerTable& pt, size_t byte_off, int nvec, const V* tile, int tid) {
  for (int c = tid; c < nvec; c += 128) {
    V v = tile[c];
    for (int r = 0; r < pt.world; ++r) reinterpret_cast<V*>((char*)pt.x[r] + byte_off)[c] = v;
  }
}

template <class Man, typename T>
__global__ void __launch_bounds__(128)
peer_optim_kernel(Man man, OptimCfg oc, PeerTable pt, T* __restrict__ buf1, T* __restrict__ buf2, long long N,
                  int use_tile) {
  constexpr int CAP = Man::CAP;
  extern __shared__ __align__(16) char tile_mem[];
  __shared__ int is_last;
  const int tid = threadIdx.x;
  unsigned long long* my_flags = pt.flags[pt.rank];
  if (blockIdx.x == 0 && tid < pt.world) flag_store_release(pt.flags[tid] + pt.rank, pt.epoch);
  if (tid < pt.world) flag_wait(my_flags + tid, pt.epoch, pt.timeout_ns);
  __syncthreads();

  const int cnt = man.count();
  const long long row0 = (long long)blockIdx.x * 128;
  const long long k = row0 + tid;
  const int rows_here = (int)((N - row0 < 128) ? (N - row0) : 128);
  const size_t tile_off = (size_t)(pt.row_lo + row0) * cnt * sizeof(T);  // byte offset of the tile in every table
  const int tile_bytes = rows_here * cnt * (int)sizeof(T);
  const bool vec16 = use_tile && (tile_off % 16 == 0) && (tile_bytes % 16 == 0);
  T* tile = reinterpret_cast<T*>(tile_mem);
  if (use_tile) {
    if (vec16) peer_pull_sum<T, float4>(pt, tile_off, tile_bytes / 16, reinterpret_cast<float4*>(tile_mem), tid);
    else peer_pull_sum<T, T>(pt, tile_off, tile_bytes / (int)sizeof(T), tile, tid);
    __syncthreads();
  }
  if (k < N) {
    const long long row = pt.row_lo + k;
    const long long base = row * cnt, lbase = k * cnt;
    T xs[CAP], gs[CAP], b1[CAP], b2[CAP];
    if constexpr (Man::kStatic) {
      if (use_tile) {
        GM_UNROLL for (int e = 0; e < CAP; ++e) gs[e] = tile[tid * CAP + e];
      } else {
        load_row<T, CAP>((const T*)pt.g[0], row, gs);
        for (int r = 1; r < pt.world; ++r) {
          T gr[CAP];
          load_row<T, CAP>((const T*)pt.g[r], row, gr);
          GM_UNROLL for (int e = 0; e < CAP; ++e) gs[e] += gr[e];
        }
      }
      load_row<T, CAP>((const T*)pt.x[pt.rank], row, xs);
      if (buf1) load_row<T, CAP>(buf1, k, b1);
      if (buf2) load_row<T, CAP>(buf2, k, b2);
      if (!buf1) { GM_UNROLL for (int e = 0; e < CAP; ++e) b1[e] = (T)0; }
      if (!buf2) { GM_UNROLL for (int e = 0; e < CAP; ++e) b2[e] = (T)0; }
    } else {
      const T* xl = (const T*)pt.x[pt.rank];
      for (int e = 0; e < cnt; ++e) {
        T sum;
        if (use_tile) {
          sum = tile[tid * cnt + e];
        } else {
          sum = ((const T*)pt.g[0])[base + e];
          for (int r = 1; r < pt.world; ++r) sum += ((const T*)pt.g[r])[base + e];
        }
        gs[e] = sum;
        xs[e] = xl[base + e];
        b1[e] = buf1 ? buf1[lbase + e] : (T)0;
        b2[e] = buf2 ? buf2[lbase + e] : (T)0;
      }
    }
    optim_update<Man, T>(man, oc, xs, gs, b1, b2);
    if constexpr (Man::kStatic) {
      if (use_tile) {
        GM_UNROLL for (int e = 0; e < CAP; ++e) tile[tid * CAP + e] = xs[e];
      } else {
        for (int r = 0; r < pt.world; ++r) store_row<T, CAP>((T*)pt.x[r], row, xs);
      }
      if (buf1) store_row<T, CAP>(buf1, k, b1);
      if (buf2) store_row<T, CAP>(buf2, k, b2);
    } else {
      for (int e = 0; e < cnt; ++e) {
        if (use_tile) {
          tile[tid * cnt + e] = xs[e];
        } else {
          for (int r = 0; r < pt.world; ++r) ((T*)pt.x[r])[base + e] = xs[e];
        }
        if (buf1) buf1[lbase + e] = b1[e];
        if (buf2) buf2[lbase + e] = b2[e];
      }
    }
  }
  if (use_tile) {
    __syncthreads();
    if (vec16) peer_push<float4>(pt, tile_off, tile_bytes / 16, reinterpret_cast<const float4*>(tile_mem), tid);
    else peer_push<T>(pt, tile_off, tile_bytes / (int)sizeof(T), tile, tid);
  }
  // per-step scalars (loss, scale gradients): every rank's accumulator is final once its ready word is seen
  if (blockIdx.x == 0 && tid >= 32 && tid < 32 + pt.n_acc) {
    double s = 0.0;
    for (int r = 0; r < pt.world; ++r) s += pt.acc[r][tid - 32];
    pt.acc_out[tid - 32] = s;
  }
  __threadfence_system();
  __syncthreads();
  if (tid == 0) {
    unsigned long long prev = atomicAdd(my_flags + 2 * kMaxPeers, 1ull);
    is_last = (prev == (unsigned long long)gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    if (tid == 0) my_flags[2 * kMaxPeers] = 0;
    if (tid < pt.world) {
      flag_store_release(pt.flags[tid] + kMaxPeers + pt.rank, pt.epoch);
      flag_wait(my_flags + kMaxPeers + tid, pt.epoch, pt.timeout_ns);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Pipelined exchange (PeerTable::gsum != nullptr).  The fused kernel above alternates, inside every block, between
// pulling (inbound NVLink), computing and pushing (outbound NVLink), so neither direction of the links is ever busy
// for long.  Here the owned rows are cut into chunks and two kernels run side by side on two streams:
//   peer_pull_kernel        chunk c+1: gsum[rows] = sum over ranks of their partial gradient rows -- nothing but 16-byte
//                           peer loads, all ranks' loads of a vector in flight before the first add, few registers, so
//                           thousands of threads keep the inbound links full
//   peer_update_push_kernel chunk c: optimizer update from gsum (local), new rows stored into every rank's point table
// The cross-GPU handshakes stay where the data dependencies are: the FIRST pull kernel of a step publishes "my partial
// gradients are final" and waits for everyone's; the LAST update kernel publishes "I have read your gradients and
// written your points" and waits for everyone's (and sums the per-step scalars).
// ---------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256, 4)
peer_pull_kernel(PeerTable pt, size_t byte_off, size_t gsum_off, long long nvec, int handshake) {
  const int tid = threadIdx.x;
  if (handshake) {
    unsigned long long* my_flags = pt.flags[pt.rank];
    if (blockIdx.x == 0 && tid < pt.world) flag_store_release(pt.flags[tid] + pt.rank, pt.epoch);
    if (tid < pt.world) flag_wait(my_flags + tid, pt.epoch, pt.timeout_ns);
    __syncthreads();
  }
  float4* out = reinterpret_cast<float4*>((char*)pt.gsum + gsum_off);
  constexpr int U = 2;
  const long long stride = (long long)gridDim.x * 256 * U;
  for (long long c0 = (long long)blockIdx.x * 256 * U + tid; c0 < nvec; c0 += stride) {
    float4 v[kMaxPeers][U];
    GM_UNROLL for (int r = 0; r < kMaxPeers; ++r) {
      if (r < pt.world) {
        const float4* src = reinterpret_cast<const float4*>((const char*)pt.g[r] + byte_off);
        GM_UNROLL for (int u = 0; u < U; ++u) {
          const long long c = c0 + (long long)u * 256;
          v[r][u] = (c < nvec) ? src[c] : float4{};
        }
      }
    }
    GM_UNROLL for (int u = 0; u < U; ++u) {
      float4 acc = v[0][u];
      GM_UNROLL for (int r = 1; r < kMaxPeers; ++r) {
        if (r < pt.world) {
          acc = vec_add<T, float4>(acc, v[r][u]);
        }
      }
      const long long c = c0 + (long long)u * 256;
      if (c < nvec) out[c] = acc;
    }
  }
}

template <class Man, typename T>
__global__ void __launch_bounds__(128)
peer_update_push_kernel(Man man, OptimCfg oc, PeerTable pt, T* __restrict__ buf1, T* __restrict__ buf2,
                        long long row0_chunk, long long rows_chunk, int last) {
  constexpr int CAP = Man::CAP;
  extern __shared__ __align__(16) char tile_mem[];
  __shared__ int is_last;
  const int tid = threadIdx.x;
  const int cnt = man.count();
  const long long row0 = row0_chunk + (long long)blockIdx.x * 128;  // first owned row of this tile
  const long long k = row0 + tid;
  const long long end = row0_chunk + rows_chunk;
  const int rows_here = (int)((end - row0 < 128) ? (end - row0) : 128);
  T* tile = reinterpret_cast<T*>(tile_mem);
  if (k < end) {
    const long long row = pt.row_lo + k;
    T xs[CAP], gs[CAP], b1[CAP], b2[CAP];
    static_assert(Man::kStatic, "the pipelined exchange is built for fixed-size points");
    load_row<T, CAP>((const T*)pt.gsum, k, gs);
    load_row<T, CAP>((const T*)pt.x[pt.rank], row, xs);
    if (buf1) load_row<T, CAP>(buf1, k, b1);
    if (buf2) load_row<T, CAP>(buf2, k, b2);
    if (!buf1) { GM_UNROLL for (int e = 0; e < CAP; ++e) b1[e] = (T)0; }
    if (!buf2) { GM_UNROLL for (int e = 0; e < CAP; ++e) b2[e] = (T)0; }
    optim_update<Man, T>(man, oc, xs, gs, b1, b2);
    GM_UNROLL for (int e = 0; e < CAP; ++e) tile[tid * CAP + e] = xs[e];
    if (buf1) store_row<T, CAP>(buf1, k, b1);
    if (buf2) store_row<T, CAP>(buf2, k, b2);
  }
  __syncthreads();
  const size_t tile_off = (size_t)(pt.row_lo + row0) * cnt * sizeof(T);
  peer_push<float4>(pt, tile_off, rows_here * cnt * (int)sizeof(T) / 16, reinterpret_cast<const float4*>(tile_mem), tid);
  if (!last) return;
  unsigned long long* my_flags = pt.flags[pt.rank];
  if (blockIdx.x == 0 && tid >= 32 && tid < 32 + pt.n_acc) {
    double s = 0.0;
    for (int r = 0; r < pt.world; ++r) s += pt.acc[r][tid - 32];
    pt.acc_out[tid - 32] = s;
  }
  __threadfence_system();
  __syncthreads();
  if (tid == 0) {
    unsigned long long prev = atomicAdd(my_flags + 2 * kMaxPeers, 1ull);
    is_last = (prev == (unsigned long long)gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    if (tid == 0) my_flags[2 * kMaxPeers] = 0;
    if (tid < pt.world) {
      flag_store_release(pt.flags[tid] + kMaxPeers + pt.rank, pt.epoch);
      flag_wait(my_flags + kMaxPeers + tid, pt.epoch, pt.timeout_ns);
    }
  }
}

// side stream + events of the pipelined exchange, one set per device, created on first use
struct PeerPipe {
  cudaStream_t side = nullptr;
  cudaEvent_t pulled[8] = {};
  cudaEvent_t joined = nullptr;
};
PeerPipe& peer_pipe();  // gm_api.cu

template <class Man, typename T>
static int launch_peer_pipelined(const Man& man, const PointArgs& a) {
  const PeerTable& pt = *a.peer;
  const int cnt = man.count();
  const size_t row_bytes = (size_t)cnt * sizeof(T);
  PeerPipe& pp = peer_pipe();
  if (!pp.side) return GM_EINVAL;
  // chunks of whole 128-row tiles, at most 4 (and at least ~8k rows each: below that the launches cost more than the
  // overlap buys)
  const long long tiles = (a.N + 127) / 128;
  int K = (int)(tiles / 64);
  const int kmax = pt.chunks >= 1 && pt.chunks <= 8 ? pt.chunks : 4;
  K = K < 1 ? 1 : (K > kmax ? kmax : K);
  const long long tiles_per = (tiles + K - 1) / K;
  int launched = 0;
  for (int c = 0; c < K; ++c) {
    const long long r0 = (long long)c * tiles_per * 128;
    if (r0 >= a.N) { K = c; break; }
    const long long rows = (a.N - r0 < tiles_per * 128) ? (a.N - r0) : tiles_per * 128;
    const size_t off = (size_t)(pt.row_lo + r0) * row_bytes;
    const long long nvec = (long long)(rows * row_bytes / 16);
    // two blocks per SM keep megabytes of 16-byte peer loads in flight and leave the rest of every SM to the
    // update+push kernel of the previous chunk, which runs beside this one
    long long pb = (nvec + 256 * 2 - 1) / (256 * 2);
    if (pb > 148 * 2) pb = 148 * 2;
    peer_pull_kernel<T><<<(unsigned)pb, 256, 0, a.stream>>>(pt, off, (size_t)r0 * row_bytes, nvec, c == 0 ? 1 : 0);
    note_launch();
    cudaEventRecord(pp.pulled[c], a.stream);
    ++launched;
  }
  for (int c = 0; c < K; ++c) {
    const long long r0 = (long long)c * tiles_per * 128;
    const long long rows = (a.N - r0 < tiles_per * 128) ? (a.N - r0) : tiles_per * 128;
    cudaStreamWaitEvent(pp.side, pp.pulled[c], 0);
    const long long ub = (rows + 127) / 128;
    peer_update_push_kernel<Man, T><<<(unsigned)ub, 128, 128 * row_bytes, pp.side>>>(
        man, a.oc, pt, (T*)a.buf1, (T*)a.buf2, r0, rows, c == K - 1 ? 1 : 0);
    note_launch();
  }
  cudaEventRecord(pp.joined, pp.side);
  cudaStreamWaitEvent(a.stream, pp.joined, 0);
  (void)launched;
  return check_launch();
}

template <class Man, typename T>
__global__ void __launch_bounds__(128)
point_op_kernel(Man man, int op, const T* __restrict__ x, const T* __restrict__ u, const T* __restrict__ v,
                T* __restrict__ out, long long N) {
  constexpr int CAP = Man::CAP;
  long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= N) return;
  const int cnt = man.count();
  T xs[CAP], us[CAP], vs[CAP], os[CAP];
  const long long base = k * cnt;
  if constexpr (Man::kStatic) {
    load_row<T, CAP>(x, k, xs);
    if (u) load_row<T, CAP>(u, k, us);
    if (v) load_row<T, CAP>(v, k, vs);
  } else {
    for (int e = 0; e < cnt; ++e) {
      xs[e] = x[base + e];
      us[e] = u ? u[base + e] : (T)0;
      vs[e] = v ? v[base + e] : (T)0;
    }
  }
  bool scalar = false;
  T sval = (T)0;
  switch (op) {
    case GM_OP_EXP: man.exp(xs, us, os); break;
    case GM_OP_RETR: man.retr(xs, us, os); break;
    case GM_OP_LOG: man.log(xs, us, os); break;
    case GM_OP_PROJU: man.proju(xs, us, os); break;
    case GM_OP_PROJX: man.projx(xs, os); break;
    case GM_OP_EGRAD2RGRAD: man.egrad2rgrad(xs, us, os); break;
    case GM_OP_INNER: scalar = true; sval = man.inner(xs, us, vs); break;
    case GM_OP_NORM2: scalar = true; sval = man.norm2(xs, us); break;
    case GM_OP_TRANSP: man.transp(xs, us, vs, os); break;
    case GM_OP_SPD_SQRTM:
      if constexpr (Man::kStatic) { man.sqrtm(xs, os); break; } else { return; }
    default: return;
  }
  if (scalar) {
    out[k] = sval;
  } else {
    if constexpr (Man::kStatic) {
      store_row<T, CAP>(out, k, os);
    } else {
      for (int e = 0; e < cnt; ++e) out[base + e] = os[e];
    }
  }
}

template <class Man, typename T>
static int launch_point(const Man& man, const PointArgs& a) {
  if (a.N <= 0) return 0;
  const int threads = 128;
  long long blocks = (a.N + threads - 1) / threads;
  if (blocks > 0x7fffffffLL) return GM_EINVAL;
  if (a.op < 0 && a.peer) {
    const size_t tile_bytes = (size_t)threads * man.count() * sizeof(T);
    const int use_tile = tile_bytes <= (size_t)kPeerTileBytes;
    if constexpr (Man::kStatic) {
      if (a.peer->gsum && use_tile && (man.count() * sizeof(T)) % 16 == 0) return launch_peer_pipelined<Man, T>(man, a);
    }
    peer_optim_kernel<Man, T><<<(unsigned)blocks, threads, use_tile ? tile_bytes : 0, a.stream>>>(
        man, a.oc, *a.peer, (T*)a.buf1, (T*)a.buf2, a.N, use_tile);
  } else if (a.op < 0)
    optim_kernel<Man, T><<<(unsigned)blocks, threads, 0, a.stream>>>(man, a.oc, (T*)a.x, (T*)const_cast<void*>(a.u),
                                                                     (T*)a.buf1, (T*)a.buf2, a.N);
  else
    point_op_kernel<Man, T><<<(unsigned)blocks, threads, 0, a.stream>>>(man, a.op, (const T*)a.x, (const T*)a.u,
                                                                        (const T*)a.v, (T*)a.out, a.N);
  note_launch();
  return check_launch();
}

}  // namespace gm
