// Ranking metrics of an embedding against the BFS layering of the graph: per-layer F1@k moments and the (local) mean
// average precision -- the GPU replacement of the reference's FastPrecision
// (graphembed/pyx/impl/precision.cpp:249-291 AveragePrecision / MeanAveragePrecision, :321-398 LayerF1Scores /
// LayerMeanF1Scores / LayerMeanAverageF1Scores; bound in graphembed/pyx/precision.pyx).
//
// One CTA per shortest-path-tree root u:
//   (1) gather row u of the condensed manifold distances (pdist order) and row u of the BFS level matrix into shared
//       memory as (key, layer) pairs -- key = the IEEE bits of the (positive) distance, so unsigned order == float
//       order; the root itself gets key 0 and therefore sorts first (precision.cpp:281,353 assert exactly that);
//   (2) bitonic sort of the pairs in shared memory (the reference: std::sort per root, SortNodeDists :103-119);
//   (3) rank statistics.  The reference walks the sorted nodes inserting each into a flat_multiset ordered by layer and
//       reads two order statistics per node (:355-385).  Both are prefix counts over the sorted sequence:
//           nodes_before(i)        = 1 + #{j < i : layer_j <= layer_i}
//           same_layer_so_far(i)   =     #{j < i : layer_j == layer_i}
//       so every warp takes a contiguous segment of the sorted sequence; a first pass builds per-segment layer
//       histograms, an exclusive scan over segments (and an inclusive one over layers) gives each warp its starting
//       cumulative counts, and a second pass finishes 32 positions per step with shuffles.
//   (4) per-layer (f1, f1^2, count) are reduced in shared memory and added to the global accumulators.
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/gm_kernels.h"

namespace gm {
void note_launch();
int check_launch();

constexpr int kRankThreads = 1024;
constexpr int kRankWarps = kRankThreads / 32;
constexpr int kMaxLayers = 256;  // uint8 levels: at most 255 hops

template <typename K> struct KeyOf;
template <> struct KeyOf<float> {
  using type = unsigned int;
  static __device__ __forceinline__ type bits(float d) { return __float_as_uint(d); }
  static constexpr type kPad = 0xffffffffu;
};
template <> struct KeyOf<double> {
  using type = unsigned long long;
  static __device__ __forceinline__ type bits(double d) { return (unsigned long long)__double_as_longlong(d); }
  static constexpr type kPad = 0xffffffffffffffffull;
};

// index of pair (a, b), a != b, in the condensed (scipy.squareform / torch.triu_indices) vector -- precision.cpp:232-236
__device__ __forceinline__ long long condensed_index(long long n, long long a, long long b) {
  if (a > b) { long long t = a; a = b; b = t; }
  return n * a - a * (a + 1) / 2 + (b - a) - 1;
}

struct RankOut {
  double* f1_m1;  double* f1_m2;  long long* f1_cnt;    // LayerMeanF1Scores accumulators, [n_layers - 1]
  double* af_m1;  double* af_m2;  long long* af_cnt;    // LayerMeanAverageF1Scores accumulators, [n_layers - 1]
  double* ap_sum;                                       // sum over roots of the average precision
};

template <typename T>
__global__ void __launch_bounds__(kRankThreads)
rank_metrics_kernel(const T* __restrict__ mpdists, const unsigned char* __restrict__ levels, int n, int M, int root_lo,
                    int min_degree, int max_degree, int n_layers, RankOut out) {
  using K = typename KeyOf<T>::type;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  K* keys = reinterpret_cast<K*>(smem_raw);
  unsigned char* lay = smem_raw + (size_t)M * sizeof(K);
  __shared__ int hist[kRankWarps][kMaxLayers];  // per-segment layer counts -> starting cumulative counts
  __shared__ int gcum[kMaxLayers];              // #{v : layer(v) <= l} over the whole tree (root included)
  __shared__ double s_m1[kMaxLayers], s_m2[kMaxLayers];
  __shared__ int s_cnt[kMaxLayers];
  __shared__ double s_ap;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int u = root_lo + blockIdx.x;

  // ---- (1) gather -------------------------------------------------------------------------------------------------
  for (int v = tid; v < M; v += kRankThreads) {
    K key = KeyOf<T>::kPad;
    unsigned char l = 255;
    if (v < n) {
      l = levels[(size_t)u * n + v];
      key = (v == u) ? (K)0 : KeyOf<T>::bits(mpdists[condensed_index(n, u, v)]);
    }
    keys[v] = key;
    lay[v] = l;
  }
  for (int i = tid; i < kRankWarps * kMaxLayers; i += kRankThreads) (&hist[0][0])[i] = 0;
  if (tid < kMaxLayers) { s_m1[tid] = 0.0; s_m2[tid] = 0.0; s_cnt[tid] = 0; }
  if (tid == 0) s_ap = 0.0;
  __syncthreads();

  // ---- (2) bitonic sort of (key, layer), ascending ------------------------------------------------------------------
  for (int k = 2; k <= M; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < (M >> 1); t += kRankThreads) {
        int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));  // index with bit j clear
        int p = i | j;
        bool up = (i & k) == 0;
        K a = keys[i], b = keys[p];
        if ((a > b) == up) {
          keys[i] = b; keys[p] = a;
          unsigned char la = lay[i]; lay[i] = lay[p]; lay[p] = la;
        }
      }
      __syncthreads();
    }
  }

  // ---- (3) rank statistics --------------------------------------------------------------------------------------------
  // positions 1 .. n-1 of the sorted sequence (position 0 is the root); segment of warp w: [1 + w*S, 1 + (w+1)*S)
  const int S = (((n - 1) + kRankWarps - 1) / kRankWarps + 31) & ~31;
  const int seg_lo = 1 + warp * S;
  const int seg_hi = min(n, seg_lo + S);
  for (int pos = seg_lo + lane; pos < seg_hi; pos += 32) atomicAdd(&hist[warp][lay[pos]], 1);
  __syncthreads();
  if (tid < kMaxLayers) {  // exclusive scan over segments, per layer; total per layer
    int run = 0;
    for (int w = 0; w < kRankWarps; ++w) { int t = hist[w][tid]; hist[w][tid] = run; run += t; }
    gcum[tid] = run + (tid == 0 ? 1 : 0);  // the root sits on layer 0
  }
  __syncthreads();
  if (lane == 0) {  // inclusive scan over layers of this warp's starting counts
    int run = 0;
    for (int l = 0; l < kMaxLayers; ++l) { run += hist[warp][l]; hist[warp][l] = run; }
  }
  if (tid == 0) {
    int run = 0;
    for (int l = 0; l < kMaxLayers; ++l) { run += gcum[l]; gcum[l] = run; }
  }
  __syncthreads();
  const int degree = gcum[1] - gcum[0];  // unweighted graph: the neighbours are exactly layer 1
  int* cum = hist[warp];                 // cum[l] = #{sorted positions before the current chunk (root excluded) with layer <= l}
  double ap_local = 0.0;
  for (int base = seg_lo; base < seg_hi; base += 32) {
    const int pos = base + lane;
    const bool valid = pos < seg_hi;
    const int L = valid ? (int)lay[pos] : 0x7fffffff;
    int le = 0, eq = 0;
    #pragma unroll 8
    for (int j = 0; j < 32; ++j) {
      int Lj = __shfl_sync(0xffffffffu, L, j);
      if (j < lane) { le += (Lj <= L); eq += (Lj == L); }
    }
    const bool scored = valid && L >= 1 && L < n_layers;
    double f1 = 0.0;
    if (scored) {
      le += cum[L];
      eq += cum[L] - cum[L - 1];
      const double nodes_before = (double)(le + 1);                 // precision.cpp:361 (+1: self)
      const double precision = nodes_before / (double)pos;          // :365, i == pos
      const double actual_before = (double)((gcum[L - 1] - 1) + eq + 1);  // :367-375
      const double recall = nodes_before / actual_before;           // :379
      f1 = 2.0 * precision * recall / (precision + recall);
      if (L == 1) ap_local += (double)(eq + 1) / (double)pos;       // :282-287: n_correct / i at every neighbour
    }
    // Per-layer sums.  fp64 shared-memory atomics are compare-and-swap loops, and consecutive sorted positions sit on
    // one or two layers: a lane-per-atomic version spent 62 % of its stall samples spinning on ~8 addresses
    // (profiles/r01_v11_aux_kernels_full.txt).  So: one reduction per DISTINCT layer of the chunk, one atomic each.
    unsigned todo = __ballot_sync(0xffffffffu, scored);
    while (todo) {
      const int leader = __ffs(todo) - 1;
      const int Lc = __shfl_sync(0xffffffffu, L, leader);
      const bool mine = scored && L == Lc;
      const unsigned group = __ballot_sync(0xffffffffu, mine);
      double s1 = mine ? f1 : 0.0, s2 = mine ? f1 * f1 : 0.0;
      #pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
      }
      if (lane == leader) {
        atomicAdd(&s_m1[Lc - 1], s1);
        atomicAdd(&s_m2[Lc - 1], s2);
        atomicAdd(&s_cnt[Lc - 1], __popc(group));
      }
      todo &= ~group;
    }
    __syncwarp();
    // advance the cumulative counts past this chunk: cum[l] += #{j in chunk : L_j <= l}
    for (int l0 = 0; l0 < n_layers; l0 += 32) {  // warp-uniform trip count: every lane takes part in the shuffles
      const int l = l0 + lane;
      int add = 0;
      #pragma unroll 8
      for (int j = 0; j < 32; ++j) add += (__shfl_sync(0xffffffffu, L, j) <= l);
      if (l < n_layers) cum[l] += add;
    }
    __syncwarp();
  }
  #pragma unroll
  for (int o = 16; o > 0; o >>= 1) ap_local += __shfl_xor_sync(0xffffffffu, ap_local, o);
  if (lane == 0 && ap_local != 0.0) atomicAdd(&s_ap, ap_local);
  __syncthreads();

  // ---- (4) flush ------------------------------------------------------------------------------------------------------
  if (tid < n_layers - 1) {
    const int c = s_cnt[tid];
    if (c > 0) {
      if (degree >= min_degree && degree <= max_degree) {  // precision.cpp:404-407
        atomicAdd(out.f1_m1 + tid, s_m1[tid]);
        atomicAdd(out.f1_m2 + tid, s_m2[tid]);
        atomicAdd((unsigned long long*)out.f1_cnt + tid, (unsigned long long)c);
      }
      const double avg = s_m1[tid] / (double)c;  // :431-437: the root's mean F1 on this layer counts once
      atomicAdd(out.af_m1 + tid, avg);
      atomicAdd(out.af_m2 + tid, avg * avg);
      atomicAdd((unsigned long long*)out.af_cnt + tid, 1ull);
    }
  }
  if (tid == 0 && degree > 0) atomicAdd(out.ap_sum, s_ap / (double)degree);  // :289
}

template <typename T>
static int rank_launch(const void* mpdists, const void* levels, int n, int root_lo, int root_hi, int min_degree,
                       int max_degree, int n_layers, const RankOut& out, cudaStream_t stream) {
  using K = typename KeyOf<T>::type;
  int M = 64;
  while (M < n) M <<= 1;
  const size_t smem = (size_t)M * (sizeof(K) + 1);
  if (smem > 160 * 1024) return GM_EUNSUPPORTED;  // static shared arrays take another ~38 KB of the 227 KB
  auto kern = rank_metrics_kernel<T>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  kern<<<root_hi - root_lo, kRankThreads, smem, stream>>>((const T*)mpdists, (const unsigned char*)levels, n, M, root_lo,
                                                          min_degree, max_degree, n_layers, out);
  note_launch();
  return check_launch();
}

}  // namespace gm

using namespace gm;

extern "C" {
#pragma GCC visibility push(default)

int gm_rank_metrics(int32_t dtype, const void* mpdists, const void* levels_u8, int32_t N, int32_t root_lo,
                    int32_t root_hi, int32_t min_degree, int32_t max_degree, int32_t n_layers, double* f1_m1,
                    double* f1_m2, int64_t* f1_cnt, double* af_m1, double* af_m2, int64_t* af_cnt, double* ap_sum,
                    gm_stream_t stream) {
  if (dtype != GM_F32 && dtype != GM_F64) return GM_EINVAL;
  if (N < 2 || root_lo < 0 || root_hi > N || root_lo > root_hi) return GM_EINVAL;
  if (n_layers < 2 || n_layers > kMaxLayers) return GM_EINVAL;
  if (root_lo == root_hi) return GM_OK;
  if (!mpdists || !levels_u8 || !f1_m1 || !f1_m2 || !f1_cnt || !af_m1 || !af_m2 || !af_cnt || !ap_sum) return GM_ENULL;
  RankOut out{f1_m1, f1_m2, (long long*)f1_cnt, af_m1, af_m2, (long long*)af_cnt, ap_sum};
  if (dtype == GM_F32)
    return rank_launch<float>(mpdists, levels_u8, N, root_lo, root_hi, min_degree, max_degree, n_layers, out,
                              (cudaStream_t)stream);
  return rank_launch<double>(mpdists, levels_u8, N, root_lo, root_hi, min_degree, max_degree, n_layers, out,
                             (cudaStream_t)stream);
}

#pragma GCC visibility pop
}  // extern "C"
