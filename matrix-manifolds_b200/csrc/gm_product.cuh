// Fused pair kernel of a PRODUCT manifold: one launch evaluates every factor's squared distance of a pair, the loss
// of the product distance  m = sum_f softplus(s_f) d2_f  (modules.py:84-88: Python's sum() over the factor list, left
// to right), and scatters every factor's gradient -- instead of the F forward launches, the loss launch and the F
// backward launches of the unfused path (gm_pairs_dist2 x F -> gm_product_loss -> gm_pairs_grad x F).
//
// A product may hold at most one SPD factor (the "lead": its size and divergence are compile-time properties of the
// translation unit that instantiates the kernel) and up to kMaxVecExtra vector factors (Lorentz / Sphere / Euclidean,
// run-time kind and length).  Other products keep the unfused path.
//
// acc[0] += sum loss, acc[1 + f] += sum dL/dm * d2_f (the scale gradient of factor f before the sigmoid(s_f) factor),
// the layout gm_product_loss writes.
#pragma once
#include "gm_launch.cuh"

namespace gm {

template <typename T>
struct VecExtraT {
  int kind, n, slot;
  const T* x;
  T* g;
  T sp;
};
template <typename T>
struct ProductExtraT {
  int F, lead_slot, nvec;
  VecExtraT<T> v[kMaxVecExtra];
};
template <typename T>
static ProductExtraT<T> typed_extras(const ProductExtra& p) {
  ProductExtraT<T> t{};
  t.F = p.F; t.lead_slot = p.lead_slot; t.nvec = p.nvec;
  for (int i = 0; i < p.nvec; ++i) {
    t.v[i].kind = p.v[i].kind; t.v[i].n = p.v[i].n; t.v[i].slot = p.v[i].slot;
    t.v[i].x = (const T*)p.v[i].x; t.v[i].g = (T*)p.v[i].g; t.v[i].sp = (T)p.v[i].sp;
  }
  return t;
}

// accumulate one scalar of row `row` (all 32 lanes call; uniform => one atomic)
template <typename T>
__device__ __forceinline__ void warp_accumulate_elem(T* p, T v, bool uniform, bool active) {
  if (uniform) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) atomicAdd(p, v);
  } else if (active) {
    atomicAdd(p, v);
  }
}

// Gradient scatter of one vector-manifold pair: ga[ra] += w d(d2)/dx, gb[rb] += w d(d2)/dy.  All 32 lanes call.
// First-endpoint rows shared by the whole warp are summed with shuffles and added once.  fp32 rows of 16-byte
// multiples go out as 128-bit reductions (REDG.E.ADD.F32x4, as the SPD kernels issue).
template <typename T, int KIND>
__device__ __forceinline__ void vec_scatter(const VecMan<T, KIND>& op, int n, const T* __restrict__ px,
                                            const T* __restrict__ py, const VecCoef<T>& c, T w, T* __restrict__ ga,
                                            T* __restrict__ gb, long long ra, long long rb, bool active) {
  const unsigned full = 0xffffffffu;
  long long ra0 = __shfl_sync(full, ra, 0);
  bool uni_a = __all_sync(full, ra == ra0) && ra0 >= 0;
  bool vectorised = false;
  if constexpr (sizeof(T) == 4) {
    if ((n & 3) == 0 && (((size_t)ga | (size_t)gb) & 15) == 0) {
      vectorised = true;
      for (int e = 0; e < n; e += 4) {
        float gx4[4], gy4[4];
        GM_UNROLL for (int j = 0; j < 4; ++j) {
          float gxe = 0.f, gye = 0.f;
          if (active) {
            op.grad_elem(e + j, px[e + j], py[e + j], c, gxe, gye);
            gxe *= w; gye *= w;
          }
          gx4[j] = gxe; gy4[j] = gye;
        }
        if (uni_a) {
          GM_UNROLL for (int j = 0; j < 4; ++j) gx4[j] = warp_sum(gx4[j]);
          if ((threadIdx.x & 31) == 0)
            atomicAdd(reinterpret_cast<float4*>(ga + ra0 * n + e), make_float4(gx4[0], gx4[1], gx4[2], gx4[3]));
        } else if (active) {
          atomicAdd(reinterpret_cast<float4*>(ga + ra * n + e), make_float4(gx4[0], gx4[1], gx4[2], gx4[3]));
        }
        if (active)
          atomicAdd(reinterpret_cast<float4*>(gb + rb * n + e), make_float4(gy4[0], gy4[1], gy4[2], gy4[3]));
      }
    }
  }
  for (int e = 0; e < n && !vectorised; ++e) {
    T gxe = (T)0, gye = (T)0;
    if (active) {
      op.grad_elem(e, px[e], py[e], c, gxe, gye);
      gxe *= w; gye *= w;
    }
    warp_accumulate_elem<T>(ga + (uni_a ? ra0 : ra) * n + e, gxe, uni_a, active);
    if (active) atomicAdd(gb + rb * n + e, gye);
  }
}

// run-time kind -> VecMan<T, KIND>
template <typename T, class Fn>
__device__ __forceinline__ void with_vec_kind(int kind, Fn&& fn) {
  const T eps = (T)1e-8;
  const T one_m = (T)(1.0 - 1e-8 * 1e-8);
  if (kind == GM_LORENTZ) fn(VecMan<T, VEC_LORENTZ>{eps, one_m, nullptr});
  else if (kind == GM_SPHERE) fn(VecMan<T, VEC_SPHERE>{eps, one_m, nullptr});
  else fn(VecMan<T, VEC_EUCLIDEAN>{eps, one_m, nullptr});
}

// Lead policy without an SPD factor (products of vector manifolds only)
template <typename T>
struct NoLead {
  static constexpr bool kHas = false;
  static constexpr int E = 1;
  __device__ __forceinline__ T dist2_grad(const T (&)[1], const T (&)[1], T (&)[1], T (&)[1]) const { return (T)0; }
};
template <class Op>
struct SpdLead : Op {
  static constexpr bool kHas = true;
};

template <class Lead, typename T>
__global__ void __launch_bounds__(128)
product_pair_kernel(Lead lead, PairSpec ps, const T* __restrict__ xl, T* __restrict__ gl, T sp_lead,
                    ProductExtraT<T> px, TargetSpec tg, LossCfg lc, double* __restrict__ acc) {
  constexpr int E = Lead::E;
  __shared__ double red[2 + kMaxVecExtra][4];
  long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = k < ps.P;
  long long ra = -1, rb = -1;
  T gx[E], gy[E];
  T d2l = (T)0, d2v[kMaxVecExtra];
  VecCoef<T> co[kMaxVecExtra];
  GM_UNROLL for (int v = 0; v < kMaxVecExtra; ++v) { d2v[v] = (T)0; co[v] = VecCoef<T>{}; }
  double loss_v = 0.0;
  T dm = (T)0;
  if (active) {
    decode_pair(ps, k, ra, rb);
    if constexpr (Lead::kHas) {
      T x[E], y[E];
      load_row<T, E>(xl, ra, x);
      load_row<T, E>(xl, rb, y);
      d2l = lead.dist2_grad(x, y, gx, gy);
    }
    GM_UNROLL for (int v = 0; v < kMaxVecExtra; ++v) {
      if (v < px.nvec) {
        const VecExtraT<T>& f = px.v[v];
        with_vec_kind<T>(f.kind, [&](auto op) { d2v[v] = op.value(f.x + ra * f.n, f.x + rb * f.n, f.n, co[v]); });
      }
    }
    // m = sum over the factor list, left to right
    T m = (T)0;
    GM_UNROLL for (int s = 0; s < 1 + kMaxVecExtra; ++s) {
      if (s < px.F) {
        T term = (T)0;
        if (Lead::kHas && s == px.lead_slot) term = sp_lead * d2l;
        GM_UNROLL for (int v = 0; v < kMaxVecExtra; ++v)
          if (v < px.nvec && px.v[v].slot == s) term = px.v[v].sp * d2v[v];
        m = (s == 0) ? term : m + term;
      }
    }
    T g = fetch_target_ps<T>(ps, tg, k, ra, rb);
    loss_v = (double)loss_term<T>(lc, g, m, dm);
  }
  // ---- gradients ----------------------------------------------------------------------------------------------
  if constexpr (Lead::kHas) {
    const T w = dm * sp_lead;
    GM_UNROLL for (int e = 0; e < E; ++e) { gx[e] *= w; gy[e] *= w; }
    warp_accumulate_row<T, E>(gl, ra, gx);
    warp_accumulate_row<T, E>(gl, rb, gy);
  }
  GM_UNROLL for (int v = 0; v < kMaxVecExtra; ++v) {
    if (v < px.nvec) {  // block-uniform
      const VecExtraT<T>& f = px.v[v];
      const T* pxr = f.x + (active ? ra : 0) * f.n;
      const T* pyr = f.x + (active ? rb : 0) * f.n;
      with_vec_kind<T>(f.kind, [&](auto op) {
        vec_scatter<T>(op, f.n, pxr, pyr, co[v], dm * f.sp, f.g, f.g, ra, rb, active);
      });
    }
  }
  // ---- loss and the per-factor scale-gradient sums ---------------------------------------------------------------
  block_accumulate(loss_v, acc, red[0]);
  if constexpr (Lead::kHas) block_accumulate((double)dm * (double)d2l, acc + 1 + px.lead_slot, red[1]);
  GM_UNROLL for (int v = 0; v < kMaxVecExtra; ++v)
    if (v < px.nvec) block_accumulate((double)dm * (double)d2v[v], acc + 1 + px.v[v].slot, red[2 + v]);
}

template <class Lead, typename T>
static int launch_product(const Lead& lead, const PairArgs& a) {
  if (a.ps.P <= 0) return 0;
  const int threads = 128;
  long long blocks = (a.ps.P + threads - 1) / threads;
  if (blocks > 0x7fffffffLL) return GM_EINVAL;
  product_pair_kernel<Lead, T><<<(unsigned)blocks, threads, 0, a.stream>>>(
      lead, a.ps, (const T*)a.xa, (T*)a.ga, (T)a.scale_sp, typed_extras<T>(*a.px), a.tg, a.lc, a.acc);
  note_launch();
  return check_launch();
}

}  // namespace gm
