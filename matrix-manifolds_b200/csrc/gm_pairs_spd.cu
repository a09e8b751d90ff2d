// SPD pair kernels (affine-invariant distance and Stein divergence), one pair
// per thread, the n x n math register resident.  Compiled once per matrix size:
//   nvcc -DGM_N=<n> ... gm_pairs_spd.cu -o gm_pairs_spd_<n>.o
// and exposes gm::spd_launch_<n>(...) to the dispatcher in gm_api.cu.
//
// Replaces SymmetricPositiveDefinite.dist / pdist / stein_div / stein_pdiv and
// their autograd backward (manifolds/spd.py:171-194,246-295), fused with the
// gather x[m[0]], x[m[1]] (base.py:62-63), the loss (objectives.py:16-45) and
// the index_put_ scatter-add of the gradients.
#include "gm_launch.cuh"

#ifndef GM_N
#error "compile with -DGM_N=<matrix size>"
#endif

namespace gm {

template <class Op, typename T, int KMODE>
__global__ void __launch_bounds__(128)
spd_pair_kernel(Op op, PairSpec ps, const T* __restrict__ xa, const T* __restrict__ xb,
                const T* __restrict__ gout, T coef, T* __restrict__ ga, T* __restrict__ gb,
                T* __restrict__ out_d2, TargetSpec tg, LossCfg lc, T scale_sp, double* __restrict__ acc) {
  constexpr int E = Op::E;
  __shared__ double red[2][4];
  long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  bool active = k < ps.P;
  long long ra = -1, rb = -1;
  double loss_v = 0.0, gd2_v = 0.0;
  T gx[E], gy[E];
  if (active) {
    decode_pair(ps, k, ra, rb);
    T x[E], y[E];
    load_row<T, E>(xa, ra, x);
    load_row<T, E>(xb, rb, y);
    if constexpr (KMODE == K_FWD) {
      out_d2[k] = op.dist2(x, y);
    } else {
      T d2 = op.dist2_grad(x, y, gx, gy);
      T w;
      if constexpr (KMODE == K_BWD) {
        w = coef * gout[k];
      } else {
        T g = fetch_target<T>(tg, k, ra, rb);
        T m = scale_sp * d2;
        T dm;
        T lv = loss_term<T>(lc, g, m, dm);
        loss_v = (double)lv;
        gd2_v = (double)dm * (double)d2;
        w = dm * scale_sp;
        if (out_d2) out_d2[k] = d2;
      }
      GM_UNROLL for (int e = 0; e < E; ++e) { gx[e] *= w; gy[e] *= w; }
    }
  }
  if constexpr (KMODE != K_FWD) {
    if (ps.mode == GM_PAIRS_ELEMENTWISE) {
      if (active) {
        store_row<T, E>(ga, ra, gx);
        store_row<T, E>(gb, rb, gy);
      }
    } else {
      warp_accumulate_row<T, E>(ga, ra, gx);
      warp_accumulate_row<T, E>(gb, rb, gy);
    }
  }
  if constexpr (KMODE == K_FUSED) {
    block_accumulate(loss_v, acc, red[0]);
    block_accumulate(gd2_v, acc + 1, red[1]);
  }
}

template <class Op, typename T>
static int launch_op(const Op& op, const PairArgs& a) {
  if (a.ps.P <= 0) return 0;
  const int threads = 128;
  long long blocks = (a.ps.P + threads - 1) / threads;
  if (blocks > 0x7fffffffLL) return GM_EINVAL;
  dim3 grid((unsigned)blocks), block(threads);
  const T* xa = (const T*)a.xa;
  const T* xb = (const T*)a.xb;
  switch (a.kmode) {
    case K_FWD:
      spd_pair_kernel<Op, T, K_FWD><<<grid, block, 0, a.stream>>>(
          op, a.ps, xa, xb, nullptr, (T)0, nullptr, nullptr, (T*)a.out_d2, a.tg, a.lc, (T)0, nullptr);
      break;
    case K_BWD:
      spd_pair_kernel<Op, T, K_BWD><<<grid, block, 0, a.stream>>>(
          op, a.ps, xa, xb, (const T*)a.gout, (T)a.coef, (T*)a.ga, (T*)a.gb, nullptr, a.tg, a.lc, (T)0, nullptr);
      break;
    default:
      spd_pair_kernel<Op, T, K_FUSED><<<grid, block, 0, a.stream>>>(
          op, a.ps, xa, xb, nullptr, (T)0, (T*)a.ga, (T*)a.gb, (T*)a.out_d2, a.tg, a.lc, (T)a.scale_sp, a.acc);
  }
  note_launch();
  return check_launch();
}

template <typename T>
static int launch_typed(const PairArgs& a) {
  constexpr int N = GM_N;
  const bool fe = (a.flags & GM_FAST_EIG) != 0;
  const bool fc = (a.flags & GM_FAST_CHOL) != 0;
  if (a.kind == GM_SPD_AI) {
    if constexpr (N == 2) {
      if (fe && fc) { SpdAI<T, 2, true, true> op{(T)a.wmin, (T)a.wmax}; return launch_op<decltype(op), T>(op, a); }
      if (fe && !fc) { SpdAI<T, 2, true, false> op{(T)a.wmin, (T)a.wmax}; return launch_op<decltype(op), T>(op, a); }
      if (!fe && fc) { SpdAI<T, 2, false, true> op{(T)a.wmin, (T)a.wmax}; return launch_op<decltype(op), T>(op, a); }
    }
    if constexpr (N == 3) {
      if (fe) { SpdAI<T, 3, true, false> op{(T)a.wmin, (T)a.wmax}; return launch_op<decltype(op), T>(op, a); }
    }
    SpdAI<T, N, false, false> op{(T)a.wmin, (T)a.wmax};
    return launch_op<decltype(op), T>(op, a);
  }
  if (a.kind == GM_SPD_STEIN) {
    if constexpr (N == 2) {
      if (fc) { SpdStein<T, 2, true> op{(T)a.wmin, (T)a.wmax}; return launch_op<decltype(op), T>(op, a); }
    }
    SpdStein<T, N, false> op{(T)a.wmin, (T)a.wmax};
    return launch_op<decltype(op), T>(op, a);
  }
  return GM_EINVAL;
}

#define GM_CAT2(a, b) a##b
#define GM_CAT(a, b) GM_CAT2(a, b)
int GM_CAT(spd_launch_, GM_N)(const PairArgs& a) {
  if (a.dtype == GM_F32) return launch_typed<float>(a);
  if (a.dtype == GM_F64) return launch_typed<double>(a);
  return GM_EINVAL;
}

}  // namespace gm
