// SPD optimizer step and Manifold-API point operations, one point per thread,
// compiled once per matrix size (-DGM_N=<n>).  Replaces RiemannianAdam._step /
// RiemannianSGD._step (optim/radam.py:43-98, optim/rsgd.py:40-82) on SPD
// parameters together with their callees egrad2rgrad / norm / exp / retr /
// transp (manifolds/spd.py:113-154,196-199), and log / projx / inner.
#include "gm_point_kernels.cuh"

#ifndef GM_N
#error "compile with -DGM_N=<matrix size>"
#endif

namespace gm {

template <typename T>
static int point_typed(const PointArgs& a) {
  constexpr int N = GM_N;
  if constexpr (N == 2) {
    if (a.flags & GM_FAST_CHOL) {
      SpdPt<T, 2, true> man{(T)a.wmin, (T)a.wmax};
      return launch_point<decltype(man), T>(man, a);
    }
  }
  SpdPt<T, N, false> man{(T)a.wmin, (T)a.wmax};
  return launch_point<decltype(man), T>(man, a);
}

#define GM_CAT2(a, b) a##b
#define GM_CAT(a, b) GM_CAT2(a, b)
int GM_CAT(spd_point_, GM_N)(const PointArgs& a) {
  if (a.dtype == GM_F32) return point_typed<float>(a);
  if (a.dtype == GM_F64) return point_typed<double>(a);
  return GM_EINVAL;
}

}  // namespace gm
