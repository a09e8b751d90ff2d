// Optimizer step and point operations for Lorentz / Sphere / Euclidean /
// Universal / Grassmann parameters (optim/radam.py:43-98, optim/rsgd.py:40-82 with the
// callees of manifolds/lorentz.py:39-86, sphere.py:41-66, euclidean.py:34-50,
// grassmann.py:46-89).  One point per thread, per-thread arrays of capacity CAP.
#include "gm_point_kernels.cuh"

namespace gm {

template <typename T, template <typename, int> class M>
static int vec_cap(const PointArgs& a) {
  const T eps = (T)1e-8;
  if (a.n <= 16) { M<T, 16> man{a.n, eps}; return launch_point<decltype(man), T>(man, a); }
  if (a.n <= 64) { M<T, 64> man{a.n, eps}; return launch_point<decltype(man), T>(man, a); }
  if (a.n <= 256) { M<T, 256> man{a.n, eps}; return launch_point<decltype(man), T>(man, a); }
  return GM_EUNSUPPORTED;
}

template <typename T, int P>
static int grass_cap(const PointArgs& a) {
  const T eps = (T)1e-8;
  if (a.n <= 16) { GrassmannPt<T, P, 16> man{a.n, eps, a.grassmann_retr_qr}; return launch_point<decltype(man), T>(man, a); }
  return GM_EUNSUPPORTED;
}

template <typename T>
static int universal_cap(const PointArgs& a) {
  const T eps = (T)1e-8;
  const T ball_eps = (T)(sizeof(T) == 4 ? 4e-3 : 1e-5);  // BALL_EPS, manifolds/impl/math.py:16
  const T* c = (const T*)a.c_dev;
  if (a.n <= 16) { UniversalPt<T, 16> man{a.n, eps, c, ball_eps}; return launch_point<decltype(man), T>(man, a); }
  if (a.n <= 64) { UniversalPt<T, 64> man{a.n, eps, c, ball_eps}; return launch_point<decltype(man), T>(man, a); }
  if (a.n <= 256) { UniversalPt<T, 256> man{a.n, eps, c, ball_eps}; return launch_point<decltype(man), T>(man, a); }
  return GM_EUNSUPPORTED;
}

// compiled once per manifold kind: -DGM_PKIND=0 (Lorentz) 1 (Sphere) 2 (Euclidean) 3 (Grassmann) 4 (Universal)
#ifndef GM_PKIND
#error "compile with -DGM_PKIND=<0..4>"
#endif

template <typename T>
static int point_typed(const PointArgs& a) {
#if GM_PKIND == 0
  return vec_cap<T, LorentzPt>(a);
#elif GM_PKIND == 1
  return vec_cap<T, SpherePt>(a);
#elif GM_PKIND == 2
  return vec_cap<T, EuclideanPt>(a);
#elif GM_PKIND == 4
  return universal_cap<T>(a);
#else
  switch (a.p) {
    case 1: return grass_cap<T, 1>(a);
    case 2: return grass_cap<T, 2>(a);
    case 3: return grass_cap<T, 3>(a);
    case 4: return grass_cap<T, 4>(a);
    case 5: return grass_cap<T, 5>(a);
    default: return GM_EUNSUPPORTED;
  }
#endif
}

#define GM_CAT2(a, b) a##b
#define GM_CAT(a, b) GM_CAT2(a, b)
int GM_CAT(vec_point_, GM_PKIND)(const PointArgs& a) {
  if (a.dtype == GM_F32) return point_typed<float>(a);
  if (a.dtype == GM_F64) return point_typed<double>(a);
  return GM_EINVAL;
}

}  // namespace gm
