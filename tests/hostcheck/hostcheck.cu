// DEV/TEST HARNESS -- not part of the product.  Compiles the *device* math of
// matrix-manifolds_b200/csrc/gm_manifolds.cuh for the host (all of it is
// __host__ __device__) so that the per-pair arithmetic the CUDA kernels execute
// can be compared with the oracle in a container without a GPU
// (tests/test_hostcheck_math.py, `-m "not gpu"`).  The shipped library
// (libgm_b200.so) never links or calls this file.
#include "../../matrix-manifolds_b200/csrc/gm_manifolds.cuh"
#include "../../include/gm_kernels.h"

using namespace gm;

template <class Op, typename T>
static void run_rows(const Op& op, const T* x, const T* y, long P, T* d2, T* gx, T* gy) {
  constexpr int E = Op::E;
  for (long k = 0; k < P; ++k) {
    T xr[E], yr[E], gxr[E], gyr[E];
    for (int e = 0; e < E; ++e) { xr[e] = x[k * E + e]; yr[e] = y[k * E + e]; }
    if (gx) {
      d2[k] = op.dist2_grad(xr, yr, gxr, gyr);
      for (int e = 0; e < E; ++e) { gx[k * E + e] = gxr[e]; gy[k * E + e] = gyr[e]; }
    } else {
      d2[k] = op.dist2(xr, yr);
    }
  }
}

template <typename T, int N>
static int spd_n(int kind, unsigned flags, double wmin, double wmax, const T* x, const T* y, long P, T* d2, T* gx,
                 T* gy) {
  const bool fe = flags & GM_FAST_EIG, fc = flags & GM_FAST_CHOL;
  if (kind == GM_SPD_AI) {
    if constexpr (N == 2) {
      if (fe && fc) { SpdAI<T, 2, true, true> op{(T)wmin, (T)wmax}; run_rows(op, x, y, P, d2, gx, gy); return 0; }
      if (fe) { SpdAI<T, 2, true, false> op{(T)wmin, (T)wmax}; run_rows(op, x, y, P, d2, gx, gy); return 0; }
      if (fc) { SpdAI<T, 2, false, true> op{(T)wmin, (T)wmax}; run_rows(op, x, y, P, d2, gx, gy); return 0; }
    }
    if constexpr (N == 3) {
      if (fe) { SpdAI<T, 3, true, false> op{(T)wmin, (T)wmax}; run_rows(op, x, y, P, d2, gx, gy); return 0; }
    }
    SpdAI<T, N, false, false> op{(T)wmin, (T)wmax};
    run_rows(op, x, y, P, d2, gx, gy);
    return 0;
  }
  if constexpr (N == 2) {
    if (fc) { SpdStein<T, 2, true> op{(T)wmin, (T)wmax}; run_rows(op, x, y, P, d2, gx, gy); return 0; }
  }
  SpdStein<T, N, false> op{(T)wmin, (T)wmax};
  run_rows(op, x, y, P, d2, gx, gy);
  return 0;
}

template <typename T>
static int spd_t(int kind, int n, unsigned flags, double wmin, double wmax, const void* x, const void* y, long P,
                 void* d2, void* gx, void* gy) {
#define CASE(N) case N: return spd_n<T, N>(kind, flags, wmin, wmax, (const T*)x, (const T*)y, P, (T*)d2, (T*)gx, (T*)gy);
  switch (n) { CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) default: return -2; }
#undef CASE
}

template <typename T, int KIND>
static void vec_rows(int n, const T* x, const T* y, long P, T* d2, T* gx, T* gy) {
  VecMan<T, KIND> op{(T)1e-8, (T)(1.0 - 1e-16), nullptr};
  for (long k = 0; k < P; ++k) {
    VecCoef<T> c{};
    d2[k] = op.value(x + k * n, y + k * n, n, c);
    if (gx)
      for (int e = 0; e < n; ++e) op.grad_elem(e, x[k * n + e], y[k * n + e], c, gx[k * n + e], gy[k * n + e]);
  }
}

template <typename T, int P_, bool FAST>
static void grass_rows(int n, const T* x, const T* y, long P, T* d2, T* gx, T* gy) {
  GrassmannCore<T, P_, FAST> op{(T)(1.0 - 1e-16)};
  for (long k = 0; k < P; ++k) {
    const T* px = x + k * n * P_;
    const T* py = y + k * n * P_;
    T a[P_ * P_], ga[P_ * P_];
    for (int i = 0; i < P_ * P_; ++i) a[i] = 0;
    for (int r = 0; r < n; ++r)
      for (int i = 0; i < P_; ++i)
        for (int j = 0; j < P_; ++j) a[i * P_ + j] += px[r * P_ + i] * py[r * P_ + j];
    d2[k] = op.run(a, gx != nullptr, ga);
    if (gx)
      for (int r = 0; r < n; ++r)
        for (int i = 0; i < P_; ++i) {
          T sx = 0, sy = 0;
          for (int j = 0; j < P_; ++j) { sx += py[r * P_ + j] * ga[i * P_ + j]; sy += px[r * P_ + j] * ga[j * P_ + i]; }
          gx[(k * n + r) * P_ + i] = sx;
          gy[(k * n + r) * P_ + i] = sy;
        }
  }
}

template <typename T>
static int any_t(int kind, int n, int p, unsigned flags, double wmin, double wmax, const void* x, const void* y,
                 long P, void* d2, void* gx, void* gy) {
  const T* X = (const T*)x; const T* Y = (const T*)y;
  T* D = (T*)d2; T* GX = (T*)gx; T* GY = (T*)gy;
  switch (kind) {
    case GM_SPD_AI: case GM_SPD_STEIN: return spd_t<T>(kind, n, flags, wmin, wmax, x, y, P, d2, gx, gy);
    case GM_LORENTZ: vec_rows<T, VEC_LORENTZ>(n, X, Y, P, D, GX, GY); return 0;
    case GM_SPHERE: vec_rows<T, VEC_SPHERE>(n, X, Y, P, D, GX, GY); return 0;
    case GM_EUCLIDEAN: vec_rows<T, VEC_EUCLIDEAN>(n, X, Y, P, D, GX, GY); return 0;
    case GM_GRASSMANN:
      if (p == 2 && (flags & GM_FAST_SVD)) { grass_rows<T, 2, true>(n, X, Y, P, D, GX, GY); return 0; }
      if (p == 1) { grass_rows<T, 1, false>(n, X, Y, P, D, GX, GY); return 0; }
      if (p == 2) { grass_rows<T, 2, false>(n, X, Y, P, D, GX, GY); return 0; }
      if (p == 3) { grass_rows<T, 3, false>(n, X, Y, P, D, GX, GY); return 0; }
      if (p == 4) { grass_rows<T, 4, false>(n, X, Y, P, D, GX, GY); return 0; }
      return -2;
  }
  return -1;
}

// Universal (kappa-stereographic): curvature passed by value; gc[k] = d(d2_k)/dc
template <typename T>
static void universal_rows(int n, double c, double wmin, const T* x, const T* y, long P, T* d2, T* gx, T* gy, T* gc) {
  T cc = (T)c;
  VecMan<T, VEC_UNIVERSAL> op{(T)wmin, (T)(1.0 - 1e-16), &cc};
  for (long k = 0; k < P; ++k) {
    VecCoef<T> co{};
    d2[k] = op.value(x + k * n, y + k * n, n, co);
    if (gx)
      for (int e = 0; e < n; ++e) op.grad_elem(e, x[k * n + e], y[k * n + e], co, gx[k * n + e], gy[k * n + e]);
    if (gc) gc[k] = co.dc;
  }
}
extern "C" int hc_universal_pairs(int dtype, int n, double c, double wmin, const void* x, const void* y, long P,
                                  void* d2, void* gx, void* gy, void* gc) {
  if (dtype == GM_F32) universal_rows<float>(n, c, wmin, (const float*)x, (const float*)y, P, (float*)d2, (float*)gx, (float*)gy, (float*)gc);
  else universal_rows<double>(n, c, wmin, (const double*)x, (const double*)y, P, (double*)d2, (double*)gx, (double*)gy, (double*)gc);
  return 0;
}
static double g_universal_c = 1.0;  // hc_point(kind = GM_UNIVERSAL) reads the curvature from here
extern "C" void hc_set_universal_c(double c) { g_universal_c = c; }

extern "C" int hc_pairs(int kind, int dtype, int n, int p, unsigned flags, double wmin, double wmax, const void* x,
                        const void* y, long P, void* d2, void* gx, void* gy) {
  if (dtype == GM_F32) return any_t<float>(kind, n, p, flags, wmin, wmax, x, y, P, d2, gx, gy);
  return any_t<double>(kind, n, p, flags, wmin, wmax, x, y, P, d2, gx, gy);
}

extern "C" void hc_loss(int dtype, int kind, int inc_l1, int inc_l2, double alpha, double eps, const void* g,
                        const void* m, long P, void* val, void* dm) {
  LossCfg c{kind, inc_l1, inc_l2, alpha, eps};
  for (long k = 0; k < P; ++k) {
    if (dtype == GM_F32) ((float*)val)[k] = loss_term<float>(c, ((const float*)g)[k], ((const float*)m)[k], ((float*)dm)[k]);
    else ((double*)val)[k] = loss_term<double>(c, ((const double*)g)[k], ((const double*)m)[k], ((double*)dm)[k]);
  }
}

// ---------------------------------------------------------------------------
// per-point ops and optimizer update (gm_pointops.cuh), same dispatch as
// gm_point_spd.cu / gm_point_vec.cu but looping on the host
// ---------------------------------------------------------------------------
#include "../../matrix-manifolds_b200/csrc/gm_pointops.cuh"

struct HcPoint {
  int op;  // gm_point_op or -1 (optimizer)
  OptimCfg oc;
  void* x; const void* u; const void* v; void* out; void* b1; void* b2;
  long N;
};

template <class Man, typename T>
static int pt_rows(const Man& man, const HcPoint& a) {
  constexpr int CAP = Man::CAP;
  const int cnt = man.count();
  for (long k = 0; k < a.N; ++k) {
    T xs[CAP], us[CAP], vs[CAP], os[CAP], b1[CAP], b2[CAP];
    for (int e = 0; e < cnt; ++e) {
      xs[e] = ((T*)a.x)[k * cnt + e];
      us[e] = a.u ? ((const T*)a.u)[k * cnt + e] : (T)0;
      vs[e] = a.v ? ((const T*)a.v)[k * cnt + e] : (T)0;
      b1[e] = a.b1 ? ((T*)a.b1)[k * cnt + e] : (T)0;
      b2[e] = a.b2 ? ((T*)a.b2)[k * cnt + e] : (T)0;
    }
    if (a.op < 0) {
      optim_update<Man, T>(man, a.oc, xs, us, b1, b2);
      for (int e = 0; e < cnt; ++e) {
        ((T*)a.x)[k * cnt + e] = xs[e];
        if (a.b1) ((T*)a.b1)[k * cnt + e] = b1[e];
        if (a.b2) ((T*)a.b2)[k * cnt + e] = b2[e];
      }
      continue;
    }
    bool scalar = false; T sval = 0;
    switch (a.op) {
      case GM_OP_EXP: man.exp(xs, us, os); break;
      case GM_OP_RETR: man.retr(xs, us, os); break;
      case GM_OP_LOG: man.log(xs, us, os); break;
      case GM_OP_PROJU: man.proju(xs, us, os); break;
      case GM_OP_PROJX: man.projx(xs, os); break;
      case GM_OP_EGRAD2RGRAD: man.egrad2rgrad(xs, us, os); break;
      case GM_OP_INNER: scalar = true; sval = man.inner(xs, us, vs); break;
      case GM_OP_NORM2: scalar = true; sval = man.norm2(xs, us); break;
      case GM_OP_TRANSP: man.transp(xs, us, vs, os); break;
      default: return -1;
    }
    if (scalar) ((T*)a.out)[k] = sval;
    else for (int e = 0; e < cnt; ++e) ((T*)a.out)[k * cnt + e] = os[e];
  }
  return 0;
}

template <typename T, int N>
static int pt_spd(unsigned flags, double wmin, double wmax, const HcPoint& a) {
  if constexpr (N == 2) {
    if (flags & GM_FAST_CHOL) { SpdPt<T, 2, true> m{(T)wmin, (T)wmax}; return pt_rows<decltype(m), T>(m, a); }
  }
  SpdPt<T, N, false> m{(T)wmin, (T)wmax};
  return pt_rows<decltype(m), T>(m, a);
}

template <typename T>
static int pt_any(int kind, int n, int p, unsigned flags, double wmin, double wmax, int retr_qr, const HcPoint& a) {
  const T eps = (T)1e-8;
  switch (kind) {
    case GM_SPD_AI: case GM_SPD_STEIN:
      switch (n) {
        case 1: return pt_spd<T, 1>(flags, wmin, wmax, a);
        case 2: return pt_spd<T, 2>(flags, wmin, wmax, a);
        case 3: return pt_spd<T, 3>(flags, wmin, wmax, a);
        case 4: return pt_spd<T, 4>(flags, wmin, wmax, a);
        case 5: return pt_spd<T, 5>(flags, wmin, wmax, a);
        case 6: return pt_spd<T, 6>(flags, wmin, wmax, a);
        default: return -2;
      }
    case GM_LORENTZ: { LorentzPt<T, 64> m{n, eps}; return pt_rows<decltype(m), T>(m, a); }
    case GM_SPHERE: { SpherePt<T, 64> m{n, eps}; return pt_rows<decltype(m), T>(m, a); }
    case GM_EUCLIDEAN: { EuclideanPt<T, 64> m{n, eps}; return pt_rows<decltype(m), T>(m, a); }
    case GM_UNIVERSAL: {
      T cc = (T)g_universal_c;
      UniversalPt<T, 64> m{n, eps, &cc, (T)(sizeof(T) == 4 ? 4e-3 : 1e-5)};
      return pt_rows<decltype(m), T>(m, a);
    }
    case GM_GRASSMANN:
      switch (p) {
        case 1: { GrassmannPt<T, 1, 16> m{n, eps, retr_qr}; return pt_rows<decltype(m), T>(m, a); }
        case 2: { GrassmannPt<T, 2, 16> m{n, eps, retr_qr}; return pt_rows<decltype(m), T>(m, a); }
        case 3: { GrassmannPt<T, 3, 16> m{n, eps, retr_qr}; return pt_rows<decltype(m), T>(m, a); }
        case 4: { GrassmannPt<T, 4, 16> m{n, eps, retr_qr}; return pt_rows<decltype(m), T>(m, a); }
        default: return -2;
      }
  }
  return -1;
}

extern "C" int hc_point(int kind, int dtype, int n, int p, unsigned flags, double wmin, double wmax, int retr_qr,
                        int op, const gm_optim_t* opt, void* x, const void* u, const void* v, void* out, void* b1,
                        void* b2, long N) {
  HcPoint a{};
  a.op = op; a.x = x; a.u = u; a.v = v; a.out = out; a.b1 = b1; a.b2 = b2; a.N = N;
  if (op < 0) {
    a.oc = make_optim_cfg(opt);
    retr_qr = opt->grassmann_retr_qr;
  }
  if (dtype == GM_F32) return pt_any<float>(kind, n, p, flags, wmin, wmax, retr_qr, a);
  return pt_any<double>(kind, n, p, flags, wmin, wmax, retr_qr, a);
}
