// Per-point manifold operations used by the fused optimizer step and by the
// Manifold API (exp / retr / log / proju / projx / egrad2rgrad / inner / norm /
// transp).  One point per thread.  Each `*Pt` struct exposes the same interface
// on fixed-capacity per-thread arrays `T[CAP]`:
//
//   int  count() const                       number of scalars per point
//   void egrad2rgrad(x, g, out)
//   T    norm2(x, u)                         manifold.norm(x,u)^2 incl. the reference's clamp
//   void retr(x, u, out) / exp(x, u, out)
//   void transp(x, y, u, out)
//   void log(x, y, out), proju(x, u, out), projx(x, out), T inner(x, u, v)
//
// Reference lines restated are cited per struct.
#pragma once
#include "../../include/gm_kernels.h"
#include "gm_manifolds.cuh"

namespace gm {

// ===========================================================================
// SPD (manifolds/spd.py:100-161,196-199); N compile time, registers only.
// FAST_CHOL selects the eps-perturbed closed-form 2x2 factor the reference uses
// for *every* chol/invchol call of SPD(2) (spd.py:43-49).
// ===========================================================================
template <typename T, int N, bool FAST_CHOL>
struct SpdPt {
  static constexpr int CAP = N * N;
  static constexpr bool kStatic = true;
  T wmin, wmax;
  GM_HD constexpr int count() const { return N * N; }

  GM_HD static void symmetrize(const T (&u)[CAP], T (&s)[CAP]) {  // tb.sym, torch_batch.py:25-27
    GM_UNROLL for (int i = 0; i < N; ++i)
      GM_UNROLL for (int j = i; j < N; ++j) {
        T v = (T)0.5 * (u[i * N + j] + u[j * N + i]);
        s[i * N + j] = v; s[j * N + i] = v;
      }
  }
  GM_HD void proju(const T (&x)[CAP], const T (&u)[CAP], T (&out)[CAP]) const { symmetrize(u, out); }  // :119-124

  GM_HD void egrad2rgrad(const T (&x)[CAP], const T (&g)[CAP], T (&out)[CAP]) const {  // :134-135, x sym(g) x
    T s[CAP];
    symmetrize(g, s);
    congr_full<T, N>(x, s, out);
  }
  GM_HD T norm2(const T (&x)[CAP], const T (&u)[CAP]) const {  // :113-117 (no clamp in this override)
    InvChol<T, N, FAST_CHOL> ic;
    ic.run(x);
    T t[CAP], m[CAP];
    // full (not mirrored) a u a^T, as einsum computes it, then sum of squares of all entries
    GM_UNROLL for (int i = 0; i < N; ++i)
      GM_UNROLL for (int j = 0; j < N; ++j) {
        T s = (T)0;
        GM_UNROLL for (int k = 0; k <= i; ++k) s += ic.a[i * N + k] * u[k * N + j];
        t[i * N + j] = s;
      }
    T acc = (T)0;
    GM_UNROLL for (int i = 0; i < N; ++i)
      GM_UNROLL for (int j = 0; j < N; ++j) {
        T s = (T)0;
        GM_UNROLL for (int k = 0; k <= j; ++k) s += t[i * N + k] * ic.a[j * N + k];
        m[i * N + j] = s;
        acc += s * s;
      }
    return acc;
  }
  // l f(l^-1 u l^-T) l^T with f applied to the eigenvalues (symapply, torch_batch.py:145-166)
  template <int F>  // 0: exp, 1: log
  GM_HD void lfl(const T (&x)[CAP], const T (&u)[CAP], T (&out)[CAP]) const {
    InvChol<T, N, FAST_CHOL> ic;
    ic.run(x);
    T m[CAP], lv[CAP], w[N];
    congr_lower<T, N>(ic.a, u, m);  // torch.symeig(upper=True) reads the upper triangle only
    GM_UNROLL for (int k = 0; k < CAP; ++k) lv[k] = ic.l[k];
    jacobi_eigh<T, N, true, false>(m, lv, w);  // sweeps started from L: lv = L V on exit
    GM_UNROLL for (int k = 0; k < N; ++k) w[k] = (F == 0) ? Num<T>::exp(w[k]) : Num<T>::log(w[k]);
    wdwt<T, N>(lv, w, out);
  }
  GM_HD void exp(const T (&x)[CAP], const T (&u)[CAP], T (&out)[CAP]) const { lfl<0>(x, u, out); }  // :137-144
  GM_HD void log(const T (&x)[CAP], const T (&y)[CAP], T (&out)[CAP]) const { lfl<1>(x, y, out); }  // :156-161
  GM_HD void retr(const T (&x)[CAP], const T (&u)[CAP], T (&out)[CAP]) const {  // :146-154
    InvChol<T, N, FAST_CHOL> ic;
    ic.run(x);
    T w[CAP];  // w = l^-1 u
    GM_UNROLL for (int i = 0; i < N; ++i)
      GM_UNROLL for (int j = 0; j < N; ++j) {
        T s = (T)0;
        GM_UNROLL for (int k = 0; k <= i; ++k) s += ic.a[i * N + k] * u[k * N + j];
        w[i * N + j] = s;
      }
    T y[CAP];
    GM_UNROLL for (int i = 0; i < N; ++i)
      GM_UNROLL for (int j = 0; j < N; ++j) {
        T s = (T)0;
        GM_UNROLL for (int k = 0; k < N; ++k) s += w[k * N + i] * w[k * N + j];
        y[i * N + j] = x[i * N + j] + u[i * N + j] + (T)0.5 * s;
      }
    symmetrize(y, out);
  }
  GM_HD void transp(const T (&x)[CAP], const T (&y)[CAP], const T (&u)[CAP], T (&out)[CAP]) const {  // :196-199
    GM_UNROLL for (int k = 0; k < CAP; ++k) out[k] = u[k];
  }
  GM_HD void projx(const T (&x)[CAP], T (&out)[CAP]) const {  // :126-132
    T s[CAP], v[CAP], w[N];
    symmetrize(x, s);
    jacobi_eigh<T, N, true>(s, v, w);
    GM_UNROLL for (int k = 0; k < N; ++k) w[k] = clampv(w[k], wmin, wmax);
    wdwt<T, N>(v, w, out);
  }
  GM_HD void sqrtm(const T (&x)[CAP], T (&out)[CAP]) const {  // tb.spdsqrtm, torch_batch.py:169-171
    T s[CAP], v[CAP], w[N];
    GM_UNROLL for (int k = 0; k < CAP; ++k) s[k] = x[k];
    jacobi_eigh<T, N, true>(s, v, w);
    GM_UNROLL for (int k = 0; k < N; ++k) w[k] = Num<T>::sqrt(clampv(w[k], wmin, wmax));
    wdwt<T, N>(v, w, out);
  }
  GM_HD T inner(const T (&x)[CAP], const T (&u)[CAP], const T (&v)[CAP]) const {  // :100-106
    InvChol<T, N, FAST_CHOL> ic;
    ic.run(x);
    T xi[CAP], a1[CAP], a2[CAP];
    GM_UNROLL for (int i = 0; i < N; ++i)
      GM_UNROLL for (int j = 0; j < N; ++j) {
        T s = (T)0;
        GM_UNROLL for (int k = (i > j ? i : j); k < N; ++k) s += ic.a[k * N + i] * ic.a[k * N + j];
        xi[i * N + j] = s;
      }
    matmul<T, N>(xi, u, a1);
    matmul<T, N>(xi, v, a2);
    T tr = (T)0;
    GM_UNROLL for (int i = 0; i < N; ++i)
      GM_UNROLL for (int k = 0; k < N; ++k) tr += a1[i * N + k] * a2[k * N + i];
    return tr;
  }
};

// ===========================================================================
// Vector manifolds, run-time length n <= CAP, arrays in per-thread local memory.
// ===========================================================================
template <typename T, int CAP_>
struct LorentzPt {  // manifolds/lorentz.py
  static constexpr int CAP = CAP_;
  static constexpr bool kStatic = false;
  int n;
  T eps;
  GM_HD int count() const { return n; }
  GM_HD T ldot(const T* u, const T* v) const {  // :101-122
    T s = -(u[0] * v[0]);
    for (int k = 1; k < n; ++k) s += u[k] * v[k];
    return s;
  }
  GM_HD void proju(const T* x, const T* u, T* out) const {  // :39-42
    T d = ldot(x, u);
    for (int k = 0; k < n; ++k) out[k] = u[k] + d * x[k];
  }
  GM_HD void egrad2rgrad(const T* x, const T* g, T* out) const {  // :52-57
    T t[CAP];
    for (int k = 0; k < n; ++k) t[k] = g[k];
    t[0] = -t[0];
    proju(x, t, out);
  }
  GM_HD T norm2(const T* x, const T* u) const { return clamp_min(ldot(u, u), eps); }  // base.py:29-32
  GM_HD void exp(const T* x, const T* u, T* out) const {                              // :59-62
    T un = clamp_min(Num<T>::sqrt(clamp_min(ldot(u, u), (T)0)), eps);
    T ch = Num<T>::cosh(un), sh = Num<T>::sinh(un);
    for (int k = 0; k < n; ++k) out[k] = x[k] * ch + (sh * u[k]) / un;
  }
  GM_HD void retr(const T* x, const T* u, T* out) const { exp(x, u, out); }  // base.py:49-50
  GM_HD void log(const T* x, const T* y, T* out) const {                     // :64-70
    T xy = clamp_max(ldot(x, y), (T)-1);
    T denom = clamp_min(Num<T>::sqrt(xy * xy - (T)1), eps);
    T z = -xy;
    T num = clamp_min(Num<T>::log(z + Num<T>::sqrt(z * z - (T)1)), eps);
    T f = num / denom;
    T t[CAP];
    for (int k = 0; k < n; ++k) t[k] = f * (y[k] + xy * x[k]);
    proju(x, t, out);
  }
  GM_HD void transp(const T* x, const T* y, const T* u, T* out) const {  // :79-82
    T xy = ldot(x, y), uy = ldot(u, y);
    T f = uy / ((T)1 - xy);
    for (int k = 0; k < n; ++k) out[k] = u[k] + f * (x[k] + y[k]);
  }
  GM_HD void projx(const T* x, T* out) const {  // :44-50
    T s = (T)0;
    for (int k = 1; k < n; ++k) { s += x[k] * x[k]; out[k] = x[k]; }
    out[0] = Num<T>::sqrt((T)1 + s);
  }
  GM_HD T inner(const T* x, const T* u, const T* v) const { return ldot(u, v); }
};

template <typename T, int CAP_>
struct SpherePt {  // manifolds/sphere.py
  static constexpr int CAP = CAP_;
  static constexpr bool kStatic = false;
  int n;
  T eps;
  GM_HD int count() const { return n; }
  GM_HD T dot(const T* u, const T* v) const {
    T s = (T)0;
    for (int k = 0; k < n; ++k) s += u[k] * v[k];
    return s;
  }
  GM_HD void proju(const T* x, const T* u, T* out) const {  // :41-44
    T d = dot(x, u);
    for (int k = 0; k < n; ++k) out[k] = u[k] - d * x[k];
  }
  GM_HD void egrad2rgrad(const T* x, const T* g, T* out) const { proju(x, g, out); }
  GM_HD T norm2(const T* x, const T* u) const { return clamp_min(dot(u, u), eps); }
  GM_HD void projx(const T* x, T* out) const {  // :46-49
    T nn = Num<T>::sqrt(norm2(x, x));
    for (int k = 0; k < n; ++k) out[k] = x[k] / nn;
  }
  GM_HD void retr(const T* x, const T* u, T* out) const {  // :58-59
    T t[CAP];
    for (int k = 0; k < n; ++k) t[k] = x[k] + u[k];
    projx(t, out);
  }
  GM_HD void exp(const T* x, const T* u, T* out) const {  // :51-56
    T nu = Num<T>::sqrt(norm2(x, u));
    if (nu > eps) {
      T c = Num<T>::cos(nu), s = Num<T>::sin(nu);
      for (int k = 0; k < n; ++k) out[k] = x[k] * c + (u[k] * s) / nu;
    } else {
      retr(x, u, out);
    }
  }
  GM_HD void log(const T* x, const T* y, T* out) const {  // :61-66
    T t[CAP], u[CAP];
    for (int k = 0; k < n; ++k) t[k] = y[k] - x[k];
    proju(x, t, u);
    T one_m = (T)(1.0 - 1e-16);
    T s = clampv(dot(x, y), -one_m, one_m);
    T d = clamp_min(Num<T>::acos(s), eps);
    if (d > eps) {
      T nu = Num<T>::sqrt(norm2(x, u));
      for (int k = 0; k < n; ++k) out[k] = (u[k] * d) / nu;
    } else {
      for (int k = 0; k < n; ++k) out[k] = u[k];
    }
  }
  GM_HD void transp(const T* x, const T* y, const T* u, T* out) const { proju(y, u, out); }  // base.py:65-66
  GM_HD T inner(const T* x, const T* u, const T* v) const { return dot(u, v); }
};

template <typename T, int CAP_>
struct EuclideanPt {  // manifolds/euclidean.py
  static constexpr int CAP = CAP_;
  static constexpr bool kStatic = false;
  int n;
  T eps;
  GM_HD int count() const { return n; }
  GM_HD void proju(const T* x, const T* u, T* out) const { for (int k = 0; k < n; ++k) out[k] = u[k]; }
  GM_HD void egrad2rgrad(const T* x, const T* g, T* out) const { proju(x, g, out); }
  GM_HD T norm2(const T* x, const T* u) const {
    T s = (T)0;
    for (int k = 0; k < n; ++k) s += u[k] * u[k];
    return clamp_min(s, eps);
  }
  GM_HD void exp(const T* x, const T* u, T* out) const { for (int k = 0; k < n; ++k) out[k] = x[k] + u[k]; }
  GM_HD void retr(const T* x, const T* u, T* out) const { exp(x, u, out); }
  GM_HD void log(const T* x, const T* y, T* out) const { for (int k = 0; k < n; ++k) out[k] = y[k] - x[k]; }
  GM_HD void transp(const T* x, const T* y, const T* u, T* out) const { proju(y, u, out); }
  GM_HD void projx(const T* x, T* out) const { for (int k = 0; k < n; ++k) out[k] = x[k]; }
  GM_HD T inner(const T* x, const T* u, const T* v) const {
    T s = (T)0;
    for (int k = 0; k < n; ++k) s += u[k] * v[k];
    return s;
  }
};

// kappa-stereographic "Universal" manifold (manifolds/universal.py over manifolds/impl/math.py); the curvature
// parameter c = get_c() is read from device memory (it is itself being optimised).
template <typename T, int CAP_>
struct UniversalPt {
  static constexpr int CAP = CAP_;
  static constexpr bool kStatic = false;
  int n;
  T eps;
  const T* c_dev;
  T ball_eps;  // BALL_EPS[dtype]: 4e-3 (fp32) / 1e-5 (fp64), math.py:16
  GM_HD int count() const { return n; }
  GM_HD T c() const { return *c_dev; }
  GM_HD T sq(const T* u) const {
    T s = (T)0;
    for (int k = 0; k < n; ++k) s += u[k] * u[k];
    return s;
  }
  GM_HD void proju(const T* x, const T* u, T* out) const { for (int k = 0; k < n; ++k) out[k] = u[k]; }  // :50-51
  GM_HD void projx(const T* x, T* out) const { Kappa<T>::project(x, n, c(), ball_eps, out); }            // :53-57
  GM_HD void egrad2rgrad(const T* x, const T* g, T* out) const {  // :59-60, math.py:1452-1453
    T lam = Kappa<T>::lambda_x(sq(x), c());
    T l2 = lam * lam;
    for (int k = 0; k < n; ++k) out[k] = g[k] / l2;
  }
  // Universal.norm calls math.norm WITHOUT c (universal.py:44-48) => lambda_x is evaluated with the default c = 1.0
  GM_HD T norm2(const T* x, const T* u) const {
    T nn = Kappa<T>::lambda_x(sq(x), (T)1) * Num<T>::sqrt(sq(u));
    return nn * nn;
  }
  GM_HD T inner(const T* x, const T* u, const T* v) const {  // :41-42, math.py:225-228
    T lam = Kappa<T>::lambda_x(sq(x), c());
    T s = (T)0;
    for (int k = 0; k < n; ++k) s += u[k] * v[k];
    return lam * lam * s;
  }
  GM_HD void exp(const T* x, const T* u, T* out) const {  // :62-67, math.py:720-727 then project
    const T cc = c();
    T sc = Num<T>::sqrt(Num<T>::abs(cc));
    T un = clamp_min(Num<T>::sqrt(sq(u)), (T)Kappa<T>::kMinNorm);
    T f = Kappa<T>::tan_func(sc / (T)2 * Kappa<T>::lambda_x(sq(x), cc) * un, cc);
    T second[CAP], g1[CAP];
    for (int k = 0; k < n; ++k) second[k] = f * u[k] / (sc * un);
    Kappa<T>::mobius_add(x, second, n, cc, g1);
    Kappa<T>::project(g1, n, cc, ball_eps, out);
  }
  GM_HD void retr(const T* x, const T* u, T* out) const {  // :69-70
    T t[CAP];
    for (int k = 0; k < n; ++k) t[k] = x[k] + u[k];
    Kappa<T>::project(t, n, c(), ball_eps, out);
  }
  GM_HD void log(const T* x, const T* y, T* out) const {  // :72-73, math.py:835-841
    const T cc = c();
    T nx[CAP], sub[CAP];
    for (int k = 0; k < n; ++k) nx[k] = -x[k];
    Kappa<T>::mobius_add(nx, y, n, cc, sub);
    T sn = clamp_min(Num<T>::sqrt(sq(sub)), (T)Kappa<T>::kMinNorm);
    T lam = Kappa<T>::lambda_x(sq(x), cc);
    T sc = Num<T>::sqrt(Num<T>::abs(cc));
    T dphi;
    T f = (T)2 / sc / lam * Kappa<T>::arctan_func(sc * sn, cc, dphi);
    for (int k = 0; k < n; ++k) out[k] = f * sub[k] / sn;
  }
  // parallel_transport (math.py:1359-1362): gyr[y, -x] u * lambda_x / lambda_y, gyration simplified as math.py:1282-1298
  GM_HD void transp(const T* x, const T* y, const T* w, T* out) const {  // :83-84
    const T cc = c();
    // u := y, v := -x
    T u2 = sq(y), v2 = sq(x), uv = (T)0, uw = (T)0, vw = (T)0;
    for (int k = 0; k < n; ++k) { uv -= y[k] * x[k]; uw += y[k] * w[k]; vw -= x[k] * w[k]; }
    T c2 = cc * cc;
    T a = -c2 * uw * v2 + cc * vw + (T)2 * c2 * uv * vw;
    T b = -c2 * vw * u2 - cc * uw;
    T d = clamp_min((T)1 + (T)2 * cc * uv + c2 * u2 * v2, (T)Kappa<T>::kMinNorm);
    T ratio_x = Kappa<T>::lambda_x(v2, cc), ratio_y = Kappa<T>::lambda_x(u2, cc);
    for (int k = 0; k < n; ++k) {
      T gyr = w[k] + (T)2 * (a * y[k] + b * (-x[k])) / d;
      out[k] = gyr * ratio_x / ratio_y;
    }
  }
};

// ===========================================================================
// Grassmann Gr(n, P), points are n x P row-major, n <= NMAX (manifolds/grassmann.py).
// Thin SVDs are taken through the P x P Gram matrix (Jacobi, registers): every
// use is of the form  m v f(sigma) v^T  with f(s)/s analytic at 0, so the small-
// singular-value loss of the Gram route is harmless.
// ===========================================================================
template <typename T, int P, int NMAX>
struct GrassmannPt {
  static constexpr int CAP = NMAX * P;
  static constexpr bool kStatic = false;
  int n;
  T eps;
  int retr_qr;
  GM_HD int count() const { return n * P; }

  GM_HD void gram(const T* a, const T* b, T (&c)[P * P]) const {  // a^T b
    GM_UNROLL for (int i = 0; i < P * P; ++i) c[i] = (T)0;
    for (int r = 0; r < n; ++r)
      GM_UNROLL for (int i = 0; i < P; ++i)
        GM_UNROLL for (int j = 0; j < P; ++j) c[i * P + j] += a[r * P + i] * b[r * P + j];
  }
  // out = a c   (n x P times P x P)
  GM_HD void mulr(const T* a, const T (&c)[P * P], T* out) const {
    for (int r = 0; r < n; ++r) {
      T row[P];
      GM_UNROLL for (int j = 0; j < P; ++j) {
        T s = (T)0;
        GM_UNROLL for (int k = 0; k < P; ++k) s += a[r * P + k] * c[k * P + j];
        row[j] = s;
      }
      GM_UNROLL for (int j = 0; j < P; ++j) out[r * P + j] = row[j];
    }
  }
  GM_HD void proju(const T* x, const T* u, T* out) const {  // :49-53
    T c[P * P];
    gram(x, u, c);
    for (int r = 0; r < n; ++r) {
      T row[P];
      GM_UNROLL for (int j = 0; j < P; ++j) {
        T s = (T)0;
        GM_UNROLL for (int k = 0; k < P; ++k) s += x[r * P + k] * c[k * P + j];
        row[j] = u[r * P + j] - s;
      }
      GM_UNROLL for (int j = 0; j < P; ++j) out[r * P + j] = row[j];
    }
  }
  GM_HD void egrad2rgrad(const T* x, const T* g, T* out) const { proju(x, g, out); }
  GM_HD T norm2(const T* x, const T* u) const {
    T s = (T)0;
    for (int k = 0; k < n * P; ++k) s += u[k] * u[k];
    return clamp_min(s, eps);
  }
  // v f(sqrt(w)) v^T for the Gram matrix g = m^T m = v diag(w) v^T
  template <int F>  // 0: cos(s), 1: sin(s)/s, 2: 1/s, 3: atan(s)/s
  GM_HD void gram_fn(const T (&g)[P * P], T (&out)[P * P]) const {
    T a[P * P], v[P * P], w[P];
    GM_UNROLL for (int i = 0; i < P * P; ++i) a[i] = g[i];
    jacobi_eigh<T, P, true>(a, v, w);
    GM_UNROLL for (int k = 0; k < P; ++k) {
      T s = Num<T>::sqrt(clamp_min(w[k], (T)0));
      T tiny = Num<T>::sqrt(Num<T>::tiny);
      if (F == 0) w[k] = Num<T>::cos(s);
      else if (F == 1) w[k] = s > tiny ? Num<T>::sin(s) / s : (T)1;
      else if (F == 2) w[k] = (T)1 / s;
      else w[k] = s > tiny ? Num<T>::atan(s) / s : (T)1;
    }
    wdwt<T, P>(v, w, out);
  }
  GM_HD void exp(const T* x, const T* u, T* out) const {  // :63-69: x v cos(s) v^T + u_s sin(s) v^T
    T g[P * P], c[P * P], s[P * P];
    gram(u, u, g);
    gram_fn<0>(g, c);
    gram_fn<1>(g, s);
    for (int r = 0; r < n; ++r) {
      T row[P];
      GM_UNROLL for (int j = 0; j < P; ++j) {
        T acc = (T)0;
        GM_UNROLL for (int k = 0; k < P; ++k) acc += x[r * P + k] * c[k * P + j] + u[r * P + k] * s[k * P + j];
        row[j] = acc;
      }
      GM_UNROLL for (int j = 0; j < P; ++j) out[r * P + j] = row[j];
    }
  }
  // Q factor of the Householder QR of y (LAPACK geqr2 + org2r sign conventions,
  // which torch.qr exposes): :55-61, :71-74
  GM_HD void qr_q(const T* y, T* q) const {
    T tau[P];
    for (int k = 0; k < n * P; ++k) q[k] = y[k];
    GM_UNROLL for (int k = 0; k < P; ++k) {
      if (k >= n) { tau[k] = (T)0; continue; }
      T alpha = q[k * P + k];
      T xn2 = (T)0;
      for (int r = k + 1; r < n; ++r) xn2 += q[r * P + k] * q[r * P + k];
      if (xn2 == (T)0) { tau[k] = (T)0; continue; }
      T beta = -Num<T>::copysign(Num<T>::sqrt(alpha * alpha + xn2), alpha);
      tau[k] = (beta - alpha) / beta;
      T sc = (T)1 / (alpha - beta);
      for (int r = k + 1; r < n; ++r) q[r * P + k] *= sc;
      q[k * P + k] = beta;
      // apply H_k to the trailing columns
      GM_UNROLL for (int j = 0; j < P; ++j) {
        if (j > k) {
          T w = q[k * P + j];
          for (int r = k + 1; r < n; ++r) w += q[r * P + k] * q[r * P + j];
          w *= tau[k];
          q[k * P + j] -= w;
          for (int r = k + 1; r < n; ++r) q[r * P + j] -= q[r * P + k] * w;
        }
      }
    }
    // form Q in place (org2r)
    GM_UNROLL for (int kk = 0; kk < P; ++kk) {
      const int k = P - 1 - kk;
      if (k >= n) continue;
      GM_UNROLL for (int j = 0; j < P; ++j) {
        if (j > k) {
          T w = q[k * P + j];  // v_k = 1 at row k
          for (int r = k + 1; r < n; ++r) w += q[r * P + k] * q[r * P + j];
          w *= tau[k];
          q[k * P + j] -= w;
          for (int r = k + 1; r < n; ++r) q[r * P + j] -= q[r * P + k] * w;
        }
      }
      for (int r = k + 1; r < n; ++r) q[r * P + k] *= -tau[k];
      q[k * P + k] = (T)1 - tau[k];
      for (int r = 0; r < k; ++r) q[r * P + k] = (T)0;
    }
  }
  GM_HD void projx(const T* x, T* out) const { qr_q(x, out); }
  GM_HD void retr(const T* x, const T* u, T* out) const {  // :71-80
    T y[CAP];
    for (int k = 0; k < n * P; ++k) y[k] = x[k] + u[k];
    if (retr_qr) { qr_q(y, out); return; }
    T g[P * P], c[P * P];
    gram(y, y, g);
    gram_fn<2>(g, c);  // polar factor u v^T = y (y^T y)^{-1/2}
    mulr(y, c, out);
  }
  GM_HD void log(const T* x, const T* y, T* out) const {  // :82-89
    // B = (y - x x^T y)(x^T y)^{-1};  log = B v atan(s)/s v^T with B^T B = v s^2 v^T
    T xty[P * P];
    gram(x, y, xty);
    // invert xty (P x P) by Gauss-Jordan with partial pivoting
    T inv[P * P], m[P * P];
    GM_UNROLL for (int i = 0; i < P * P; ++i) { m[i] = xty[i]; inv[i] = (i / P == i % P) ? (T)1 : (T)0; }
    for (int c = 0; c < P; ++c) {
      int piv = c;
      T best = Num<T>::abs(m[c * P + c]);
      for (int r = c + 1; r < P; ++r)
        if (Num<T>::abs(m[r * P + c]) > best) { best = Num<T>::abs(m[r * P + c]); piv = r; }
      if (piv != c)
        for (int j = 0; j < P; ++j) {
          T t1 = m[c * P + j]; m[c * P + j] = m[piv * P + j]; m[piv * P + j] = t1;
          T t2 = inv[c * P + j]; inv[c * P + j] = inv[piv * P + j]; inv[piv * P + j] = t2;
        }
      T d = (T)1 / m[c * P + c];
      for (int j = 0; j < P; ++j) { m[c * P + j] *= d; inv[c * P + j] *= d; }
      for (int r = 0; r < P; ++r)
        if (r != c) {
          T f = m[r * P + c];
          for (int j = 0; j < P; ++j) { m[r * P + j] -= f * m[c * P + j]; inv[r * P + j] -= f * inv[c * P + j]; }
        }
    }
    T b[CAP];
    for (int r = 0; r < n; ++r) {
      T row[P];
      GM_UNROLL for (int j = 0; j < P; ++j) {
        T s = (T)0;
        GM_UNROLL for (int k = 0; k < P; ++k) s += x[r * P + k] * xty[k * P + j];
        row[j] = y[r * P + j] - s;
      }
      GM_UNROLL for (int j = 0; j < P; ++j) {
        T s = (T)0;
        GM_UNROLL for (int k = 0; k < P; ++k) s += row[k] * inv[k * P + j];
        b[r * P + j] = s;
      }
    }
    // b = us diag(s) vs^T by one-sided (Hestenes) Jacobi on the columns of b -- every singular value to relative
    // accuracy.  (Going through the Gram matrix b^T b squares the condition number: with tan(angle) ~ 30 for nearly
    // orthogonal subspaces the small singular values, and with them atan(s)/s, lose 3 digits in fp32.)
    T v[P * P];
    GM_UNROLL for (int i = 0; i < P * P; ++i) v[i] = (i / P == i % P) ? (T)1 : (T)0;
    for (int sweep = 0; sweep < JacobiCfg<T>::max_sweeps + 4; ++sweep) {
      bool rotated = false;
      GM_UNROLL for (int p = 0; p < P - 1; ++p) {
        GM_UNROLL for (int q = p + 1; q < P; ++q) {
          T alpha = (T)0, beta = (T)0, gamma = (T)0;
          for (int r = 0; r < n; ++r) {
            alpha += b[r * P + p] * b[r * P + p];
            beta += b[r * P + q] * b[r * P + q];
            gamma += b[r * P + p] * b[r * P + q];
          }
          if (Num<T>::abs(gamma) > (Num<T>::eps * (T)0.25) * Num<T>::sqrt(alpha * beta) &&
              Num<T>::abs(gamma) > Num<T>::tiny) {
            rotated = true;
            // exact (not MUFU-approximate) rotation: this is an API op, not the training hot path
            T zeta = (beta - alpha) / (gamma + gamma);
            T t = Num<T>::copysign((T)1, zeta) / (Num<T>::abs(zeta) + Num<T>::sqrt((T)1 + zeta * zeta));
            T cs = (T)1 / Num<T>::sqrt((T)1 + t * t), sn = t * cs;
            for (int r = 0; r < n; ++r) {
              T bp = b[r * P + p], bq = b[r * P + q];
              b[r * P + p] = cs * bp - sn * bq;
              b[r * P + q] = sn * bp + cs * bq;
            }
            GM_UNROLL for (int r = 0; r < P; ++r) {
              T vp = v[r * P + p], vq = v[r * P + q];
              v[r * P + p] = cs * vp - sn * vq;
              v[r * P + q] = sn * vp + cs * vq;
            }
          }
        }
      }
      if (!rotated) break;
    }
    // out = sum_k (b_k / s_k) atan(s_k) v_k^T  with b_k = u_k s_k the rotated columns
    T f[P];
    const T tiny = Num<T>::sqrt(Num<T>::tiny);
    GM_UNROLL for (int k = 0; k < P; ++k) {
      T n2 = (T)0;
      for (int r = 0; r < n; ++r) n2 += b[r * P + k] * b[r * P + k];
      T sk = Num<T>::sqrt(n2);
      f[k] = sk > tiny ? Num<T>::atan(sk) / sk : (T)1;
    }
    T c[P * P];  // c = diag(f) v^T
    GM_UNROLL for (int k = 0; k < P; ++k)
      GM_UNROLL for (int j = 0; j < P; ++j) c[k * P + j] = f[k] * v[j * P + k];
    mulr(b, c, out);
  }
  GM_HD void transp(const T* x, const T* y, const T* u, T* out) const { proju(y, u, out); }  // base.py:65-66
  GM_HD T inner(const T* x, const T* u, const T* v) const {
    T s = (T)0;
    for (int k = 0; k < n * P; ++k) s += u[k] * v[k];
    return s;
  }
};

// ===========================================================================
// Optimizer update of one point (optim/radam.py:43-98, optim/rsgd.py:40-82).
// ===========================================================================
struct OptimCfg {
  int kind, exact, has_clip, step, has_momentum, first_step, zero_grad;
  double lr, beta1, beta2, momentum, dampening, max_grad_norm, eps;
  double alpha;  // RAdam step size lr * (1 - beta2^t)^0.5 / (1 - beta1^t), evaluated once on the host (radam.py:89-91)
};

inline OptimCfg make_optim_cfg(const gm_optim_t* opt) {
  OptimCfg c;
  c.kind = opt->kind; c.exact = opt->exact; c.has_clip = opt->has_clip; c.step = opt->step;
  c.has_momentum = opt->has_momentum; c.first_step = opt->first_step; c.zero_grad = opt->zero_grad;
  c.lr = opt->lr; c.beta1 = opt->beta1; c.beta2 = opt->beta2; c.momentum = opt->momentum;
  c.dampening = opt->dampening; c.max_grad_norm = opt->max_grad_norm; c.eps = opt->eps;
  c.alpha = opt->lr * ::sqrt(1.0 - ::pow(opt->beta2, (double)opt->step)) / (1.0 - ::pow(opt->beta1, (double)opt->step));
  return c;
}

// x, g: current point and Euclidean gradient.  b1/b2: optimizer buffers
// (exp_avg / exp_avg_sq, or the momentum buffer in b1).  All updated in place.
template <class Man, typename T>
GM_HD void optim_update(const Man& man, const OptimCfg& c, T* x, const T* g, T* b1, T* b2) {
  constexpr int CAP = Man::CAP;
  const int cnt = Man::kStatic ? CAP : man.count();
  T rg[CAP], dir[CAP], nx[CAP];
  if (c.kind == GM_OPT_RADAM) {
    man.egrad2rgrad(*(T(*)[CAP])x, *(const T(*)[CAP])g, rg);
    T gn = Num<T>::sqrt(man.norm2(*(T(*)[CAP])x, rg));  // keepdim norm (radam.py:72)
    if (c.has_clip) {
      T f = clamp_max((T)c.max_grad_norm / gn, (T)1);
      for (int k = 0; k < cnt; ++k) rg[k] *= f;
    }
    T gn2 = gn * gn;  // grad_norm.pow_(2): the UNclipped norm (radam.py:87)
    // alpha = lr * (1 - beta2^t)^0.5 / (1 - beta1^t) in Python doubles, then applied in T (radam.py:89-91)
    const T nalpha = (T)(-c.alpha), be1 = (T)c.beta1, om1 = (T)(1.0 - c.beta1), be2 = (T)c.beta2;
    const T add2 = (T)(1.0 - c.beta2) * gn2;
    // The second moment is one scalar per point stored at full parameter shape (radam.py:56-60,87), so all cnt
    // entries normally agree: then one sqrt and one reciprocal serve the whole point.
    bool same = true;
    for (int k = 1; k < cnt; ++k) same = same && (b2[k] == b2[0]);
    if (same) {
      T v = b2[0] * be2 + add2;
      T rden = Num<T>::recip(Num<T>::sqrt(v) + (T)c.eps);
      for (int k = 0; k < cnt; ++k) {
        T m = b1[k] * be1 + om1 * rg[k];
        b1[k] = m;
        b2[k] = v;
        dir[k] = (m * rden) * nalpha;  // denom.div_(exp_avg).reciprocal_().mul_(-alpha)
      }
    } else {
      for (int k = 0; k < cnt; ++k) {
        T m = b1[k] * be1 + om1 * rg[k];
        T v = b2[k] * be2 + add2;
        b1[k] = m;
        b2[k] = v;
        T denom = Num<T>::sqrt(v) + (T)c.eps;
        dir[k] = ((T)1 / (denom / m)) * nalpha;
      }
    }
    if (c.exact) man.exp(*(T(*)[CAP])x, dir, nx); else man.retr(*(T(*)[CAP])x, dir, nx);
    T mt[CAP];
    man.transp(*(T(*)[CAP])x, nx, *(T(*)[CAP])b1, mt);
    for (int k = 0; k < cnt; ++k) { x[k] = nx[k]; b1[k] = mt[k]; }
  } else {
    if (c.has_momentum && c.first_step)
      for (int k = 0; k < cnt; ++k) b1[k] = g[k];  // rsgd.py:53-54: clone of the *Euclidean* gradient
    man.egrad2rgrad(*(T(*)[CAP])x, *(const T(*)[CAP])g, rg);
    if (c.has_clip) {
      T gn = Num<T>::sqrt(man.norm2(*(T(*)[CAP])x, rg));
      T f = clamp_max((T)c.max_grad_norm / gn, (T)1);
      for (int k = 0; k < cnt; ++k) rg[k] *= f;
    }
    if (c.has_momentum) {
      for (int k = 0; k < cnt; ++k) {
        b1[k] = b1[k] * (T)c.momentum + (T)(1.0 - c.dampening) * rg[k];
        dir[k] = (T)(-c.lr) * b1[k];
      }
      if (c.exact) man.exp(*(T(*)[CAP])x, dir, nx); else man.retr(*(T(*)[CAP])x, dir, nx);
      T mt[CAP];
      man.transp(*(T(*)[CAP])x, nx, *(T(*)[CAP])b1, mt);
      for (int k = 0; k < cnt; ++k) { x[k] = nx[k]; b1[k] = mt[k]; }
    } else {
      for (int k = 0; k < cnt; ++k) dir[k] = (T)(-c.lr) * rg[k];
      if (c.exact) man.exp(*(T(*)[CAP])x, dir, nx); else man.retr(*(T(*)[CAP])x, dir, nx);
      for (int k = 0; k < cnt; ++k) x[k] = nx[k];
    }
  }
}

}  // namespace gm
