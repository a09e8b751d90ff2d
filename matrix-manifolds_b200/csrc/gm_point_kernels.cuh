// Kernels shared by gm_point_spd.cu / gm_point_vec.cu: one point per thread,
// fused optimizer update and the single-op Manifold API calls.
#pragma once
#include "gm_launch.cuh"
#include "gm_pointops.cuh"

namespace gm {

struct PointArgs {
  int kind, dtype, n, p;
  unsigned flags;
  double wmin, wmax;
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
};

template <class Man, typename T>
__global__ void __launch_bounds__(128)
optim_kernel(Man man, OptimCfg oc, T* __restrict__ x, const T* __restrict__ grad, T* __restrict__ buf1,
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
  if (a.op < 0)
    optim_kernel<Man, T><<<(unsigned)blocks, threads, 0, a.stream>>>(man, a.oc, (T*)a.x, (const T*)a.u, (T*)a.buf1,
                                                                     (T*)a.buf2, a.N);
  else
    point_op_kernel<Man, T><<<(unsigned)blocks, threads, 0, a.stream>>>(man, a.op, (const T*)a.x, (const T*)a.u,
                                                                        (const T*)a.v, (T*)a.out, a.N);
  note_launch();
  return check_launch();
}

}  // namespace gm
