// DEV TOOL (host only, never shipped): measures Jacobi sweep counts and fp32 accuracy of SPD pair-math variants
// against an fp64 evaluation of the same formulas, for the distributions the bench and the tests use.
//   nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr -w -o /tmp/eig_lab tools/eig_lab.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <random>
#include <vector>
#include <algorithm>
#include "../matrix-manifolds_b200/csrc/gm_manifolds.cuh"
using namespace gm;

static std::mt19937_64 rng(1234);
static std::normal_distribution<double> nd(0.0, 1.0);

static void rand_spd(double ir, double (&x)[16]) {
  double s[16];
  double nrm = 0;
  for (int i = 0; i < 4; ++i) for (int j = i; j < 4; ++j) { double u = nd(rng); s[i*4+j] = s[j*4+i] = u; nrm += u*u; }
  nrm = std::sqrt(nrm);
  for (int k = 0; k < 16; ++k) s[k] *= ir / nrm;
  double v[16], w[4];
  jacobi_eigh<double, 4, true>(s, v, w);
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) {
    double a = 0; for (int k = 0; k < 4; ++k) a += v[i*4+k] * std::exp(w[k]) * v[j*4+k];
    x[i*4+j] = a;
  }
}

struct Cfg { float thr_mult; int fixed; int max_sweeps; int order = 0; float cnoise = 0; int first_check = 0; };
static std::uniform_real_distribution<float> ud(-1.f, 1.f);
static int g_last_sweeps;

// instrumented restatement of gm::jacobi_eigh<float,4>
static void jac(const Cfg& cfg, float (&a)[16], float (&v)[16], float (&w)[4]) {
  constexpr int N = 4;
  for (int i = 0; i < 16; ++i) v[i] = (i % 5 == 0) ? 1.f : 0.f;
  int sweep = 0;
  for (; sweep < cfg.max_sweeps; ++sweep) {
    if (!cfg.fixed && sweep >= cfg.first_check) {
      float off = 0, dia = 0;
      for (int i = 0; i < N; ++i) { dia += a[i*N+i]*a[i*N+i]; for (int j = i+1; j < N; ++j) off += a[i*N+j]*a[i*N+j]; }
      float e = FLT_EPSILON * cfg.thr_mult;
      if (off <= e*e*0.0625f*dia || off < FLT_MIN) break;
    } else if (cfg.fixed && sweep >= cfg.fixed) break;
    static const int ord[2][6][2] = {{{0,1},{0,2},{0,3},{1,2},{1,3},{2,3}}, {{0,1},{2,3},{0,2},{1,3},{0,3},{1,2}}};
    for (int rot = 0; rot < 6; ++rot) { int p = ord[cfg.order][rot][0], q = ord[cfg.order][rot][1];
      float apq = a[p*N+q];
      float d = a[q*N+q] - a[p*N+p];
      float two = apq + apq;
      float den = fabsf(d) + sqrtf(d*d + two*two);
      float t = den > 0 ? (d >= 0 ? two : -two) / den : 0.f;
      float c = 1.0f / sqrtf(t*t + 1.f); if (cfg.cnoise > 0) c *= 1.f + cfg.cnoise * ud(rng);
      float s = t * c;
      a[p*N+p] -= t*apq; a[q*N+q] += t*apq; a[p*N+q] = a[q*N+p] = 0;
      for (int r = 0; r < N; ++r) if (r != p && r != q) {
        float arp = a[r*N+p], arq = a[r*N+q];
        float nrp = c*arp - s*arq, nrq = s*arp + c*arq;
        a[r*N+p] = a[p*N+r] = nrp; a[r*N+q] = a[q*N+r] = nrq;
      }
      for (int r = 0; r < N; ++r) { float vp = v[r*N+p], vq = v[r*N+q]; v[r*N+p] = c*vp - s*vq; v[r*N+q] = s*vp + c*vq; }
    }
  }
  g_last_sweeps = sweep;
  for (int i = 0; i < N; ++i) w[i] = a[i*N+i];
}

static float pair_f32(const Cfg& cfg, const float (&x)[16], const float (&y)[16], float (&gx)[16], float (&gy)[16]) {
  InvChol<float, 4, false> ic; ic.run(x);
  float m[16]; congr_lower<float, 4>(ic.a, y, m);
  float v[16], w[4], cx[4], cy[4];
  jac(cfg, m, v, w);
  float phi = 0;
  for (int k = 0; k < 4; ++k) { float lg = logf(w[k]); phi += lg*lg; float c = 2*lg/w[k]; cy[k] = c; cx[k] = -c*w[k]; }
  float wm[16]; lowerT_mul<float, 4>(ic.a, v, wm);
  wdwt<float, 4>(wm, cx, gx); wdwt<float, 4>(wm, cy, gy);
  return phi;
}

int main(int argc, char** argv) {
  const int P = argc > 1 ? atoi(argv[1]) : 100000;
  std::vector<Cfg> cfgs = {{1, 0, 10}, {1, 0, 10, 1}, {4, 0, 10, 1}, {4, 0, 10, 1, 2.4e-7f}, {4, 0, 10, 1, 2.4e-7f, 3}, {8, 0, 10, 1, 2.4e-7f, 3}};
  for (double ir : {0.1, 1.0, 3.0}) {
    std::vector<double> X(16 * (size_t)P), Y(16 * (size_t)P);
    for (int p = 0; p < P; ++p) { rand_spd(ir, *(double(*)[16])&X[16*p]); rand_spd(ir, *(double(*)[16])&Y[16*p]); }
    for (auto& cfg : cfgs) {
      SpdAI<double, 4, false, false> opd{1e-8, 1e8};
      double max_rel_d2 = 0, max_rel_g = 0, sum_rel_g = 0;
      std::vector<double> relg;
      long long hist[16] = {0}, whist[16] = {0};
      int wmax = 0;
      for (int p = 0; p < P; ++p) {
        double xd[16], yd[16]; float xf[16], yf[16];
        for (int k = 0; k < 16; ++k) { xf[k] = (float)X[16*p+k]; yf[k] = (float)Y[16*p+k]; xd[k] = xf[k]; yd[k] = yf[k]; }
        float gxf[16], gyf[16]; double gxd[16], gyd[16];
        float d2f = pair_f32(cfg, xf, yf, gxf, gyf);
        hist[g_last_sweeps]++; wmax = std::max(wmax, g_last_sweeps);
        if (p % 32 == 31) { whist[wmax]++; wmax = 0; }
        double d2d = opd.dist2_grad(xd, yd, gxd, gyd);
        double num = 0, den = 0;
        for (int k = 0; k < 16; ++k) { num = std::max(num, std::fabs(gxf[k]-gxd[k])); num = std::max(num, std::fabs(gyf[k]-gyd[k]));
                                       den = std::max(den, std::fabs(gxd[k])); den = std::max(den, std::fabs(gyd[k])); }
        double rg = num / den; relg.push_back(rg);
        max_rel_g = std::max(max_rel_g, rg); sum_rel_g += rg;
        max_rel_d2 = std::max(max_rel_d2, std::fabs(d2f - d2d) / d2d);
      }
      std::sort(relg.begin(), relg.end());
      printf("ir=%.1f thr x%-5g fixed %d ord %d noise %.1e fc %d | d2 max %.2e | grad max %.2e mean %.2e p99 %.2e | sweeps", ir, cfg.thr_mult, cfg.fixed, cfg.order, cfg.cnoise, cfg.first_check,
             max_rel_d2, max_rel_g, sum_rel_g / P, relg[(size_t)(0.99*P)]);
      for (int k = 0; k < 9; ++k) if (hist[k]) printf(" %d:%.1f%%", k, 100.0*hist[k]/P);
      printf(" | warp-max"); for (int k = 0; k < 9; ++k) if (whist[k]) printf(" %d:%.1f%%", k, 100.0*whist[k]/(P/32));
      printf("\n");
    }
  }
  return 0;
}
