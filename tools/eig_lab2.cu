// DEV TOOL (host only): warp-cooperative Jacobi policy simulation for 4x4 pair matrices.
// Simulates warps of 32 pairs: round-robin order, 3 unconditional sweeps, then sweeps in which a rotation is executed
// only if ANY lane of the warp needs it (apq^2 > thr^2 app aqq); reports executed rotations per warp and accuracy.
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
  double s[16]; double nrm = 0;
  for (int i = 0; i < 4; ++i) for (int j = i; j < 4; ++j) { double u = nd(rng); s[i*4+j] = s[j*4+i] = u; nrm += u*u; }
  nrm = std::sqrt(nrm);
  for (int k = 0; k < 16; ++k) s[k] *= ir / nrm;
  double v[16], w[4];
  jacobi_eigh<double, 4, true>(s, v, w);
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { double a = 0; for (int k = 0; k < 4; ++k) a += v[i*4+k]*std::exp(w[k])*v[j*4+k]; x[i*4+j] = a; }
}
struct Lane { float a[16], v[16]; };
static void rotate(Lane& L, int p, int q) {
  float* a = L.a; float* v = L.v; const int N = 4;
  float apq = a[p*N+q]; float t, c, s;
  Num<float>::rotation(a[q*N+q]-a[p*N+p], apq, t, c, s);
  a[p*N+p] -= t*apq; a[q*N+q] += t*apq; a[p*N+q] = a[q*N+p] = 0;
  for (int r = 0; r < N; ++r) if (r != p && r != q) { float arp = a[r*N+p], arq = a[r*N+q]; float nrp = c*arp - s*arq, nrq = s*arp + c*arq; a[r*N+p]=a[p*N+r]=nrp; a[r*N+q]=a[q*N+r]=nrq; }
  for (int r = 0; r < N; ++r) { float vp = v[r*N+p], vq = v[r*N+q]; v[r*N+p] = c*vp - s*vq; v[r*N+q] = s*vp + c*vq; }
}
int main(int argc, char** argv) {
  const int W = argc > 1 ? atoi(argv[1]) : 3000;
  static const int ord[6][2] = {{0,1},{2,3},{0,2},{1,3},{0,3},{1,2}};
  for (double ir : {0.1, 1.0, 3.0}) for (float thrm : {0.25f, 1.0f, 4.0f}) for (int uncond : {2, 3}) {
    double rot_sum = 0, max_rel = 0, sum_rel = 0; long cnt = 0; long hist[40] = {0};
    for (int w = 0; w < W; ++w) {
      Lane L[32]; double gxd[32][16], gyd[32][16]; InvChol<float,4,false> ic[32]; float yf[32][16];
      for (int l = 0; l < 32; ++l) {
        double xd[16], yd[16]; rand_spd(ir, xd); rand_spd(ir, yd);
        float xf[16]; for (int k = 0; k < 16; ++k) { xf[k] = (float)xd[k]; yf[l][k] = (float)yd[k]; xd[k] = xf[k]; yd[k] = yf[l][k]; }
        SpdAI<double,4,false,false> opd{1e-8,1e8}; opd.dist2_grad(xd, yd, gxd[l], gyd[l]);
        ic[l].run(xf); congr_lower<float,4>(ic[l].a, yf[l], L[l].a);
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) L[l].v[i*4+j] = (j >= i) ? ic[l].a[j*4+i] : 0.f;
      }
      int rots = 0;
      for (int sweep = 0; sweep < 10; ++sweep) {
        int done_rots = 0;
        for (int r = 0; r < 6; ++r) {
          int p = ord[r][0], q = ord[r][1];
          bool any = sweep < uncond;
          if (!any) for (int l = 0; l < 32; ++l) { float apq = L[l].a[p*4+q]; float th = FLT_EPSILON*thrm; if (apq*apq > th*th*fabsf(L[l].a[p*4+p]*L[l].a[q*4+q])) any = true; }
          if (any) { for (int l = 0; l < 32; ++l) rotate(L[l], p, q); ++done_rots; }
        }
        rots += done_rots;
        if (done_rots == 0) break;
      }
      rot_sum += rots; hist[rots]++;
      for (int l = 0; l < 32; ++l) {
        float cx[4], cy[4], gx[16], gy[16];
        for (int k = 0; k < 4; ++k) { float wv = L[l].a[k*5]; float lg = logf(wv); float c = 2*lg/wv; cy[k] = c; cx[k] = -c*wv; }
        wdwt<float,4>(L[l].v, cx, gx); wdwt<float,4>(L[l].v, cy, gy);
        double num = 0, den = 0;
        for (int k = 0; k < 16; ++k) { num = std::max(num, std::fabs(gx[k]-gxd[l][k])); num = std::max(num, std::fabs(gy[k]-gyd[l][k])); den = std::max(den, std::fabs(gxd[l][k])); den = std::max(den, std::fabs(gyd[l][k])); }
        max_rel = std::max(max_rel, num/den); sum_rel += num/den; ++cnt;
      }
    }
    printf("ir=%.1f thr x%-4g uncond %d | rotations/warp mean %.2f | grad rel err max %.2e mean %.2e | hist", ir, thrm, uncond, rot_sum/W, max_rel, sum_rel/cnt);
    for (int k = 0; k < 40; ++k) if (hist[k] > W/100) printf(" %d:%.0f%%", k, 100.0*hist[k]/W);
    printf("\n");
  }
}
