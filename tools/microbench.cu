// DEV TOOL (not shipped): issue-rate microbenchmarks for the instruction classes the pair kernel is made of.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/microbench tools/microbench.cu && /tmp/microbench
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long pk(float lo, float hi){ unsigned long long r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk(unsigned long long v, float& lo, float& hi){ asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c){
  unsigned long long d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

constexpr int ITERS = 4096, CH = 8;
template <int MODE>
__global__ void __launch_bounds__(256) kern(float* out, float a, float b) {
  float x[CH]; unsigned long long y[CH];
  for (int i = 0; i < CH; ++i) { x[i] = threadIdx.x * 1e-3f + i; y[i] = pk(x[i], x[i] + 1.f); }
  unsigned long long bb = pk(b, b * 1.01f);
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      if (MODE == 0) x[i] = fmaf(x[i], a, b);                                  // FFMA
      if (MODE == 1) y[i] = fma2(y[i], pk(a, a), bb);                          // FFMA2 (broadcast scalar)
      if (MODE == 2) { float r; asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x[i])); x[i] = r; }  // MUFU
      if (MODE == 3) { x[i] = fmaf(x[i], a, b); y[i] = fma2(y[i], pk(a, a), bb); }  // mixed 1:1
      if (MODE == 4) { x[i] = fmaf(x[i], a, b); x[i] = __int_as_float(__float_as_int(x[i]) ^ 0x1234); }  // FFMA + LOP3
    }
  }
  float s = 0; for (int i = 0; i < CH; ++i) { float lo, hi; upk(y[i], lo, hi); s += x[i] + lo + hi; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name, int per_iter) {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  kern<MODE><<<148 * 8, 256>>>(out, 1.0001f, 0.5f);
  cudaEventRecord(e0);
  kern<MODE><<<148 * 8, 256>>>(out, 1.0001f, 0.5f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double winst = 148.0 * 8 * 8 * ITERS * CH * per_iter;  // warp instructions
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double cyc = ms * 1e-3 * clk * 1e3;
  printf("%-28s %.3f ms  %.2f warp-inst/clk/SM (%.2f per SMSP)\n", name, ms, winst / cyc / 148, winst / cyc / 148 / 4);
  cudaFree(out);
}
int main() {
  run<0>("FFMA", 1); run<1>("FFMA2 (packed)", 1); run<2>("MUFU.RSQ", 1); run<3>("FFMA + FFMA2 1:1", 2); run<4>("FFMA + LOP3 1:1", 2);
  return 0;
}
