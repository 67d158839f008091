// Microbenchmark: FFMA vs FFMA2 (fma.rn.f32x2) vs DFMA vs MUFU.EX2 issue throughput on one B200 SM set.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma_rate ffma_rate.cu && ./ffma_rate
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(reinterpret_cast<unsigned long long&>(d))
               : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)),
                 "l"(reinterpret_cast<unsigned long long&>(c)));
  return d;
}

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float x) {
  constexpr int NACC = 16;
  float2 a[NACC];
  double da[NACC];
  for (int i = 0; i < NACC; ++i) { a[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f); da[i] = i + threadIdx.x; }
  const float2 b = make_float2(x, x * 0.999f), c = make_float2(1e-6f, 2e-6f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      if (MODE == 0) { a[i].x = fmaf(a[i].x, b.x, c.x); a[i].y = fmaf(a[i].y, b.y, c.y); }      // 2 FFMA
      if (MODE == 1) a[i] = ffma2(a[i], b, c);                                                     // 1 FFMA2
      if (MODE == 2) da[i] = fma(da[i], (double)b.x, (double)c.x);                                 // 1 DFMA
      if (MODE == 3) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i].x)); }               // 1 MUFU
      if (MODE == 4) { a[i] = ffma2(a[i], b, c); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i].y)); }
    }
  }
  float s = 0; double ds = 0;
  for (int i = 0; i < NACC; ++i) { s += a[i].x + a[i].y; ds += da[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)ds;
}

template <int MODE> void run(const char* name, double ops_per_iter_per_thread, int ctas_per_sm) {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
  const int iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148 * ctas_per_sm, 256>>>(out, 100, 0.999f);
  cudaEventRecord(e0);
  k<MODE><<<148 * ctas_per_sm, 256>>>(out, iters, 0.999f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double total = ops_per_iter_per_thread * iters * 148.0 * ctas_per_sm * 256;
  printf("%-28s ctas/SM %d: %.3f ms  %.2f Tops/s  -> %.1f ops/clk/SM @1.9GHz\n", name, ctas_per_sm, ms, total / ms * 1e-9,
         total / (ms * 1e-3) / 148 / 1.9e9);
  cudaFree(out);
}

int main() {
  for (int c : {1, 2, 4}) {
    run<0>("FFMA (scalar, 2 per slot)", 32, c);
    run<1>("FFMA2 (counted as 2 fma)", 32, c);
    run<2>("DFMA", 16, c);
    run<3>("MUFU.EX2", 16, c);
    run<4>("FFMA2 + MUFU interleaved", 48, c);
  }
  return 0;
}
