// Microbenchmark of the phase-B inner loop variants: acc[g][k] += e[g] * f[k] with f streamed from shared memory.
//   V0 scalar FFMA                       (19 FFMA per bin)
//   V1 FFMA2 paired over MOMENTS         (e broadcast .F32, f pair, acc pair)
//   V2 FFMA2 paired over BINS            (e pair, f broadcast .F32, acc pair)
// G = 4 bins per thread, 256 threads per CTA, 1 CTA per SM; reports G-bin*point pairs per clock per SM.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(reinterpret_cast<unsigned long long&>(d))
      : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)),
        "l"(reinterpret_cast<unsigned long long&>(c)));
  return d;
}
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
constexpr int G = 4, NPTS = 256, STRIDE = 28;
template <int V, bool WITH_E>
__global__ void __launch_bounds__(256, 1) k(float* out, int reps) {
  __shared__ __align__(16) float tile[NPTS * STRIDE];
  for (int i = threadIdx.x; i < NPTS * STRIDE; i += 256) tile[i] = 1e-3f * (i % 97);
  __syncthreads();
  float bx[G], by[G], bz[G];
  for (int g = 0; g < G; ++g) { bx[g] = 0.1f * g + threadIdx.x * 1e-3f; by[g] = 0.2f; bz[g] = 0.3f * g; }
  float acc[G][20];
  float2 acc2[G][10];
  for (int g = 0; g < G; ++g) for (int j = 0; j < 10; ++j) { acc2[g][j] = make_float2(0.f, 0.f); acc[g][2 * j] = acc[g][2 * j + 1] = 0.f; }
  const int slice = threadIdx.x >> 6;  // 4 slices of 64 lanes
  for (int r = 0; r < reps; ++r) {
    for (int pt = slice; pt < NPTS; pt += 4) {
      const float4* r4 = reinterpret_cast<const float4*>(tile + pt * STRIDE);
      const float4 d = r4[0], f0 = r4[1], f1 = r4[2], f2 = r4[3], f3 = r4[4], f4 = r4[5];
      float e[G];
#pragma unroll
      for (int g = 0; g < G; ++g) e[g] = WITH_E ? ex2(fmaf(d.x, bx[g], fmaf(d.y, by[g], fmaf(d.z, bz[g], d.w)))) : bx[g] + d.x;
      const float f[20] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w, f2.x, f2.y, f2.z, f2.w, f3.x, f3.y, f3.z, f3.w, f4.x, f4.y, f4.z, f4.w};
      if (V == 0) {
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
          for (int j = 0; j < 19; ++j) acc[g][j] = fmaf(e[g], f[j], acc[g][j]);
      } else if (V == 1) {
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const float2 ee = make_float2(e[g], e[g]);
#pragma unroll
          for (int j = 0; j < 9; ++j) acc2[g][j] = ffma2(ee, make_float2(f[2 * j], f[2 * j + 1]), acc2[g][j]);
          acc[g][18] = fmaf(e[g], f[18], acc[g][18]);
        }
      } else {
#pragma unroll
        for (int gp = 0; gp < G / 2; ++gp) {
          const float2 ee = make_float2(e[2 * gp], e[2 * gp + 1]);
#pragma unroll
          for (int j = 0; j < 19; ++j) acc2[gp][j % 10 + 0 * j] = ffma2(ee, make_float2(f[j], f[j]), acc2[gp][j % 10]);  // placeholder index
        }
      }
    }
  }
  float s = 0;
  for (int g = 0; g < G; ++g) for (int j = 0; j < 10; ++j) s += acc2[g][j].x + acc2[g][j].y + acc[g][2 * j] + acc[g][2 * j + 1];
  out[blockIdx.x * 256 + threadIdx.x] = s;
}
// V2 needs 19 accumulator pairs per bin pair: separate kernel to keep register arrays exact
template <bool WITH_E>
__global__ void __launch_bounds__(256, 1) k2(float* out, int reps) {
  __shared__ __align__(16) float tile[NPTS * STRIDE];
  for (int i = threadIdx.x; i < NPTS * STRIDE; i += 256) tile[i] = 1e-3f * (i % 97);
  __syncthreads();
  float2 bx[G / 2], by[G / 2], bz[G / 2];
  for (int g = 0; g < G / 2; ++g) { bx[g] = make_float2(0.1f * g + threadIdx.x * 1e-3f, 0.05f); by[g] = make_float2(0.2f, 0.1f); bz[g] = make_float2(0.3f * g, 0.2f); }
  float2 acc2[G / 2][19];
  for (int g = 0; g < G / 2; ++g) for (int j = 0; j < 19; ++j) acc2[g][j] = make_float2(0.f, 0.f);
  const int slice = threadIdx.x >> 6;
  for (int r = 0; r < reps; ++r) {
    for (int pt = slice; pt < NPTS; pt += 4) {
      const float4* r4 = reinterpret_cast<const float4*>(tile + pt * STRIDE);
      const float4 d = r4[0], f0 = r4[1], f1 = r4[2], f2 = r4[3], f3 = r4[4], f4 = r4[5];
      const float f[20] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w, f2.x, f2.y, f2.z, f2.w, f3.x, f3.y, f3.z, f3.w, f4.x, f4.y, f4.z, f4.w};
#pragma unroll
      for (int gp = 0; gp < G / 2; ++gp) {
        float2 s = ffma2(make_float2(d.z, d.z), bz[gp], make_float2(d.w, d.w));
        s = ffma2(make_float2(d.y, d.y), by[gp], s);
        s = ffma2(make_float2(d.x, d.x), bx[gp], s);
        const float2 ee = WITH_E ? make_float2(ex2(s.x), ex2(s.y)) : s;
#pragma unroll
        for (int j = 0; j < 19; ++j) acc2[gp][j] = ffma2(ee, make_float2(f[j], f[j]), acc2[gp][j]);
      }
    }
  }
  float s = 0;
  for (int g = 0; g < G / 2; ++g) for (int j = 0; j < 19; ++j) s += acc2[g][j].x + acc2[g][j].y;
  out[blockIdx.x * 256 + threadIdx.x] = s;
}
template <typename F> void run(const char* name, F kern) {
  float* out; cudaMalloc(&out, 148 * 256 * sizeof(float));
  const int reps = 200;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  kern<<<148, 256>>>(out, 2);
  cudaEventRecord(e0);
  kern<<<148, 256>>>(out, reps);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  // per CTA: reps * NPTS points * (256 threads / 4 slices... every point visits 64 lanes * G bins = 256 bins)
  const double pairs = (double)reps * NPTS * 256.0 * 148;
  printf("%-44s %.3f ms  %.1f Gpairs/s  %.2f pairs/clk/SM (floor 19 fma: 6.74, 22 fma: 5.82)\n", name, ms, pairs / ms * 1e-6,
         pairs / (ms * 1e-3) / 148 / 1.9e9);
  cudaFree(out);
}
int main() {
  run("V0 scalar FFMA, no exp", k<0, false>);
  run("V0 scalar FFMA, with sim+exp", k<0, true>);
  run("V1 FFMA2 over moments, no exp", k<1, false>);
  run("V1 FFMA2 over moments, with sim+exp", k<1, true>);
  run("V2 FFMA2 over bins, no exp", k2<false>);
  run("V2 FFMA2 over bins, with sim+exp", k2<true>);
  return 0;
}
