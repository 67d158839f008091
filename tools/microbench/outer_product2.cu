// Phase-B inner loop, FFMA2 paired over bins (e pair, f broadcast), sweeping bins/thread G, threads/CTA NT and an
// explicit register prefetch of the next point row.  B = 256 bins per CTA; reports pairs/clk/SM.
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
constexpr int NPTS = 512, STRIDE = 28, B = 256;
struct Row { float4 d, f0, f1, f2, f3, f4; };
__device__ __forceinline__ Row load_row(const float* tile, int pt) {
  const float4* r4 = reinterpret_cast<const float4*>(tile + pt * STRIDE);
  Row r; r.d = r4[0]; r.f0 = r4[1]; r.f1 = r4[2]; r.f2 = r4[3]; r.f3 = r4[4]; r.f4 = r4[5];
  return r;
}
template <int G, int NT, bool PREFETCH>
__global__ void __launch_bounds__(NT, 1) k2(float* out, int reps) {
  extern __shared__ __align__(16) float tile[];
  for (int i = threadIdx.x; i < NPTS * STRIDE; i += NT) tile[i] = 1e-3f * (i % 97);
  __syncthreads();
  constexpr int L = B / G, PS = NT / L;
  float2 bx[G / 2], by[G / 2], bz[G / 2];
  for (int g = 0; g < G / 2; ++g) { bx[g] = make_float2(0.1f * g + threadIdx.x * 1e-3f, 0.05f); by[g] = make_float2(0.2f, 0.1f); bz[g] = make_float2(0.3f * g, 0.2f); }
  float2 acc2[G / 2][19];
  for (int g = 0; g < G / 2; ++g) for (int j = 0; j < 19; ++j) acc2[g][j] = make_float2(0.f, 0.f);
  const int slice = threadIdx.x / L;
  for (int r = 0; r < reps; ++r) {
    Row cur = load_row(tile, slice);
    for (int pt = slice; pt < NPTS; pt += PS) {
      Row nxt;
      if (PREFETCH) nxt = load_row(tile, (pt + PS < NPTS) ? pt + PS : pt); else cur = load_row(tile, pt);
      const float f[20] = {cur.f0.x, cur.f0.y, cur.f0.z, cur.f0.w, cur.f1.x, cur.f1.y, cur.f1.z, cur.f1.w, cur.f2.x, cur.f2.y, cur.f2.z, cur.f2.w,
                           cur.f3.x, cur.f3.y, cur.f3.z, cur.f3.w, cur.f4.x, cur.f4.y, cur.f4.z, cur.f4.w};
#pragma unroll
      for (int gp = 0; gp < G / 2; ++gp) {
        float2 s = ffma2(make_float2(cur.d.z, cur.d.z), bz[gp], make_float2(cur.d.w, cur.d.w));
        s = ffma2(make_float2(cur.d.y, cur.d.y), by[gp], s);
        s = ffma2(make_float2(cur.d.x, cur.d.x), bx[gp], s);
        const float2 ee = make_float2(ex2(s.x), ex2(s.y));
#pragma unroll
        for (int j = 0; j < 19; ++j) acc2[gp][j] = ffma2(ee, make_float2(f[j], f[j]), acc2[gp][j]);
      }
      if (PREFETCH) cur = nxt;
    }
  }
  float s = 0;
  for (int g = 0; g < G / 2; ++g) for (int j = 0; j < 19; ++j) s += acc2[g][j].x + acc2[g][j].y;
  out[blockIdx.x * NT + threadIdx.x] = s;
}
template <int G, int NT, bool PF> void run(const char* name) {
  float* out; cudaMalloc(&out, 148 * NT * sizeof(float));
  auto kern = k2<G, NT, PF>;
  const size_t smem = NPTS * STRIDE * 4;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int reps = 100;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  kern<<<148, NT, smem>>>(out, 2);
  cudaEventRecord(e0);
  kern<<<148, NT, smem>>>(out, reps);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double pairs = (double)reps * NPTS * B * 148;
  cudaFuncAttributes a; cudaFuncGetAttributes(&a, kern);
  printf("%-34s regs %3d  %.3f ms  %7.1f Gpairs/s  %.2f pairs/clk/SM\n", name, a.numRegs, ms, pairs / ms * 1e-6, pairs / (ms * 1e-3) / 148 / 1.9e9);
  cudaFree(out);
}
int main() {
  run<4, 256, false>("G=4 NT=256");
  run<4, 256, true>("G=4 NT=256 prefetch");
  run<2, 512, false>("G=2 NT=512");
  run<2, 512, true>("G=2 NT=512 prefetch");
  run<8, 256, false>("G=8 NT=256");
  run<8, 256, true>("G=8 NT=256 prefetch");
  run<8, 128, false>("G=8 NT=128");
  run<8, 128, true>("G=8 NT=128 prefetch");
  run<4, 512, false>("G=4 NT=512 (8 slices)");
  run<4, 128, true>("G=4 NT=128 prefetch");
  return 0;
}
