// Microbenchmark: FP64 tensor-core MMA (mma.sync.m8n8k4.f64 -> SASS DMMA) against DFMA on one B200: is there anything to gain by
// moving the fp64 phase B of scan_moments (19 moments x bins x points, a dense contraction) off the DFMA loop?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_rate dmma_rate.cu && ./dmma_rate
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// MODE 0: DMMA only, NACC independent 8x8 accumulators per warp; MODE 1: DFMA only (2 * NACC chains per thread);
// MODE 2: DMMA with its A / B fragments re-read from shared memory every instruction group (3 A + 4 B loads per 12 MMAs: the phase-B shape)
template <int MODE, int NACC>
__global__ void __launch_bounds__(256) k(double* out, int iters, double x) {
  __shared__ double sa[8 * 3 * 32], sb[8 * 4 * 32];
  double c[NACC][2];
  for (int i = 0; i < NACC; ++i) { c[i][0] = i + threadIdx.x; c[i][1] = i * 0.5; }
  for (int i = threadIdx.x; i < 8 * 3 * 32; i += 256) sa[i] = 1e-3 * i;
  for (int i = threadIdx.x; i < 8 * 4 * 32; i += 256) sb[i] = 1.0 - 1e-6 * i;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double a = x + lane * 1e-9, b = 1.0 - x * 1e-3;
  if (MODE == 3) {   // half of the warps (one of the two on every scheduler) on DMMA, the other half on DFMA, each in its own loop (a predicated
                     // mix of both in one loop issues every instruction from every warp): do the two rates add up?  The DFMA warps run 4 x the
                     // iterations, i.e. the same number of multiply-adds as the DMMA warps
    if (warp & 4) {
      for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma(c[i][0], c[i][1], a, b);
    } else {
      for (int it = 0; it < 4 * iters; ++it)
#pragma unroll
        for (int i = 0; i < NACC; ++i) { c[i][0] = fma(c[i][0], b, a); c[i][1] = fma(c[i][1], b, a); }
    }
  }
  for (int it = 0; it < (MODE == 3 ? 0 : iters); ++it) {
    if (MODE == 2) {
      double av[3], bv[4];
      const int slot = (it & 7);
#pragma unroll
      for (int m = 0; m < 3; ++m) av[m] = sa[(slot * 3 + m) * 32 + lane];
#pragma unroll
      for (int n = 0; n < 4; ++n) bv[n] = sb[(slot * 4 + n) * 32 + lane + (warp & 0)];
#pragma unroll
      for (int m = 0; m < 3; ++m)
#pragma unroll
        for (int n = 0; n < 4; ++n) dmma(c[(m * 4 + n) % NACC][0], c[(m * 4 + n) % NACC][1], av[m], bv[n]);
    } else {
#pragma unroll
      for (int i = 0; i < NACC; ++i) {
        if (MODE == 0) dmma(c[i][0], c[i][1], a, b);
        if (MODE == 1) { c[i][0] = fma(c[i][0], b, a); c[i][1] = fma(c[i][1], b, a); }
      }
    }
  }
  double s = 0;
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE, int NACC> void run(const char* name, double fma_per_iter_per_thread, int ctas_per_sm, int threads = 256) {
  double* out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(double));
  const int iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE, NACC><<<148 * ctas_per_sm, threads>>>(out, 100, 0.999);
  cudaEventRecord(e0);
  k<MODE, NACC><<<148 * ctas_per_sm, threads>>>(out, iters, 0.999);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double total = fma_per_iter_per_thread * iters * 148.0 * ctas_per_sm * threads;
  printf("%-44s threads %d ctas/SM %d: %.3f ms  %.2f TFMA/s  -> %.1f FMA/clk/SM @1.965GHz\n", name, threads, ctas_per_sm, ms, total / ms * 1e-9,
         total / (ms * 1e-3) / 148 / 1.965e9);
  cudaFree(out);
}

int main() {
  // one DMMA = 8 x 8 x 4 = 256 FMA per warp = 8 per thread
  for (int c : {1, 2}) {
    run<0, 12>("DMMA m8n8k4, 12 accumulators/warp", 12 * 8, c);
    run<0, 4>("DMMA m8n8k4, 4 accumulators/warp", 4 * 8, c);
    run<2, 12>("DMMA, 3 A + 4 B fragments from smem per 12", 12 * 8, c);
    run<1, 12>("DFMA, 24 chains/thread", 24, c);
  }
  // how much of the FP64 pipe can ONE warp per scheduler fill with independent DFMAs?  (128 threads = one warp per scheduler; 64 = two schedulers used)
  run<1, 12>("DFMA, 24 chains/thread", 24, 1, 128);
  run<0, 12>("DMMA m8n8k4, 12 accumulators/warp", 12 * 8, 1, 128);
  // DMMA and DFMA side by side (4 warps each, one of each per scheduler): FMA per iteration per thread averaged over the two halves
  run<3, 12>("4 warps DMMA + 4 warps DFMA (sum of both)", 12 * 8, 1);
  run<0, 12>("DMMA alone, 4 warps (128 threads)", 12 * 8, 1, 128);
  return 0;
}
