// gcl_finalize.cu -- K2: fixed-order reduction of the per-CTA partial records and the per-bin non-linear part.
//
// One CTA per job.  Step 1 sums `n_parts` records element-wise in a fixed order (deterministic); the result is the
// buffer that is exchanged between GPUs when points are sharded (all sums are linear in the points, the mass
// rescale is applied afterwards).  Step 2 (finalize) reproduces, per bin, archive/legacy_operators/binning.py:175-209
// and the certificate scalars of the resample / deskew / soft-assign operators, including the contribution of the
// zero-weight padded rows to the soft-assign entropy.
#include "gcl_internal.cuh"
#include "gcl_finalize_body.cuh"

namespace gcl {


__global__ void __launch_bounds__(kFinThreads) reduce_finalize_kernel(const __grid_constant__ FinalizeParams P) {
  __shared__ FinalizeShared sh;
  reduce_finalize_job(P, blockIdx.x, sh);
}

// Wide records with many partials (one big scan: 148 CTAs x 19 x 256 moments) would leave the single finalising CTA summing
// ~1400 values per thread: a grid of CTAs does the fixed-order sum first (one element per thread, eight loads in flight, the
// additions still in partial order p = 0, 1, 2, ...), and the finalising CTA then sees one "partial".
__global__ void __launch_bounds__(256) reduce_partials_kernel(const __grid_constant__ FinalizeParams P) {
  const int job = blockIdx.y, rec = P.nmom * P.n_bins + GCL_NCERT;
  const int e = blockIdx.x * 256 + threadIdx.x;
  if (e >= rec) return;
  const double* src = P.parts + (size_t)job * P.job_stride + e;
  const bool is_max = (e == P.nmom * P.n_bins + GCL_CS_MAX_RESP);
  double s = src[0];
  for (int p0 = 1; p0 < P.n_parts; p0 += 8) {
    double v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (p0 + j < P.n_parts) ? src[(size_t)(p0 + j) * P.part_stride] : (is_max ? s : 0.0);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (p0 + j < P.n_parts) s = is_max ? fmax(s, v[j]) : s + v[j];
  }
  P.reduced[(size_t)job * rec + e] = s;
}

cudaError_t launch_reduce_finalize(const FinalizeParams& P, int n_jobs, cudaStream_t stream, int* launches) {
  const int rec = P.nmom * P.n_bins + GCL_NCERT;
  *launches = 1;
  if (P.n_parts >= 16 && (int64_t)rec * P.n_parts >= 256 * 1024) {   // (a 48-bin scan, 79 x 920 values, is faster in one launch)
    // parallel pre-reduction into the reduced record (all there is to do without finalisation), which the finalising CTA
    // then reads as its one partial
    reduce_partials_kernel<<<dim3((rec + 255) / 256, n_jobs), 256, 0, stream>>>(P);
    if (!P.finalize) return cudaGetLastError();
    FinalizeParams F = P;
    F.parts = P.reduced; F.n_parts = 1; F.part_stride = rec; F.job_stride = rec;
    reduce_finalize_kernel<<<n_jobs, kFinThreads, 0, stream>>>(F);
    *launches = 2;
    return cudaGetLastError();
  }
  reduce_finalize_kernel<<<n_jobs, kFinThreads, 0, stream>>>(P);
  return cudaGetLastError();
}

}  // namespace gcl
