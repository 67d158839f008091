// gcl_scan_moments.cu -- moment sums from MATERIALISED responsibilities: the stand-alone ScanBinMomentMatch operator
// (archive/legacy_operators/binning.py:139-209, :212-324), including per-point covariances (28 moments per bin:
// 1, d, ddT(6), p, ppT(6), C(9)).  The fused hot path (which never materialises (N,B)) is gcl_scan_fused.cu.
//
// Per CTA (256 threads, a contiguous slice of the rows), tile = 256 points:
//   phase A  thread per point: ray direction, row w_i * [1, d, ddT, p, ppT, C] to shared memory;
//   phase B  thread = (bin lane, point slice) owns G bins x 28 accumulators in registers, reads the row with warp-broadcast
//            loads and its responsibilities r_ib from global memory (row-major (N,B): consecutive lanes, consecutive bins);
//   flush    fixed-order cross-slice sum into this CTA's fp64 record (no atomics => bit-reproducible).
#include "gcl_internal.cuh"

namespace gcl {

constexpr int kMmThreads = 256;
constexpr int kMmMom = 28;
constexpr int kMmRow = 29;   // doubles per tile row; odd => conflict-free 64-bit stores
constexpr int kMmKC = 5;     // moments per flush pass

template <int G>
__global__ void __launch_bounds__(kMmThreads, 1) moments_from_resp_kernel(const __grid_constant__ ScanParams P) {
  extern __shared__ __align__(16) double s_mm[];
  const int B = P.n_bins, L = P.lanes, Bp = G * L, n_chunks = P.n_chunks, PS = kMmThreads / L;
  double* s_tile = s_mm;                                  // [256][kMmRow]
  double* s_stage = s_mm + (size_t)kMmThreads * kMmRow;   // [PS][kMmKC][Bp]
  __shared__ double s_red[kMmThreads / 32][GCL_NCERT];
  const int tid = threadIdx.x, binlane = tid & (L - 1), slice = tid / L;
  const gcl_job& J = P.job;
  const int64_t rows = J.n_raw;
  const int64_t rpc = (rows + gridDim.x - 1) / gridDim.x;
  const int64_t lo = ((int64_t)blockIdx.x * rpc < rows) ? (int64_t)blockIdx.x * rpc : rows;
  const int64_t hi = (lo + rpc < rows) ? lo + rpc : rows;
  double* part = P.partials + (size_t)blockIdx.x * P.rec_doubles;
  for (int i = tid; i < kMmMom * B + GCL_NCERT; i += kMmThreads) part[i] = 0.0;

  double a[G][kMmMom];
  auto zero = [&]() {
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
      for (int k = 0; k < kMmMom; ++k) a[g][k] = 0.0;
  };
  zero();
  double c_w = 0, c_w2 = 0, c_rows = 0;

  auto flush = [&](int chunk) {
#pragma unroll
    for (int k0 = 0; k0 < kMmMom; k0 += kMmKC) {
#pragma unroll
      for (int kc = 0; kc < kMmKC; ++kc)
        if (k0 + kc < kMmMom) {
#pragma unroll
          for (int g = 0; g < G; ++g) s_stage[((size_t)slice * kMmKC + kc) * Bp + g * L + binlane] = a[g][(k0 + kc < kMmMom) ? k0 + kc : 0];
        }
      __syncthreads();
      for (int e = tid; e < kMmKC * Bp; e += kMmThreads) {
        const int kc = e / Bp, bl = e - kc * Bp, k = k0 + kc, b = chunk * Bp + bl;
        if (k < kMmMom && b < B) {
          double s = 0.0;
          for (int sl = 0; sl < PS; ++sl) s += s_stage[((size_t)sl * kMmKC + kc) * Bp + bl];
          part[(size_t)k * B + b] += s;
        }
      }
      __syncthreads();
    }
    zero();
  };

  for (int64_t tile0 = lo; tile0 < hi; tile0 += kMmThreads) {
    const int count = (int)((hi - tile0 < kMmThreads) ? hi - tile0 : kMmThreads);
    __syncthreads();
    if (tid < count) {
      const int64_t r = tile0 + tid;
      const double* pts = static_cast<const double*>(J.points);
      const double px = pts[r * 3], py = pts[r * 3 + 1], pz = pts[r * 3 + 2];
      const double lam = J.point_lambda ? J.point_lambda[r] : 1.0;
      const double w0 = static_cast<const double*>(J.weights)[r];
      const double w = w0 * lam;   // binning.py:157
      const double r0 = px - P.c.origin[0], r1 = py - P.c.origin[1], r2 = pz - P.c.origin[2];
      const double inv = 1.0 / (sqrt(r0 * r0 + r1 * r1 + r2 * r2) + P.c.eps_mass);   // binning.py:162-164
      const double d0 = r0 * inv, d1 = r1 * inv, d2 = r2 * inv;
      double* f = s_tile + (size_t)tid * kMmRow;
      f[0] = w;
      f[1] = w * d0; f[2] = w * d1; f[3] = w * d2;
      f[4] = f[1] * d0; f[5] = f[1] * d1; f[6] = f[1] * d2; f[7] = f[2] * d1; f[8] = f[2] * d2; f[9] = f[3] * d2;
      f[10] = w * px; f[11] = w * py; f[12] = w * pz;
      f[13] = f[10] * px; f[14] = f[10] * py; f[15] = f[10] * pz; f[16] = f[11] * py; f[17] = f[11] * pz; f[18] = f[12] * pz;
#pragma unroll
      for (int k = 0; k < 9; ++k) f[19 + k] = P.point_cov ? w * P.point_cov[r * 9 + k] : 0.0;   // all nine entries, binning.py:173
      c_w += w0; c_w2 += w0 * w0; c_rows += 1.0;
    }
    __syncthreads();
    for (int chunk = 0; chunk < n_chunks; ++chunk) {
      for (int pt = slice; pt < count; pt += PS) {
        const double* row = s_tile + (size_t)pt * kMmRow;
        double e[G];
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const int b = chunk * Bp + g * L + binlane;
          e[g] = (b < B) ? P.resp[(tile0 + pt) * B + b] : 0.0;
        }
#pragma unroll
        for (int k = 0; k < kMmMom; ++k) {
          const double fk = row[k];
#pragma unroll
          for (int g = 0; g < G; ++g) a[g][k] = fma(e[g], fk, a[g][k]);
        }
      }
      if (n_chunks > 1) { __syncthreads(); flush(chunk); }
    }
  }
  __syncthreads();
  if (n_chunks == 1) flush(0);

  // certificate scalars (the resample / deskew / soft-assign entries stay zero for this operator)
  double cs[GCL_NCERT] = {c_w, c_w, c_w2, c_w, 0.0, 0.0, c_rows, 0.0};
#pragma unroll
  for (int k = 0; k < GCL_NCERT; ++k) {
    double v = cs[k];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((tid & 31) == 0) s_red[tid >> 5][k] = v;
  }
  __syncthreads();
  if (tid < GCL_NCERT) {
    double v = s_red[0][tid];
    for (int w = 1; w < kMmThreads / 32; ++w) v += s_red[w][tid];
    part[(size_t)kMmMom * B + tid] = v;
  }
}

static inline int pow2_ceil_mm(int x) { int p = 1; while (p < x) p <<= 1; return p; }

// bins laid out as n_chunks chunks of G*L (L a power of two in [8,256] lanes, G <= 2 bins per thread: 56 fp64 accumulators)
void plan_bins(int n_bins, int precision, int* G_out, int* L_out, int* chunks_out) {
  (void)precision;
  int bestG = 1, bestL = 256, bestChunks = 1 << 20, bestWaste = 1 << 30;
  for (int G = 2; G >= 1; --G) {
    int L = pow2_ceil_mm((n_bins + G - 1) / G);
    if (L < 8) L = 8;
    if (L > 256) L = 256;
    const int Bp = G * L, chunks = (n_bins + Bp - 1) / Bp, waste = chunks * Bp - n_bins;
    if (chunks < bestChunks || (chunks == bestChunks && waste < bestWaste)) { bestChunks = chunks; bestWaste = waste; bestG = G; bestL = L; }
  }
  *G_out = bestG; *L_out = bestL; *chunks_out = bestChunks;
}

template <int G>
static cudaError_t launch_mm(const ScanParams& P, int ctas, size_t smem, cudaStream_t stream) {
  auto kern = moments_from_resp_kernel<G>;
  static std::atomic<uint64_t> optin{0};
  if (cudaError_t e = optin_dynamic_smem(kern, 200 * 1024, optin)) return e;
  kern<<<ctas, kMmThreads, smem, stream>>>(P);
  return cudaGetLastError();
}

cudaError_t launch_moments_from_resp(const ScanParams& P, int ctas, int G, cudaStream_t stream) {
  const size_t smem = ((size_t)kMmThreads * kMmRow + (size_t)(kMmThreads / P.lanes) * kMmKC * G * P.lanes) * sizeof(double) + 16;
  return G == 1 ? launch_mm<1>(P, ctas, smem, stream) : launch_mm<2>(P, ctas, smem, stream);
}

}  // namespace gcl
