// gcl_ops.cu -- the reference operators one by one (fp64, device arrays), for the drop-in Python operators that
// keep the reference's per-operator signatures.  The fused pipeline (gcl_scan_moments.cu) does not use these.
// Reductions are two-stage and fixed-order (per-block partials, then one block sums them in index order).
#include "gcl_internal.cuh"

namespace gcl {

constexpr int kOpThreads = 256;
constexpr int kMaxBlocks = 512;  // 512 blocks x <= 8 partials = 4096 doubles of workspace

template <int NV>
__device__ __forceinline__ void block_partials(double (&v)[NV], const bool (&is_max)[NV], double* ws_block) {
  __shared__ double s[kOpThreads / 32][NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double x = v[k];
    for (int o = 16; o > 0; o >>= 1) {
      const double y = __shfl_xor_sync(0xffffffffu, x, o);
      x = is_max[k] ? fmax(x, y) : x + y;
    }
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5][k] = x;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double x = s[0][threadIdx.x];
    for (int w = 1; w < kOpThreads / 32; ++w) x = is_max[threadIdx.x] ? fmax(x, s[w][threadIdx.x]) : x + s[w][threadIdx.x];
    ws_block[threadIdx.x] = x;
  }
}

// ---- PointBudgetResample (pkg/backend/operators/point_budget.py:50-109) --------------------------------------------
__global__ void __launch_bounds__(kOpThreads) resample_sums_kernel(const double* __restrict__ w, int64_t n, int64_t stride,
                                                                   int64_t n_sel, double* ws) {
  double v[3] = {0, 0, 0};
  const bool mx[3] = {false, false, false};
  const int64_t gsz = (int64_t)gridDim.x * kOpThreads;
  for (int64_t i = (int64_t)blockIdx.x * kOpThreads + threadIdx.x; i < n; i += gsz) v[0] += w[i];
  for (int64_t j = (int64_t)blockIdx.x * kOpThreads + threadIdx.x; j < n_sel; j += gsz) {
    const double x = w[j * stride];
    v[1] += x;
    v[2] += x * x;
  }
  block_partials<3>(v, mx, ws + blockIdx.x * 3);
}

__global__ void __launch_bounds__(kOpThreads) resample_gather_kernel(
    const double* __restrict__ pts, const double* __restrict__ ts, const double* __restrict__ w,
    const uint8_t* __restrict__ ring, const uint8_t* __restrict__ tag, int64_t n, int64_t stride, int64_t n_sel, int cap,
    double eps, const double* __restrict__ ws, int n_sum_blocks, double* pts_out, double* ts_out, double* w_out,
    uint8_t* ring_out, uint8_t* tag_out, double* scalars) {
  __shared__ double s_tot[3];
  if (threadIdx.x < 3) {
    double x = 0;
    for (int b = 0; b < n_sum_blocks; ++b) x += ws[b * 3 + threadIdx.x];
    s_tot[threadIdx.x] = x;
  }
  __syncthreads();
  const double M = s_tot[0], Msel = s_tot[1];
  const double scale = M / (Msel + eps);
  const int64_t gsz = (int64_t)gridDim.x * kOpThreads;
  for (int64_t j = (int64_t)blockIdx.x * kOpThreads + threadIdx.x; j < cap; j += gsz) {
    if (j < n_sel) {
      const int64_t r = j * stride;
      pts_out[j * 3] = pts[r * 3]; pts_out[j * 3 + 1] = pts[r * 3 + 1]; pts_out[j * 3 + 2] = pts[r * 3 + 2];
      ts_out[j] = ts[r];
      w_out[j] = w[r] * scale;
      ring_out[j] = ring ? ring[r] : 0;
      tag_out[j] = tag ? tag[r] : 0;
    } else {  // zero padding, point_budget.py:74-91
      pts_out[j * 3] = 0; pts_out[j * 3 + 1] = 0; pts_out[j * 3 + 2] = 0;
      ts_out[j] = 0; w_out[j] = 0; ring_out[j] = 0; tag_out[j] = 0;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const double wn2 = scale * scale * s_tot[2] / ((M + eps) * (M + eps));
    scalars[0] = (double)n; scalars[1] = (double)n_sel; scalars[2] = (double)stride; scalars[3] = M; scalars[4] = Msel;
    scalars[5] = scale; scalars[6] = 1.0 / (wn2 + (double)cap * eps); scalars[7] = 0.0;
  }
}

cudaError_t op_point_budget(const double* pts, const double* ts, const double* w, const uint8_t* ring, const uint8_t* tag,
                            int64_t n, int cap, double eps, double* pts_out, double* ts_out, double* w_out,
                            uint8_t* ring_out, uint8_t* tag_out, double* scalars, double* ws, cudaStream_t st, int* launches) {
  const int64_t c = cap > 0 ? cap : 1;
  int64_t stride = (n + c - 1) / c;
  if (stride < 1) stride = 1;
  const int64_t n_sel = (n + stride - 1) / stride;
  int blocks = (int)((n + kOpThreads - 1) / kOpThreads);
  blocks = blocks < 1 ? 1 : (blocks > kMaxBlocks ? kMaxBlocks : blocks);
  resample_sums_kernel<<<blocks, kOpThreads, 0, st>>>(w, n, stride, n_sel, ws);
  int gblocks = (cap + kOpThreads - 1) / kOpThreads;
  gblocks = gblocks < 1 ? 1 : (gblocks > 4 * kMaxBlocks ? 4 * kMaxBlocks : gblocks);
  resample_gather_kernel<<<gblocks, kOpThreads, 0, st>>>(pts, ts, w, ring, tag, n, stride, n_sel, cap, eps, ws, blocks, pts_out,
                                                          ts_out, w_out, ring_out, tag_out, scalars);
  *launches += 2;
  return cudaGetLastError();
}

// ---- DeskewConstantTwist (pkg/backend/operators/deskew_constant_twist.py:31-69) --------------------------------------
struct Twist { double xi[6]; };

__global__ void __launch_bounds__(kOpThreads) deskew_kernel(const double* __restrict__ pts, const double* __restrict__ ts,
                                                            const double* __restrict__ w, int64_t n, double t0, double t1,
                                                            const __grid_constant__ Twist tw, double* pts_out, double* w_out,
                                                            double* ws) {
  const double denom = fmax(t1 - t0, 1e-12);
  const double sigma = fmax(0.1 * denom, 1e-6);
  double v[2] = {0, 0};
  const bool mx[2] = {false, false};
  const int64_t gsz = (int64_t)gridDim.x * kOpThreads;
  for (int64_t i = (int64_t)blockIdx.x * kOpThreads + threadIdx.x; i < n; i += gsz) {
    const double t = ts[i];
    const double p[3] = {pts[i * 3], pts[i * 3 + 1], pts[i * 3 + 2]};
    double p0[3];
    deskew_point64(tw.xi, (t - t0) / denom, p, p0);
    const double wo = w[i] * window_weight64(t, t0, t1, sigma, 1e-12);
    pts_out[i * 3] = p0[0]; pts_out[i * 3 + 1] = p0[1]; pts_out[i * 3 + 2] = p0[2];
    w_out[i] = wo;
    v[0] += wo;
    v[1] += w[i];
  }
  block_partials<2>(v, mx, ws + blockIdx.x * 2);
}

__global__ void deskew_finish_kernel(const double* ws, int nblocks, double eps, double* scalars) {
  if (threadIdx.x == 0) {
    double a = 0, b = 0;
    for (int i = 0; i < nblocks; ++i) { a += ws[i * 2]; b += ws[i * 2 + 1]; }
    scalars[0] = a; scalars[1] = b; scalars[2] = a / (b + eps); scalars[3] = 0;
  }
}

cudaError_t op_deskew(const double* pts, const double* ts, const double* w, int64_t n, double t0, double t1, const double* xi,
                      double eps, double* pts_out, double* w_out, double* scalars, double* ws, cudaStream_t st, int* launches) {
  Twist tw;
  for (int i = 0; i < 6; ++i) tw.xi[i] = xi[i];
  int blocks = (int)((n + kOpThreads - 1) / kOpThreads);
  blocks = blocks < 1 ? 1 : (blocks > kMaxBlocks ? kMaxBlocks : blocks);
  deskew_kernel<<<blocks, kOpThreads, 0, st>>>(pts, ts, w, n, t0, t1, tw, pts_out, w_out, ws);
  deskew_finish_kernel<<<1, 32, 0, st>>>(ws, blocks, eps, scalars);
  *launches += 2;
  return cudaGetLastError();
}

// ---- ray directions (pkg/backend/pipeline.py:591-593) ------------------------------------------------------------------
__global__ void ray_dirs_kernel(const double* __restrict__ pts, int64_t n, double ox, double oy, double oz, double eps,
                                double* out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double r0 = pts[i * 3] - ox, r1 = pts[i * 3 + 1] - oy, r2 = pts[i * 3 + 2] - oz;
  const double dn = sqrt(r0 * r0 + r1 * r1 + r2 * r2) + eps;
  out[i * 3] = r0 / dn; out[i * 3 + 1] = r1 / dn; out[i * 3 + 2] = r2 / dn;
}

cudaError_t op_ray_dirs(const double* pts, int64_t n, const double* o, double eps, double* out, cudaStream_t st, int* launches) {
  if (n > 0) ray_dirs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(pts, n, o[0], o[1], o[2], eps, out);
  *launches += (n > 0);
  return cudaGetLastError();
}

// ---- BinSoftAssign, materialised (archive/legacy_operators/binning.py:56-76) ---------------------------------------------
// one warp per point, lanes over bins: the (N,B) write is coalesced; row max / sum by shuffles (jax.nn.softmax order
// differs only in the association of the sum)
__global__ void __launch_bounds__(kOpThreads) soft_assign_kernel(const double* __restrict__ dirs, int64_t n,
                                                                 const double* __restrict__ bins, int B, double tau, double eps,
                                                                 double* __restrict__ resp, double* ws) {
  extern __shared__ double s_b[];  // [3B]
  for (int i = threadIdx.x; i < 3 * B; i += kOpThreads) s_b[i] = bins[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t wstride = (int64_t)gridDim.x * (kOpThreads / 32);
  double v[2] = {0, 0};  // entropy sum, max resp
  const bool mx[2] = {false, true};
  for (int64_t i = (int64_t)blockIdx.x * (kOpThreads / 32) + warp; i < n; i += wstride) {
    const double d0 = dirs[i * 3], d1 = dirs[i * 3 + 1], d2 = dirs[i * 3 + 2];
    double m = -1e300;
    for (int b = lane; b < B; b += 32) m = fmax(m, (d0 * s_b[b * 3] + d1 * s_b[b * 3 + 1] + d2 * s_b[b * 3 + 2]) / tau);
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    double z = 0;
    for (int b = lane; b < B; b += 32) z += exp((d0 * s_b[b * 3] + d1 * s_b[b * 3 + 1] + d2 * s_b[b * 3 + 2]) / tau - m);
    for (int o = 16; o > 0; o >>= 1) z += __shfl_xor_sync(0xffffffffu, z, o);
    double h = 0, rm = 0;
    for (int b = lane; b < B; b += 32) {
      const double r = exp((d0 * s_b[b * 3] + d1 * s_b[b * 3 + 1] + d2 * s_b[b * 3 + 2]) / tau - m) / z;
      resp[i * B + b] = r;
      h -= r * log(r + eps);
      rm = fmax(rm, r);
    }
    v[0] += h;  // lane-local; reduced across the block below
    v[1] = fmax(v[1], rm);
  }
  block_partials<2>(v, mx, ws + blockIdx.x * 2);
}

__global__ void soft_assign_finish_kernel(const double* ws, int nblocks, int64_t n, double eps, double* scalars) {
  if (threadIdx.x == 0) {
    double a = 0, b = (n > 0) ? 0.0 : -INFINITY;
    for (int i = 0; i < nblocks; ++i) { a += ws[i * 2]; b = fmax(b, ws[i * 2 + 1]); }
    const double avg = a / ((double)n + eps);
    scalars[0] = avg; scalars[1] = b; scalars[2] = exp(avg); scalars[3] = 0;
  }
}

cudaError_t op_soft_assign(const double* dirs, int64_t n, const double* bins, int B, double tau, double eps, double* resp,
                           double* scalars, double* ws, cudaStream_t st, int* launches) {
  int64_t want = (n + (kOpThreads / 32) - 1) / (kOpThreads / 32);
  int blocks = (int)(want < 1 ? 1 : (want > kMaxBlocks ? kMaxBlocks : want));
  soft_assign_kernel<<<blocks, kOpThreads, 3 * B * sizeof(double), st>>>(dirs, n, bins, B, tau, eps, resp, ws);
  soft_assign_finish_kernel<<<1, 32, 0, st>>>(ws, blocks, n, eps, scalars);
  *launches += 2;
  return cudaGetLastError();
}

// ---- KappaFromResultant (pkg/backend/operators/kappa.py:130-169) ---------------------------------------------------------
__global__ void kappa_kernel(const double* __restrict__ R, int64_t n, double eps_r, double d, double r0, double tau, double* out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = kappa_from_resultant(R[i], eps_r, d, r0, tau);
}
cudaError_t op_kappa(const double* R, int64_t n, double eps_r, int d, double r0, double tau, double* out, cudaStream_t st, int* launches) {
  if (n > 0) kappa_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(R, n, eps_r, (double)d, r0, tau, out);
  *launches += (n > 0);
  return cudaGetLastError();
}

// ---- Fibonacci atlas (archive/bin_atlas.py:40-61) -------------------------------------------------------------------------
__global__ void atlas_kernel(int B, double eps, double* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  const double idx = (double)i + 0.5;
  const double phi = acos(1.0 - 2.0 * idx / (double)B);
  const double theta = (3.141592653589793 * (1.0 + sqrt(5.0))) * idx;
  double x = sin(phi) * cos(theta), y = sin(phi) * sin(theta), z = cos(phi);
  const double dn = sqrt(x * x + y * y + z * z) + eps;
  out[i * 3] = x / dn; out[i * 3 + 1] = y / dn; out[i * 3 + 2] = z / dn;
}
cudaError_t op_atlas(int B, double eps, double* out, cudaStream_t st, int* launches) {
  atlas_kernel<<<(B + 127) / 128, 128, 0, st>>>(B, eps, out);
  *launches += 1;
  return cudaGetLastError();
}

// ---- pack to the 20 B/point layout ------------------------------------------------------------------------------------------
__global__ void pack_kernel(const double* __restrict__ pts, const double* __restrict__ ts, const double* __restrict__ w,
                            int64_t n, double t0, float4* xyzt, float* wf) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  xyzt[i] = make_float4((float)pts[i * 3], (float)pts[i * 3 + 1], (float)pts[i * 3 + 2], (float)(ts[i] - t0));
  wf[i] = (float)w[i];
}
cudaError_t op_pack(const double* pts, const double* ts, const double* w, int64_t n, double t0, void* xyzt, float* wf,
                    cudaStream_t st, int* launches) {
  if (n > 0) pack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(pts, ts, w, n, t0, static_cast<float4*>(xyzt), wf);
  *launches += (n > 0);
  return cudaGetLastError();
}

}  // namespace gcl
