// gcl_ingest.cu -- PointCloud2 ingest on the GPU: the message's raw bytes -> the arrays the hot path reads.
//
// Replaces parse_pointcloud2_vlp16 (pkg/backend/backend_node.py:377-468, datatype table :351-374) and the base-frame transform
// (:1675-1681).  The raw message (point_step bytes per point, typically 22 for the Kimera VLP-16 layout against 40 bytes of
// parsed float64) is what crosses PCIe; each CTA stages a contiguous run of 256 points' bytes in shared memory with 16-byte
// loads and decodes from there.  Byte / integer results (coordinates as exact float32 -> float64 widening, sentinel
// substitution, ring, per-point time) are bit-exact; the range weights go through exp (<= 1 ulp from NumPy's), the 3x3
// transform through fma (last-ulp differences from a BLAS dot product).
#include "gcl_internal.cuh"

namespace gcl {

constexpr int kIngestThreads = 256;

__device__ __forceinline__ double decode_field(const unsigned char* p, int datatype) {
  // sensor_msgs/PointField datatype codes, little endian, no alignment guarantee
  switch (datatype) {
    case 1: return (double)(signed char)p[0];
    case 2: return (double)p[0];
    case 3: return (double)(short)((unsigned)p[0] | ((unsigned)p[1] << 8));
    case 4: return (double)((unsigned)p[0] | ((unsigned)p[1] << 8));
    case 5: return (double)(int)((unsigned)p[0] | ((unsigned)p[1] << 8) | ((unsigned)p[2] << 16) | ((unsigned)p[3] << 24));
    case 6: return (double)((unsigned)p[0] | ((unsigned)p[1] << 8) | ((unsigned)p[2] << 16) | ((unsigned)p[3] << 24));
    case 7: return (double)__uint_as_float((unsigned)p[0] | ((unsigned)p[1] << 8) | ((unsigned)p[2] << 16) | ((unsigned)p[3] << 24));
    default: {
      unsigned long long v = 0;
      for (int k = 7; k >= 0; --k) v = (v << 8) | p[k];
      return __longlong_as_double((long long)v);
    }
  }
}

// Scan time bounds from the per-point times and the header stamp (pkg/backend/backend_node.py:1776-1786):
//   |max - min| < 1e-9 : start = min, end = header stamp;   else  start = min(stamp, min), end = max(stamp, max).
// One CTA: reduces the per-CTA extrema of pass 1, applies the time unit, writes the bounds record
// [start, end, min, max, unit flag] and patches the window of the job records of this scan.
__device__ __forceinline__ void scan_bounds_block(const IngestParams& P, int n_parts) {
  __shared__ double s_lo[kIngestThreads / 32], s_hi[kIngestThreads / 32];
  __shared__ double s_b[2];
  double lo = 1.0e300, hi = -1.0e300;
  if (P.off[4] >= 0) {
    for (int k = threadIdx.x; k < n_parts; k += kIngestThreads) { lo = fmin(lo, __ldcg(P.bounds_parts + 2 * k)); hi = fmax(hi, __ldcg(P.bounds_parts + 2 * k + 1)); }
  }
  for (int o = 16; o > 0; o >>= 1) {
    lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  __syncthreads();   // (the caller's use of these arrays is over)
  if ((threadIdx.x & 31) == 0) { s_lo[threadIdx.x >> 5] = lo; s_hi[threadIdx.x >> 5] = hi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kIngestThreads / 32; ++w) { lo = fmin(lo, s_lo[w]); hi = fmax(hi, s_hi[w]); }
    const int ns = __ldcg(P.flag);
    double tmin, tmax;
    if (P.off[4] >= 0 && P.n_points > 0) { tmin = ns ? lo * 1e-9 : lo; tmax = ns ? hi * 1e-9 : hi; }
    else { tmin = tmax = P.header_stamp; }
    double t0, t1;
    if (fabs(tmax - tmin) < 1e-9) { t0 = tmin; t1 = P.header_stamp; }
    else { t0 = fmin(P.header_stamp, tmin); t1 = fmax(P.header_stamp, tmax); }
    P.bounds_out[0] = t0; P.bounds_out[1] = t1; P.bounds_out[2] = tmin; P.bounds_out[3] = tmax; P.bounds_out[4] = (double)ns;
    s_b[0] = t0; s_b[1] = t1;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < P.bounds_n_jobs; j += kIngestThreads) { P.bounds_jobs[j].scan_start_time = s_b[0]; P.bounds_jobs[j].scan_end_time = s_b[1]; }
}

// pass 1 (only when the message carries per-point time): does ANY raw time value exceed 1e6?  (backend_node.py:437-441:
// such a message is in nanoseconds and every value is scaled by 1e-9).  With bounds_parts it also leaves this CTA's
// minimum / maximum raw time value (scaling by a positive constant is monotone, so the unit can be applied afterwards), and
// the CTA that finishes last (ticket in flag[1]) turns the extrema into the scan bounds.
__global__ void __launch_bounds__(kIngestThreads) ingest_time_unit_kernel(const __grid_constant__ IngestParams P) {
  const int64_t i = (int64_t)blockIdx.x * kIngestThreads + threadIdx.x;
  bool big = false;
  double lo = 1.0e300, hi = -1.0e300;
  if (i < P.n_points) {
    const double v = decode_field(P.data + i * P.point_step + P.off[4], P.dtype[4]);
    big = v > 1e6;
    lo = hi = v;
  }
  if (__syncthreads_or(big) && threadIdx.x == 0) atomicOr(P.flag, 1);
  if (!P.bounds_parts) return;
  __shared__ double s_lo[kIngestThreads / 32], s_hi[kIngestThreads / 32];
  __shared__ int s_last;
  for (int o = 16; o > 0; o >>= 1) {
    lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) { s_lo[threadIdx.x >> 5] = lo; s_hi[threadIdx.x >> 5] = hi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kIngestThreads / 32; ++w) { lo = fmin(lo, s_lo[w]); hi = fmax(hi, s_hi[w]); }
    P.bounds_parts[2 * blockIdx.x] = lo; P.bounds_parts[2 * blockIdx.x + 1] = hi;
    __threadfence();
    s_last = atomicAdd(P.flag + 1, 1) == (int)gridDim.x - 1;
  }
  __syncthreads();
  if (s_last && P.bounds_out) { __threadfence(); scan_bounds_block(P, (int)gridDim.x); }
}

// the bounds alone (message without a per-point time field, or an empty message)
__global__ void __launch_bounds__(kIngestThreads) scan_bounds_kernel(const __grid_constant__ IngestParams P) { scan_bounds_block(P, 0); }

__global__ void __launch_bounds__(kIngestThreads) ingest_kernel(const __grid_constant__ IngestParams P) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  const int64_t first = (int64_t)blockIdx.x * kIngestThreads;
  const int count = (int)((P.n_points - first < kIngestThreads) ? P.n_points - first : kIngestThreads);
  const size_t nbytes = (size_t)count * P.point_step;
  const unsigned char* src = P.data + first * P.point_step;
  if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
    uint4* d4 = reinterpret_cast<uint4*>(s_raw);
    for (size_t k = threadIdx.x; k < nbytes / 16; k += kIngestThreads) d4[k] = s4[k];
    for (size_t k = (nbytes / 16) * 16 + threadIdx.x; k < nbytes; k += kIngestThreads) s_raw[k] = src[k];
  } else {
    for (size_t k = threadIdx.x; k < nbytes; k += kIngestThreads) s_raw[k] = src[k];
  }
  __syncthreads();
  if ((int)threadIdx.x >= count) return;
  const int64_t i = first + threadIdx.x;
  const unsigned char* rec = s_raw + (size_t)threadIdx.x * P.point_step;
  double c[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double v = decode_field(rec + P.off[k], P.dtype[k]);
    c[k] = isnan(v) ? P.sentinel : (isinf(v) ? (v > 0 ? P.sentinel : -P.sentinel) : v);   // np.nan_to_num, :431-433
  }
  const double ringv = decode_field(rec + P.off[3], P.dtype[3]);
  double t;
  if (P.off[4] >= 0) {
    const double t_raw = decode_field(rec + P.off[4], P.dtype[4]);
    t = (*P.flag) ? t_raw * 1e-9 : t_raw;
  } else {
    t = P.header_stamp;   // :443
  }
  // range window in the LiDAR frame, :448-458 (products and sums rounded separately, as NumPy evaluates x*x + y*y + z*z)
  const double dist = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(c[0], c[0]), __dmul_rn(c[1], c[1])), __dmul_rn(c[2], c[2])));
  const double a = (dist - P.min_r) / P.sigma, b = (P.max_r - dist) / P.sigma;
  const double w_raw = (1.0 / (1.0 + exp(-a))) * (1.0 / (1.0 + exp(-b)));
  const double w = w_raw * (1.0 - P.weight_floor) + P.weight_floor;
  double q[3];
  if (P.transform) {   // p_base = R_base_lidar p_lidar + t_base_lidar, :1677
#pragma unroll
    for (int r = 0; r < 3; ++r) q[r] = fma(P.R[r * 3 + 2], c[2], fma(P.R[r * 3 + 1], c[1], P.R[r * 3] * c[0])) + P.t[r];
  } else {
    q[0] = c[0]; q[1] = c[1]; q[2] = c[2];
  }
  // compact rows (GCL_JOB_COMPACT_ROWS): points / times only for the rows PointBudgetResample selects (raw row = k * stride,
  // point_budget.py:160-163), at index k; weights for EVERY raw row (mass preservation needs their sum, :69)
  const int64_t cs = P.compact_stride > 1 ? P.compact_stride : 1;
  const bool keep = (i % cs) == 0;
  const int64_t ip = i / cs;
  if (P.points_out && keep) { P.points_out[ip * 3] = q[0]; P.points_out[ip * 3 + 1] = q[1]; P.points_out[ip * 3 + 2] = q[2]; }
  if (P.timestamps_out && keep) P.timestamps_out[ip] = t;
  if (P.weights_out) P.weights_out[i] = w;
  if (P.ring_out) P.ring_out[i] = (unsigned char)(long long)ringv;   // astype(uint8) keeps the low byte
  if (P.tag_out) P.tag_out[i] = 0;
  if (P.packed_xyzt) {
    const double pt0 = P.bounds_dev ? P.bounds_dev[0] : P.packed_t0;
    if (keep) P.packed_xyzt[ip] = make_float4((float)q[0], (float)q[1], (float)q[2], (float)(t - pt0));
    P.packed_w[i] = (float)w;
  }
}

cudaError_t launch_ingest(const IngestParams& P, cudaStream_t stream, int* launches) {
  *launches = 0;
  if (P.n_points <= 0 && !P.bounds_out) return cudaSuccess;
  const int grid = (int)((P.n_points + kIngestThreads - 1) / kIngestThreads);
  cudaError_t e;
  if (!P.flag_is_zero && (e = cudaMemsetAsync(P.flag, 0, 2 * sizeof(int), stream)) != cudaSuccess) return e;
  if (P.off[4] >= 0 && grid > 0) {
    ingest_time_unit_kernel<<<grid, kIngestThreads, 0, stream>>>(P);   // with bounds_out: its last CTA writes the bounds
    *launches += 1;
  } else if (P.bounds_out) {
    scan_bounds_kernel<<<1, kIngestThreads, 0, stream>>>(P);
    *launches += 1;
  }
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  if (P.defer_decode || grid == 0) return cudaSuccess;
  return launch_ingest_decode(P, stream, launches);
}

// the decode pass alone (after launch_ingest with defer_decode: other work that needs the bounds may be enqueued in between)
cudaError_t launch_ingest_decode(const IngestParams& P, cudaStream_t stream, int* launches) {
  if (P.n_points <= 0) return cudaSuccess;
  const int grid = (int)((P.n_points + kIngestThreads - 1) / kIngestThreads);
  const size_t smem = (size_t)kIngestThreads * P.point_step;
  static std::atomic<uint64_t> optin{0};
  if (cudaError_t e = optin_dynamic_smem(ingest_kernel, 200 * 1024, optin)) return e;
  ingest_kernel<<<grid, kIngestThreads, smem, stream>>>(P);
  *launches += 1;
  return cudaGetLastError();
}

}  // namespace gcl
