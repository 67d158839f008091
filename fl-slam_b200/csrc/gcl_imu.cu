// gcl_imu.cu -- IMU preintegration over a scan window and the deskew twist it yields.
//
// Replaces smooth_window_weights + preintegrate_imu_relative_pose_jax (pkg/backend/operators/imu_preintegration.py:19-43,
// :46-147) + se3_log (pkg/common/geometry/se3_jax.py:179-256) as called at pkg/backend/pipeline.py:437-483.  The reference
// walks the M samples with a sequential lax.scan (its dominant per-scan cost, docs/PIPELINE_DESIGN_GAPS.md:200-214).  Here
// one CTA of 256 threads handles one hypothesis (start orientation + biases); the recurrences are evaluated as scans:
//   R_k = R_start dR_0 ... dR_{k-1}      exclusive prefix PRODUCT of the per-sample rotations (thread-local products, then a
//                                        warp scan of 3x3 matrices and a pass over the warp totals)
//   v, p                                 prefix sums once R_k is known (p needs the incoming v of every lane's chunk)
// Re-association changes the result at the 1e-15 level against the sample-by-sample order.
#include "gcl_internal.cuh"

namespace gcl {

__device__ __forceinline__ void se3_log6(const double* T, double* xi) {   // se3_jax.py:221-256
  double R[9], phi[3];
  so3_exp(T + 3, R);
  so3_log(R, phi);
  const double th2 = phi[0] * phi[0] + phi[1] * phi[1] + phi[2] * phi[2];
  const double th = sqrt(th2);
  const double D = (th < kSmallAngle) ? (1.0 / 12.0 + th2 / 720.0)
                                      : (1.0 / th2 - (1.0 + cos(th)) / (2.0 * th * sin(th) + 1e-12));   // :201-215
  const double K[9] = {0.0, -phi[2], phi[1], phi[2], 0.0, -phi[0], -phi[1], phi[0], 0.0};
  double K2[9];
  mat3_mul(K, K, K2);
  for (int i = 0; i < 3; ++i) {
    double r = 0.0;
    for (int j = 0; j < 3; ++j) r += (((i == j) ? 1.0 : 0.0) - 0.5 * K[i * 3 + j] + D * K2[i * 3 + j]) * T[j];
    xi[i] = r;
    xi[3 + i] = phi[i];
  }
}

__device__ __forceinline__ double imu_weight(const ImuParams& P, int i) {
  if (P.weights) return P.weights[i];
  const double sigma = fmax(P.sigma, 1e-6);
  const double t0 = P.bounds_dev ? P.bounds_dev[0] : P.t0, t1 = P.bounds_dev ? P.bounds_dev[1] : P.t1;
  return window_weight64(P.stamps[i], t0, t1, sigma, 1e-12);   // GC_WEIGHT_FLOOR, constants.py:256
}

constexpr int kImuThreads = 256;            // one CTA per hypothesis; 512 samples -> 2 per thread
constexpr int kImuWarps = kImuThreads / 32;

__global__ void __launch_bounds__(kImuThreads) imu_preintegrate_kernel(const __grid_constant__ ImuParams P) {
  __shared__ double s_wtot[kImuWarps][9];   // per-warp rotation products
  __shared__ double s_vtot[kImuWarps][3];   // per-warp velocity gains
  __shared__ double s_sum[kImuWarps][17];   // per-warp totals
  const int h = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, M = P.M;
  const int C = (M + kImuThreads - 1) / kImuThreads;
  const int i0 = tid * C < M ? tid * C : M, i1 = (tid + 1) * C < M ? (tid + 1) * C : M;
  const double bg[3] = {P.gyro_bias[h * 3], P.gyro_bias[h * 3 + 1], P.gyro_bias[h * 3 + 2]};
  const double ba[3] = {P.accel_bias[h * 3], P.accel_bias[h * 3 + 1], P.accel_bias[h * 3 + 2]};
  double R_start[9];
  so3_exp(P.rotvec_start + h * 3, R_start);

  // ---- pass 1: product of this thread's per-sample rotations, sum of weights --------------------------------------------
  // (with at most kKeep samples per thread -- 512 samples: 2 -- the window weight and the sample rotation are kept in registers
  // for pass 2 instead of being evaluated again: two exponentials and a Rodrigues formula per sample)
  constexpr int kKeep = 2;
  const bool keep = C <= kKeep;
  double k_dte[kKeep] = {0, 0}, k_dR[kKeep][9];
  double Pm[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, ess = 0.0;
  for (int i = i0; i < i1; ++i) {
    const double w = imu_weight(P, i);
    if (h == 0 && P.weights_out) P.weights_out[i] = w;
    const double dt = (i + 1 < M) ? fmax(P.stamps[i + 1] - P.stamps[i], 0.0) : 0.0;   // last dt forced to 0, :81-82
    const double dte = w * dt;
    const double om[3] = {(P.gyro[i * 3] - bg[0]) * dte, (P.gyro[i * 3 + 1] - bg[1]) * dte, (P.gyro[i * 3 + 2] - bg[2]) * dte};
    double dR[9], T[9];
    so3_exp(om, dR);
    mat3_mul(Pm, dR, T);
    for (int k = 0; k < 9; ++k) Pm[k] = T[k];
    ess += w;
    if (keep) {
#pragma unroll
      for (int j = 0; j < kKeep; ++j)
        if (i - i0 == j) {
          k_dte[j] = dte;
#pragma unroll
          for (int k = 0; k < 9; ++k) k_dR[j][k] = dR[k];
        }
    }
  }
  // ---- inclusive scan of the thread products (earlier threads multiply from the LEFT): warp scan, then warp totals --------
  double Q[9];
  for (int k = 0; k < 9; ++k) Q[k] = Pm[k];
  for (int o = 1; o < 32; o <<= 1) {
    double L[9], T[9];
    for (int k = 0; k < 9; ++k) L[k] = __shfl_up_sync(0xffffffffu, Q[k], o);
    if (lane >= o) {
      mat3_mul(L, Q, T);
      for (int k = 0; k < 9; ++k) Q[k] = T[k];
    }
  }
  if (lane == 31)
    for (int k = 0; k < 9; ++k) s_wtot[warp][k] = Q[k];
  __syncthreads();
  double E[9];   // exclusive prefix of this thread = (warps before) * (lanes before)
  for (int k = 0; k < 9; ++k) {
    const double up = __shfl_up_sync(0xffffffffu, Q[k], 1);
    E[k] = (lane == 0) ? ((k % 4 == 0) ? 1.0 : 0.0) : up;
  }
  double Wp[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, Qtot[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  for (int w = 0; w < kImuWarps; ++w) {
    double T[9];
    if (w == warp)
      for (int k = 0; k < 9; ++k) Wp[k] = Qtot[k];
    mat3_mul(Qtot, s_wtot[w], T);
    for (int k = 0; k < 9; ++k) Qtot[k] = T[k];
  }
  double Rc[9];   // orientation at the start of this thread's chunk
  {
    double T[9];
    mat3_mul(Wp, E, T);
    mat3_mul(R_start, T, Rc);
  }

  // ---- pass 2: accelerations in the world frame, local velocity / position integrals, weighted means ---------------------
  double v[3] = {0, 0, 0}, p[3] = {0, 0, 0}, Tl = 0.0, s_ab[3] = {0, 0, 0}, s_awn[3] = {0, 0, 0}, s_aw[3] = {0, 0, 0};
  for (int i = i0; i < i1; ++i) {
    double dte;
    if (keep) {
      dte = (i - i0 == 0) ? k_dte[0] : k_dte[1];
    } else {
      const double w = imu_weight(P, i);
      const double dt = (i + 1 < M) ? fmax(P.stamps[i + 1] - P.stamps[i], 0.0) : 0.0;
      dte = w * dt;
    }
    const double ab[3] = {P.accel[i * 3] - ba[0], P.accel[i * 3 + 1] - ba[1], P.accel[i * 3 + 2] - ba[2]};
    double awn[3], aw[3];
    for (int r = 0; r < 3; ++r) {
      awn[r] = Rc[r * 3] * ab[0] + Rc[r * 3 + 1] * ab[1] + Rc[r * 3 + 2] * ab[2];
      aw[r] = awn[r] + P.gravity[r];
    }
    Tl += dte;
    for (int r = 0; r < 3; ++r) {
      s_ab[r] += ab[r] * dte; s_awn[r] += awn[r] * dte; s_aw[r] += aw[r] * dte;
      p[r] = p[r] + v[r] * dte + 0.5 * aw[r] * (dte * dte);
      v[r] = v[r] + aw[r] * dte;
    }
    if (i + 1 < i1) {
      double dR[9], T[9];
      if (keep) {   // only the first kept sample can have a successor in this chunk
#pragma unroll
        for (int k = 0; k < 9; ++k) dR[k] = k_dR[0][k];
      } else {
        const double om[3] = {(P.gyro[i * 3] - bg[0]) * dte, (P.gyro[i * 3 + 1] - bg[1]) * dte, (P.gyro[i * 3 + 2] - bg[2]) * dte};
        so3_exp(om, dR);
      }
      mat3_mul(Rc, dR, T);
      for (int k = 0; k < 9; ++k) Rc[k] = T[k];
    }
  }
  // incoming velocity of every chunk = exclusive prefix sum of the local velocity gains; position: p_end = sum_l (p_l + v_in,l T_l)
  double vin[3];
  for (int r = 0; r < 3; ++r) {
    double inc = v[r];
    for (int o = 1; o < 32; o <<= 1) {
      const double up = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += up;
    }
    if (lane == 31) s_vtot[warp][r] = inc;
    vin[r] = inc - v[r];
  }
  __syncthreads();
  for (int r = 0; r < 3; ++r)
    for (int w = 0; w < warp; ++w) vin[r] += s_vtot[w][r];
  double tot[17];
  for (int r = 0; r < 3; ++r) {
    tot[r] = p[r] + vin[r] * Tl; tot[3 + r] = v[r]; tot[6 + r] = s_ab[r]; tot[9 + r] = s_awn[r]; tot[12 + r] = s_aw[r];
  }
  tot[15] = Tl; tot[16] = ess;
  for (int k = 0; k < 17; ++k) {
    for (int o = 16; o > 0; o >>= 1) tot[k] += __shfl_xor_sync(0xffffffffu, tot[k], o);
    if (lane == 0) s_sum[warp][k] = tot[k];
  }
  __syncthreads();
  if (tid != 0) return;
  for (int k = 0; k < 17; ++k) {
    double a = s_sum[0][k];
    for (int w = 1; w < kImuWarps; ++w) a += s_sum[w][k];
    tot[k] = a;
  }

  double* out = P.out + (size_t)h * GCL_IMU_DOUBLES;
  double R_end[9], dR[9], rot[3];
  mat3_mul(R_start, Qtot, R_end);
  mat3_mul_tn(R_start, R_end, dR);          // delta_R = R_start^T R_end, :122-123
  so3_log(dR, rot);
  double dp[3], dv[3];
  for (int r = 0; r < 3; ++r) {             // world -> start body frame, :131-132
    dp[r] = R_start[0 * 3 + r] * tot[0] + R_start[1 * 3 + r] * tot[1] + R_start[2 * 3 + r] * tot[2];
    dv[r] = R_start[0 * 3 + r] * tot[3] + R_start[1 * 3 + r] * tot[4] + R_start[2 * 3 + r] * tot[5];
  }
  const double den = fmax(tot[15], 1e-12);
  for (int r = 0; r < 3; ++r) {
    out[GCL_IMU_DELTA_POSE + r] = dp[r]; out[GCL_IMU_DELTA_POSE + 3 + r] = rot[r];
    out[GCL_IMU_DELTA_P + r] = dp[r]; out[GCL_IMU_DELTA_V + r] = dv[r];
    out[GCL_IMU_A_BODY_MEAN + r] = tot[6 + r] / den; out[GCL_IMU_A_WORLD_NOG_MEAN + r] = tot[9 + r] / den;
    out[GCL_IMU_A_WORLD_MEAN + r] = tot[12 + r] / den;
  }
  for (int k = 0; k < 9; ++k) out[GCL_IMU_DELTA_R + k] = dR[k];
  out[GCL_IMU_ESS] = tot[16];
  out[GCL_IMU_DT_SUM] = tot[15];
  double xi[6];
  se3_log6(out + GCL_IMU_DELTA_POSE, xi);   // xi_body = se3_log(delta_pose), pipeline.py:476
  for (int r = 0; r < 3; ++r) xi[r] *= P.trans_scale;   // rotation-only deskew scales the translation part, :477-481
  for (int r = 0; r < 6; ++r) out[GCL_IMU_XI_BODY + r] = xi[r];
  if (P.jobs_out)
    for (int r = 0; r < 6; ++r) P.jobs_out[h].xi_body[r] = xi[r];
}

cudaError_t launch_imu_preintegrate(const ImuParams& P, int n_hyp, cudaStream_t stream) {
  imu_preintegrate_kernel<<<n_hyp, kImuThreads, 0, stream>>>(P);
  return cudaGetLastError();
}

}  // namespace gcl
