// gcl_evidence.cu -- K3: map-derived statistics, Matrix Fisher rotation evidence, planar translation evidence and
// the 22-D embedding, batched over hypotheses (one warp per hypothesis); the map maintenance kernel; and the
// sequential "chain" kernel that carries map + measurement-noise state from scan to scan on the device.
//
// Replaces archive/bin_atlas.py:166-224 and archive/legacy_operators/matrix_fisher_evidence.py:83-147 (scatter
// metrics), :155-256 + :264-394 (Matrix Fisher), :413-499 + :502-671 (planar translation), :729-756 (22-D);
// pkg/backend/operators/measurement_noise_iw_jax.py:37-127 (LiDAR measurement-noise inverse-Wishart block).
// The 3x3 SVD / symmetric eigen / LU routines are in gcl_math.cuh.  All arithmetic is binary64.
#include "gcl_internal.cuh"

namespace gcl {

// ---- map-derived statistics of one bin (archive/bin_atlas.py:166-200) -----------------------------------------------
__device__ __forceinline__ void map_derived_bin(const double* S_dir, const double* N_dir, const double* N_pos,
                                                const double* sum_p, const double* sum_ppT, int B, int b, double eps_mass,
                                                double eps_psd, double* out) {
  double* mu = out;           // [3B]
  double* kap = out + 3 * B;  // [B]
  double* cen = out + 4 * B;  // [3B]
  double* Sig = out + 7 * B;  // [9B]
  const double s0 = S_dir[b * 3], s1 = S_dir[b * 3 + 1], s2 = S_dir[b * 3 + 2];
  const double n = sqrt(s0 * s0 + s1 * s1 + s2 * s2);
  const double dn = n + eps_mass;  // _safe_normalize_jax, primitives.py:478-484
  mu[b * 3] = s0 / dn; mu[b * 3 + 1] = s1 / dn; mu[b * 3 + 2] = s2 / dn;
  const double invNd = 1.0 / (N_dir[b] + eps_mass + kF64Eps);
  kap[b] = kappa_from_resultant(n * invNd, 1e-6, 3.0, 0.8, 0.03);
  const double invNp = 1.0 / (N_pos[b] + eps_mass + kF64Eps);
  const double c[3] = {sum_p[b * 3] * invNp, sum_p[b * 3 + 1] * invNp, sum_p[b * 3 + 2] * invNp};
  double Sraw[9], Spsd[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) Sraw[i * 3 + j] = sum_ppT[b * 9 + i * 3 + j] * invNp - c[i] * c[j];
  psd_project3_fast(Sraw, eps_psd, Spsd);
  for (int i = 0; i < 3; ++i) cen[b * 3 + i] = c[i];
  for (int i = 0; i < 9; ++i) Sig[b * 9 + i] = Spsd[i];
}

__global__ void map_derived_kernel(const double* __restrict__ S_dir, const double* __restrict__ N_dir,
                                   const double* __restrict__ N_pos, const double* __restrict__ sum_p,
                                   const double* __restrict__ sum_ppT, int B, double eps_mass, double eps_psd,
                                   double* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) map_derived_bin(S_dir, N_dir, N_pos, sum_p, sum_ppT, B, b, eps_mass, eps_psd, out);
}

cudaError_t launch_map_derived(const double* S_dir, const double* N_dir, const double* N_pos, const double* sum_p,
                               const double* sum_ppT, int n_bins, double eps_mass, double eps_psd, double* out,
                               cudaStream_t stream) {
  const int threads = 64;
  map_derived_kernel<<<(n_bins + threads - 1) / threads, threads, 0, stream>>>(S_dir, N_dir, N_pos, sum_p, sum_ppT, n_bins,
                                                                                eps_mass, eps_psd, out);
  return cudaGetLastError();
}

// compute_scatter_metrics, matrix_fisher_evidence.py:83-147 -> [l1,l2,l3, linearity, planarity, sphericity, anisotropy, eff_rank]
__device__ inline void scatter_metrics(const double* S, double N_total, double eps, double* out8, double* vecs9) {
  double T[9];
  const double invN = 1.0 / (N_total + eps);
  for (int i = 0; i < 9; ++i) T[i] = S[i] * invN;
  double val[3], V[9];
  sym_eig3(T, val, V);  // ascending
  const double l1 = fmax(val[2], 0.0), l2 = fmax(val[1], 0.0), l3 = fmax(val[0], 0.0);
  for (int k = 0; k < 3; ++k) { vecs9[k * 3 + 0] = V[k * 3 + 2]; vecs9[k * 3 + 1] = V[k * 3 + 1]; vecs9[k * 3 + 2] = V[k * 3 + 0]; }
  const double inv1 = 1.0 / (l1 + eps);
  const double tot = l1 + l2 + l3 + eps;
  const double p1 = l1 / tot, p2 = l2 / tot, p3 = l3 / tot;
  const double ent = -(p1 * log(p1 + eps) + p2 * log(p2 + eps) + p3 * log(p3 + eps));
  out8[0] = l1; out8[1] = l2; out8[2] = l3;
  out8[3] = (l1 - l2) * inv1; out8[4] = (l2 - l3) * inv1; out8[5] = l3 * inv1; out8[6] = 1.0 - l3 * inv1;
  out8[7] = exp(ent);
}

__device__ __forceinline__ double warp_sum(double v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- one hypothesis, one warp ---------------------------------------------------------------------------------------
// iw_out (may be null): [sum_b w_b r_b r_b^T (9), sum_b w_b] with r_b = t_b - t_wls, the sufficient statistics of
// lidar_meas_iw_suffstats_from_translation_residuals_jax (measurement_noise_iw_jax.py:104-127).
__device__ void pose_evidence_warp(int lane, const double* st, const double* map_S_dir, const double* map_S_sc,
                                   const double* map_N_dir, const double* map_N_pos, const double* map_derived,
                                   const double* pose, const double* R_hat_override, int B, double eps_psd, double eps,
                                   int do_rotation, int do_translation, double* ev, double* iw_out) {
  const StatsLayout SL(B);
  const double* sN = st + SL.N();
  const double* s_dir = st + SL.s_dir();
  const double* sS = st + SL.S_scatter();
  const double* p_bar = st + SL.p_bar();
  const double* Sig_p = st + SL.Sigma_p();
  const double* m_cen = map_derived + 4 * B;
  const double* m_Sig = map_derived + 7 * B;
  for (int i = lane; i < GCL_EVID_DOUBLES; i += 32) ev[i] = 0.0;
  __syncwarp();

  // ---- pass 1 over bins: cross-covariance H, effective mass, scatter totals ------------------------------------
  double H[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, Neff = 0, scanT[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, mapT[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  double sumNs = 0, sumNm = 0;
  for (int b = lane; b < B; b += 32) {
    const double Ns = sN[b], Nm = map_N_dir[b];
    const double ss[3] = {s_dir[b * 3], s_dir[b * 3 + 1], s_dir[b * 3 + 2]};
    const double sm[3] = {map_S_dir[b * 3], map_S_dir[b * 3 + 1], map_S_dir[b * 3 + 2]};
    const double wb = sqrt(Ns * Nm + eps);
    const double ns = sqrt(ss[0] * ss[0] + ss[1] * ss[1] + ss[2] * ss[2]);
    const double nm = sqrt(sm[0] * sm[0] + sm[1] * sm[1] + sm[2] * sm[2]);
    const double Rs = ns * (1.0 / (Ns + eps)), Rm = nm * (1.0 / (Nm + eps));
    const double w = wb * (Rs * Rm);
    const double ins = 1.0 / (ns + eps), inm = 1.0 / (nm + eps);
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) H[i * 3 + j] += w * (sm[i] * inm) * (ss[j] * ins);
    Neff += w;
    for (int i = 0; i < 9; ++i) { scanT[i] += sS[b * 9 + i]; mapT[i] += map_S_sc[b * 9 + i]; }
    sumNs += Ns; sumNm += Nm;
  }
  for (int i = 0; i < 9; ++i) { H[i] = warp_sum(H[i]); scanT[i] = warp_sum(scanT[i]); mapT[i] = warp_sum(mapT[i]); }
  Neff = warp_sum(Neff); sumNs = warp_sum(sumNs); sumNm = warp_sum(sumNm);

  double R_hat[9];
  double Rp[9];
  so3_exp(pose + 3, Rp);  // R_pred = so3_exp(pose[3:6]), matrix_fisher_evidence.py:311-312
  // independent 3x3 eigen problems run on separate lanes: 0 = SVD chain, 1 = scan scatter metrics, 2 = map scatter metrics
  double map_eval[3] = {0, 0, 0};
  // lanes 1 and 2 run the same code on different matrices (scan / map scatter), so they execute together
  if (lane == 1 || lane == 2) {
    double m8[8], v9[9];
    scatter_metrics(lane == 1 ? scanT : mapT, lane == 1 ? sumNs : sumNm, eps, m8, v9);  // map: same matrix as T_map (:572-587)
    if (do_rotation) {
      const int om = lane == 1 ? GCL_EV_SCAN_METRICS : GCL_EV_MAP_METRICS, ov = lane == 1 ? GCL_EV_SCAN_EIGVECS : GCL_EV_MAP_EIGVECS;
      for (int i = 0; i < 8; ++i) ev[om + i] = m8[i];
      for (int i = 0; i < 9; ++i) ev[ov + i] = v9[i];
    }
    map_eval[0] = m8[0]; map_eval[2] = m8[2];
  }
  if (do_rotation) {
    if (lane == 0) {
      double U[9], s[3], V[9];
      svd3(H, U, s, V);
      double UVt[9];
      mat3_mul_nt(U, V, UVt);
      const double det = mat3_det(UVt);
      const double sg = (det > 0.0) ? 1.0 : ((det < 0.0) ? -1.0 : 0.0);  // jnp.sign, :219-222
      U[2] *= sg; U[5] *= sg; U[8] *= sg;
      double Rmf[9];
      mat3_mul_nt(U, V, Rmf);
      const double ld[3] = {s[1] + s[2], s[0] + s[2], s[0] + s[1]};    // unsigned singular values, :240-249
      double Lraw[9];
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) Lraw[i * 3 + j] = V[i * 3 + 0] * ld[0] * V[j * 3 + 0] + V[i * 3 + 1] * ld[1] * V[j * 3 + 1] + V[i * 3 + 2] * ld[2] * V[j * 3 + 2];
      double Rerr[9], delta[3];
      mat3_mul_tn(Rp, Rmf, Rerr);
      so3_log(Rerr, delta);
      double Lpsd[9], pc[6];
      pc[0] = psd_project3_fast(Lraw, eps_psd, Lpsd);
      double h[3];
      for (int i = 0; i < 3; ++i) h[i] = Lpsd[i * 3] * delta[0] + Lpsd[i * 3 + 1] * delta[1] + Lpsd[i * 3 + 2] * delta[2];
      const double nll = 0.5 * (delta[0] * h[0] + delta[1] * h[1] + delta[2] * h[2]);
      for (int i = 0; i < 9; ++i) { ev[GCL_EV_R_MF + i] = Rmf[i]; ev[GCL_EV_L_ROT + i] = Lpsd[i]; }
      for (int i = 0; i < 3; ++i) { ev[GCL_EV_H_ROT + i] = h[i]; ev[GCL_EV_DELTA_ROT + i] = delta[i]; ev[GCL_EV_SVD_S + i] = s[i]; }
      const double emin = fmin(s[0], fmin(s[1], s[2])), emax = fmax(s[0], fmax(s[1], s[2]));
      double* c = ev + GCL_EV_ROT_CERT;
      c[0] = emin; c[1] = emax; c[2] = emax / (emin + eps);
      c[3] = (double)((s[0] < eps) + (s[1] < eps) + (s[2] < eps));
      c[4] = nll / (Neff + eps); c[5] = s[0] + s[1] + s[2]; c[6] = pc[0]; c[7] = eps / (Neff + eps); c[8] = nll; c[9] = Neff;
      for (int i = 0; i < 3; ++i) {
        ev[GCL_EV_H22 + 3 + i] = h[i];
        for (int j = 0; j < 3; ++j) ev[GCL_EV_L22 + (3 + i) * 22 + 3 + j] = Lpsd[i * 3 + j];
      }
      for (int i = 0; i < 9; ++i) R_hat[i] = Rmf[i];
    }
    for (int i = 0; i < 9; ++i) R_hat[i] = __shfl_sync(0xffffffffu, R_hat[i], 0);
  }
  const double l1m = __shfl_sync(0xffffffffu, map_eval[0], 2), l3m = __shfl_sync(0xffffffffu, map_eval[2], 2);
  if (R_hat_override)
    for (int i = 0; i < 9; ++i) R_hat[i] = R_hat_override[i];
  if (!do_translation) return;

  // ---- pass 2 over bins: per-bin WLS terms ----------------------------------------------------------------------
  double Lf[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, hf[3] = {0, 0, 0}, NeffT = 0;
  for (int b = lane; b < B; b += 32) {
    double tb[3];
    for (int i = 0; i < 3; ++i)
      tb[i] = m_cen[b * 3 + i] - (R_hat[i * 3] * p_bar[b * 3] + R_hat[i * 3 + 1] * p_bar[b * 3 + 1] + R_hat[i * 3 + 2] * p_bar[b * 3 + 2]);
    double RS[9], RSR[9], Sg[9], Si[9];
    mat3_mul(R_hat, Sig_p + b * 9, RS);
    mat3_mul_nt(RS, R_hat, RSR);
    for (int i = 0; i < 9; ++i) Sg[i] = m_Sig[b * 9 + i] + RSR[i] + ((i % 4 == 0) ? eps : 0.0);
    mat3_inv_lu(Sg, Si);  // plain inverse, matrix_fisher_evidence.py:462-466
    const double wb = sqrt(sN[b] * map_N_pos[b] + eps);
    for (int i = 0; i < 9; ++i) Si[i] *= wb;
    for (int i = 0; i < 9; ++i) Lf[i] += Si[i];
    for (int i = 0; i < 3; ++i) hf[i] += Si[i * 3] * tb[0] + Si[i * 3 + 1] * tb[1] + Si[i * 3 + 2] * tb[2];
    NeffT += wb;
  }
  for (int i = 0; i < 9; ++i) Lf[i] = warp_sum(Lf[i]);
  for (int i = 0; i < 3; ++i) hf[i] = warp_sum(hf[i]);
  NeffT = warp_sum(NeffT);
  double tw[3] = {hf[0], hf[1], hf[2]};
  if (lane == 0) {
    // self-adaptive z precision from the map scatter, :572-587.  scatter_metrics clamps eigenvalues at 0, which is what
    // the rule does to lambda3; lambda1 is additionally floored at eps.
    const double zs = fmax(l3m, 0.0) / fmax(l1m, eps);
    double Lreg[9];
    for (int i = 0; i < 9; ++i) Lreg[i] = Lf[i] + ((i % 4 == 0) ? eps : 0.0);
    int piv[3];
    lu3(Lreg, piv);
    lu3_solve(Lreg, piv, tw);  // jnp.linalg.solve, :479
    const double mask[3] = {1.0, 1.0, zs};
    double Lraw[9], Lpsd[9], pc[6], lv[3];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) Lraw[i * 3 + j] = Lf[i * 3 + j] * mask[i] * mask[j];
    double dt[3] = {tw[0] - pose[0], tw[1] - pose[1], tw[2] - pose[2]};
    psd_project3(Lraw, eps_psd, Lpsd, pc, lv);
    double h[3];
    for (int i = 0; i < 3; ++i) h[i] = Lpsd[i * 3] * dt[0] + Lpsd[i * 3 + 1] * dt[1] + Lpsd[i * 3 + 2] * dt[2];
    const double nll = 0.5 * (dt[0] * h[0] + dt[1] * h[1] + dt[2] * h[2]);
    for (int i = 0; i < 3; ++i) { ev[GCL_EV_T_WLS + i] = tw[i]; ev[GCL_EV_H_TRANS + i] = h[i]; ev[GCL_EV_DELTA_TRANS + i] = dt[i]; }
    for (int i = 0; i < 9; ++i) ev[GCL_EV_L_TRANS + i] = Lpsd[i];
    double* c = ev + GCL_EV_TRANS_CERT;
    // eigvalsh(L_trans_psd) = the floored eigenvalues the projection just produced (pc[2] = min, pc[3] = max)
    c[0] = pc[2]; c[1] = pc[3]; c[2] = pc[3] / (pc[2] + eps);
    c[3] = (double)((lv[0] < eps) + (lv[1] < eps) + (lv[2] < eps));
    c[4] = nll / (NeffT + eps); c[5] = pc[0]; c[6] = eps / (NeffT + eps); c[7] = nll; c[8] = NeffT;
    c[9] = 0.5 * (Lpsd[0] + Lpsd[4]); c[10] = Lpsd[8]; c[11] = zs;
    for (int i = 0; i < 3; ++i) {
      ev[GCL_EV_H22 + i] = h[i];
      for (int j = 0; j < 3; ++j) ev[GCL_EV_L22 + i * 22 + j] = Lpsd[i * 3 + j];
    }
  }
  if (iw_out) {  // residual sufficient statistics for the measurement-noise IW update (uniform branch)
    for (int i = 0; i < 3; ++i) tw[i] = __shfl_sync(0xffffffffu, tw[i], 0);
    double rr[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, ws = 0;
    for (int b = lane; b < B; b += 32) {
      double r[3];
      for (int i = 0; i < 3; ++i)
        r[i] = m_cen[b * 3 + i] - (R_hat[i * 3] * p_bar[b * 3] + R_hat[i * 3 + 1] * p_bar[b * 3 + 1] + R_hat[i * 3 + 2] * p_bar[b * 3 + 2]) - tw[i];
      const double wb = sqrt(sN[b] * map_N_pos[b] + eps);
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) rr[i * 3 + j] += wb * r[i] * r[j];
      ws += wb;
    }
    for (int i = 0; i < 9; ++i) rr[i] = warp_sum(rr[i]);
    ws = warp_sum(ws);
    if (lane == 0) {
      for (int i = 0; i < 9; ++i) iw_out[i] = rr[i];
      iw_out[9] = ws;
    }
  }
}

__global__ void __launch_bounds__(32) pose_evidence_kernel(const __grid_constant__ EvidenceParams P) {
  const int job = blockIdx.x;
  pose_evidence_warp(threadIdx.x, P.stats + (size_t)job * P.stats_stride, P.map_S_dir, P.map_S_dir_scatter, P.map_N_dir,
                     P.map_N_pos, P.map_derived, P.pose_pred + (size_t)job * 6,
                     P.R_hat_override ? P.R_hat_override + (size_t)job * 9 : nullptr, P.n_bins, P.eps_psd, P.eps_mass,
                     P.do_rotation, P.do_translation, P.evidence + (size_t)job * GCL_EVID_DOUBLES, nullptr);
}

cudaError_t launch_pose_evidence(const EvidenceParams& P, int n_jobs, cudaStream_t stream) {
  pose_evidence_kernel<<<n_jobs, 32, 0, stream>>>(P);
  return cudaGetLastError();
}

// ---- map maintenance of one bin ("next" row; rule declared in DESIGN.md) -----------------------------
// forget, then add the scan's sums pushed to the world frame by (R, t) with t_z forced to 0 (CHANGELOG.md:693-722).
__device__ __forceinline__ void map_update_bin(const double* stats, const double* R, const double* tw, int B, int b, double gamma,
                                               double* S_dir, double* S_sc, double* N_dir, double* N_pos, double* sum_p,
                                               double* sum_ppT) {
  const StatsLayout SL(B);
  const double t[3] = {tw[0], tw[1], 0.0};
  const double N = stats[SL.N() + b];
  const double* s = stats + SL.s_dir() + b * 3;
  const double* S = stats + SL.S_scatter() + b * 9;
  const double* sp = stats + SL.sum_p() + b * 3;
  const double* spp = stats + SL.sum_ppT() + b * 9;
  double Rs[3], Rp[3], T1[9], RSR[9], RPR[9];
  for (int i = 0; i < 3; ++i) {
    Rs[i] = R[i * 3] * s[0] + R[i * 3 + 1] * s[1] + R[i * 3 + 2] * s[2];
    Rp[i] = R[i * 3] * sp[0] + R[i * 3 + 1] * sp[1] + R[i * 3 + 2] * sp[2];
  }
  mat3_mul(R, S, T1); mat3_mul_nt(T1, R, RSR);
  mat3_mul(R, spp, T1); mat3_mul_nt(T1, R, RPR);
  for (int i = 0; i < 3; ++i) {
    S_dir[b * 3 + i] = gamma * S_dir[b * 3 + i] + Rs[i];
    sum_p[b * 3 + i] = gamma * sum_p[b * 3 + i] + (Rp[i] + N * t[i]);
    for (int j = 0; j < 3; ++j) {
      S_sc[b * 9 + i * 3 + j] = gamma * S_sc[b * 9 + i * 3 + j] + RSR[i * 3 + j];
      sum_ppT[b * 9 + i * 3 + j] = gamma * sum_ppT[b * 9 + i * 3 + j] + (RPR[i * 3 + j] + Rp[i] * t[j] + t[i] * Rp[j] + N * t[i] * t[j]);
    }
  }
  N_dir[b] = gamma * N_dir[b] + N;
  N_pos[b] = gamma * N_pos[b] + N;
}

__global__ void map_update_kernel(const double* __restrict__ stats, const double* __restrict__ Rw,
                                  const double* __restrict__ tw, int B, double gamma, double* S_dir, double* S_sc,
                                  double* N_dir, double* N_pos, double* sum_p, double* sum_ppT) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double R[9], t[3] = {tw[0], tw[1], tw[2]};
  for (int i = 0; i < 9; ++i) R[i] = Rw[i];
  map_update_bin(stats, R, t, B, b, gamma, S_dir, S_sc, N_dir, N_pos, sum_p, sum_ppT);
}

cudaError_t launch_map_update(const double* stats, const double* R, const double* t, int n_bins, double forgetting,
                              double* S_dir, double* S_dir_scatter, double* N_dir, double* N_pos, double* sum_p,
                              double* sum_ppT, cudaStream_t stream) {
  const int threads = 64;
  map_update_kernel<<<(n_bins + threads - 1) / threads, threads, 0, stream>>>(stats, R, t, n_bins, forgetting, S_dir,
                                                                               S_dir_scatter, N_dir, N_pos, sum_p, sum_ppT);
  return cudaGetLastError();
}

// ---- LiDAR measurement-noise inverse-Wishart block -------------------------------------------------------------------
__device__ __forceinline__ double softplus64(double x) { return (x > 0.0 ? x : 0.0) + log1p(exp(-fabs(x))); }

// IW mode Sigma = PSD(Psi / (nu + p + 1)), p = 3   (measurement_noise_iw_jax.py:37-57)
__device__ inline void iw_sigma_mode(const double* iw /* [nu, Psi(9)] */, double eps_psd, double* Sigma) {
  double S[9];
  const double den = iw[0] + 3.0 + 1.0;
  for (int i = 0; i < 9; ++i) S[i] = iw[1 + i] / den;
  psd_project3_fast(S, eps_psd, Sigma);
}

// suffstats from residuals (:104-127) then the forgetful update with the smooth nu projection (:60-100)
__device__ inline void iw_update(double* iw, const double* rr_w /* [sum w r r^T (9), sum w] */, double rho, double nu_weak_add,
                                 double nu_max, double eps_mass, double eps_psd) {
  double rrT[9], dPsi[9];
  const double wsum = rr_w[9] + eps_mass;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) rrT[i * 3 + j] = 0.5 * (rr_w[i * 3 + j] + rr_w[j * 3 + i]) / wsum;
  psd_project3_fast(rrT, eps_psd, dPsi);
  double Praw[9], Ppsd[9];
  for (int i = 0; i < 9; ++i) Praw[i] = rho * iw[1 + i] + dPsi[i];
  psd_project3_fast(Praw, eps_psd, Ppsd);  // symmetrises first
  const double nu_raw = rho * iw[0] + 1.0;
  const double nu_min = 3.0 + 1.0 + nu_weak_add;
  const double nu_floor = nu_min + softplus64(nu_raw - nu_min);
  iw[0] = nu_max - softplus64(nu_max - nu_floor);
  for (int i = 0; i < 9; ++i) iw[1 + i] = Ppsd[i];
}

// ---- the sequential chain: one CTA walks a stream of scans, carrying map and IW state on the device -------------------
// per scan s:  [Sigma_p += Sigma_meas]  ->  map-derived stats  ->  pose evidence (H hypotheses, one warp each)
//              ->  measurement-noise IW update  ->  map update with pose_update[s]
constexpr int kChainThreads = 128;

__global__ void __launch_bounds__(kChainThreads) chain_kernel(const __grid_constant__ ChainParams P) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int B = P.n_bins, H = P.n_hyp;
  const StatsLayout SL(B);
  __shared__ double s_iw[10];
  __shared__ double s_sigma[9];
  for (int s = 0; s < P.n_scans; ++s) {
    double* st0 = P.stats + (size_t)s * H * P.stats_stride;
    // 1. measurement noise enters every bin covariance: C_i = Sigma_meas  =>  sum_C / N = Sigma_meas N/(N+eps)
    if (P.iw_state && P.use_sigma_meas) {
      if (tid == 0) iw_sigma_mode(P.iw_state, P.eps_psd, s_sigma);
      __syncthreads();
      for (int idx = tid; idx < B * H; idx += kChainThreads) {
        const int h = idx / B, b = idx - h * B;
        double* st = st0 + (size_t)h * P.stats_stride;
        const double N = st[SL.N() + b];
        const double invN = 1.0 / (N + P.eps_mass + kF64Eps);
        const double* pb = st + SL.p_bar() + b * 3;
        const double* spp = st + SL.sum_ppT() + b * 9;
        double Sraw[9], Spsd[9];
        for (int i = 0; i < 3; ++i)
          for (int j = 0; j < 3; ++j) Sraw[i * 3 + j] = spp[i * 3 + j] * invN - pb[i] * pb[j] + s_sigma[i * 3 + j] * (N * invN);
        psd_project3_fast(Sraw, P.eps_psd, Spsd);
        for (int i = 0; i < 9; ++i) st[SL.Sigma_p() + b * 9 + i] = Spsd[i];
      }
    }
    // 2. map-derived statistics
    for (int b = tid; b < B; b += kChainThreads)
      map_derived_bin(P.map_S_dir, P.map_N_dir, P.map_N_pos, P.map_sum_p, P.map_sum_ppT, B, b, P.eps_mass, P.eps_psd, P.map_derived);
    __syncthreads();
    // 3. pose evidence, one warp per hypothesis
    for (int h = warp; h < H; h += kChainThreads / 32)
      pose_evidence_warp(lane, st0 + (size_t)h * P.stats_stride, P.map_S_dir, P.map_S_dir_scatter, P.map_N_dir, P.map_N_pos,
                         P.map_derived, P.pose_pred + ((size_t)s * H + h) * 6, nullptr, B, P.eps_psd, P.eps_mass, 1, 1,
                         P.evidence + ((size_t)s * H + h) * GCL_EVID_DOUBLES, (h == 0 && P.iw_state) ? s_iw : nullptr);
    __syncthreads();
    // 4. IW update from hypothesis 0's translation residuals
    if (P.iw_state && tid == 0) iw_update(P.iw_state, s_iw, P.iw_rho, P.iw_nu_weak_add, P.iw_nu_max, P.eps_mass, P.eps_psd);
    // 5. map update (hypothesis 0's statistics, pose_update[s])
    if (P.pose_update) {
      double R[9];
      so3_exp(P.pose_update + (size_t)s * 6 + 3, R);
      for (int b = tid; b < B; b += kChainThreads)
        map_update_bin(st0, R, P.pose_update + (size_t)s * 6, B, b, P.forgetting, P.map_S_dir, P.map_S_dir_scatter, P.map_N_dir,
                       P.map_N_pos, P.map_sum_p, P.map_sum_ppT);
    }
    __syncthreads();
  }
}

cudaError_t launch_chain(const ChainParams& P, cudaStream_t stream) {
  chain_kernel<<<1, kChainThreads, 0, stream>>>(P);
  return cudaGetLastError();
}

}  // namespace gcl
