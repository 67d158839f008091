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

constexpr int kChainThreads = 128;   // CTA of the chain / evidence kernels
#ifdef GCL_CHAIN_TRACE
__device__ long long g_chain_trace[32];
#define TR(i) do { if ((threadIdx.x & 31) == 0) { g_chain_trace[i] = clock64(); if ((i) == 0 || (i) == 19) { long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt)); g_chain_trace[(i) == 0 ? 30 : 31] = gt; } } } while (0)
#else
#define TR(i) do { } while (0)
#endif

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

// ---- planar translation evidence, in pieces (matrix_fisher_evidence.py:413-499) -------------------------------------
// pass over bins (lane-strided) + warp sums: L_full = sum W_b, h_full = sum W_b t_b, N_eff = sum w_b; every lane gets the sums
__device__ __forceinline__ void translation_sums_warp(int lane, int B, const double* R_hat, const double* p_bar,
                                                      const double* Sig_p, const double* sN, const double* m_cen,
                                                      const double* m_Sig, const double* map_N_pos, double eps, double* Lf,
                                                      double* hf, double& NeffT) {
  for (int i = 0; i < 9; ++i) Lf[i] = 0.0;
  hf[0] = hf[1] = hf[2] = 0.0;
  NeffT = 0.0;
  for (int b = lane; b < B; b += 32) {
    double tb[3];
    for (int i = 0; i < 3; ++i)
      tb[i] = m_cen[b * 3 + i] - (R_hat[i * 3] * p_bar[b * 3] + R_hat[i * 3 + 1] * p_bar[b * 3 + 1] + R_hat[i * 3 + 2] * p_bar[b * 3 + 2]);
    double RS[9], RSR[9], Sg[9], Si[9];
    mat3_mul(R_hat, Sig_p + b * 9, RS);
    mat3_mul_nt(RS, R_hat, RSR);
    for (int i = 0; i < 9; ++i) Sg[i] = m_Sig[b * 9 + i] + RSR[i] + ((i % 4 == 0) ? eps : 0.0);
    mat3_inv_adj(Sg, Si);  // plain inverse, matrix_fisher_evidence.py:462-466
    const double wb = sqrt(sN[b] * map_N_pos[b] + eps);
    for (int i = 0; i < 9; ++i) Si[i] *= wb;
    for (int i = 0; i < 9; ++i) Lf[i] += Si[i];
    for (int i = 0; i < 3; ++i) hf[i] += Si[i * 3] * tb[0] + Si[i * 3 + 1] * tb[1] + Si[i * 3 + 2] * tb[2];
    NeffT += wb;
  }
  for (int i = 0; i < 9; ++i) Lf[i] = warp_sum(Lf[i]);
  for (int i = 0; i < 3; ++i) hf[i] = warp_sum(hf[i]);
  NeffT = warp_sum(NeffT);
}

// t_wls = solve(L_full + eps I, h_full)  (jnp.linalg.solve, :479); evaluated identically on every lane
__device__ __forceinline__ void translation_solve(const double* Lf, const double* hf, double eps, double* tw) {
  // L_full + eps I is symmetric positive definite: the cofactor inverse needs no pivot search (no dynamically indexed
  // arrays => everything stays in registers) and is as accurate as LU here
  double Lreg[9], Li[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) Lreg[i] = Lf[i] + ((i % 4 == 0) ? eps : 0.0);
  mat3_inv_adj(Lreg, Li);
#pragma unroll
  for (int i = 0; i < 3; ++i) tw[i] = Li[i * 3] * hf[0] + Li[i * 3 + 1] * hf[1] + Li[i * 3 + 2] * hf[2];
}

// planar mask, PSD projection, information vector, certificate and the 22-D blocks (one thread)
__device__ inline void translation_finish(const double* Lf, const double* tw, double NeffT, double l1m, double l3m,
                                          const double* pose, double eps_psd, double eps, double* ev) {
  // self-adaptive z precision from the map scatter, :572-587.  scatter_metrics clamps eigenvalues at 0, which is what
  // the rule does to lambda3; lambda1 is additionally floored at eps.
  const double zs = fmax(l3m, 0.0) / fmax(l1m, eps);
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

// sum_b w_b r_b r_b^T and sum_b w_b with r_b = t_b - t_wls: the sufficient statistics of
// lidar_meas_iw_suffstats_from_translation_residuals_jax (measurement_noise_iw_jax.py:104-127); every lane gets the sums
__device__ __forceinline__ void residual_sums_warp(int lane, int B, const double* R_hat, const double* p_bar, const double* sN,
                                                   const double* m_cen, const double* map_N_pos, const double* tw, double eps,
                                                   double* rr, double& ws) {
  for (int i = 0; i < 9; ++i) rr[i] = 0.0;
  ws = 0.0;
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
}

// ---- one hypothesis, one warp ---------------------------------------------------------------------------------------
// iw_out (may be null): [sum_b w_b r_b r_b^T (9), sum_b w_b] with r_b = t_b - t_wls, the sufficient statistics of
// lidar_meas_iw_suffstats_from_translation_residuals_jax (measurement_noise_iw_jax.py:104-127).
__device__ void pose_evidence_warp(int lane, const double* st, const double* map_S_dir, const double* map_S_sc,
                                   const double* map_N_dir, const double* map_N_pos, const double* map_derived,
                                   const double* pose, const double* R_hat_override, int B, double eps_psd, double eps,
                                   int do_rotation, int do_translation, double* ev, double* iw_out, int role = 0,
                                   double* s_l13 = nullptr, int bar_id = 0, int pre_bar = 0, int iw_bar = 0, int iw_bar_count = 64) {
  // role 0: everything on this warp.  Roles 1 / 2 split one hypothesis over a PAIR of warps (the latency-critical single-scan
  // chain): role 2 computes the scan / map scatter metrics (two symmetric eigen problems) while role 1 runs the SVD chain;
  // they meet at named barrier `bar_id` (64 threads) where role 2 hands over the map eigenvalues through s_l13.
  // pre_bar (role 1, single-scan chain): the rotation part needs neither the bin covariances nor the map-derived statistics,
  // so it runs while other warps prepare those; this warp waits on named barrier `pre_bar` (whole CTA) only before the
  // translation part.  iw_bar: once the residual statistics are written (and the map is no longer read) this warp arrives on
  // named barrier `iw_bar`, BEFORE its lane 0 finishes the translation evidence: the IW update and, with one hypothesis, the
  // map update of the other warps run side by side with that finish.
  const StatsLayout SL(B);
  const double* sN = st + SL.N();
  const double* s_dir = st + SL.s_dir();
  const double* sS = st + SL.S_scatter();
  const double* p_bar = st + SL.p_bar();
  const double* Sig_p = st + SL.Sigma_p();
  const double* m_cen = map_derived + 4 * B;
  const double* m_Sig = map_derived + 7 * B;
  if (role != 2)
    for (int i = lane; i < GCL_EVID_DOUBLES; i += 32)
      if (role == 0 || i < GCL_EV_SCAN_METRICS || i >= GCL_EV_T_WLS) ev[i] = 0.0;   // role 2 owns [SCAN_METRICS, T_WLS)
  __syncwarp();

  // ---- pass 1 over bins: cross-covariance H, effective mass, scatter totals ------------------------------------
  double H[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, Neff = 0, scanT[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, mapT[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  double sumNs = 0, sumNm = 0;
  for (int b = lane; b < B; b += 32) {
    const double Ns = sN[b], Nm = map_N_dir[b];
    if (role != 1) {
      for (int i = 0; i < 9; ++i) { scanT[i] += sS[b * 9 + i]; mapT[i] += map_S_sc[b * 9 + i]; }
      sumNs += Ns; sumNm += Nm;
    }
    if (role == 2) continue;
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
  }
  if (role == 1) TR(2);
  if (role != 2) {
    for (int i = 0; i < 9; ++i) H[i] = warp_sum(H[i]);
    Neff = warp_sum(Neff);
  }
  if (role == 1) TR(3);
  if (role != 1) {
    for (int i = 0; i < 9; ++i) { scanT[i] = warp_sum(scanT[i]); mapT[i] = warp_sum(mapT[i]); }
    sumNs = warp_sum(sumNs); sumNm = warp_sum(sumNm);
  }

  double R_hat[9];
  double Rp[9];
  so3_exp(pose + 3, Rp);  // R_pred = so3_exp(pose[3:6]), matrix_fisher_evidence.py:311-312
  // independent 3x3 eigen problems run on separate lanes: 0 = SVD chain, 1 = scan scatter metrics, 2 = map scatter metrics
  double map_eval[3] = {0, 0, 0};
  // lanes 1 and 2 run the same code on different matrices (scan / map scatter), so they execute together
  if (role != 1 && (lane == 1 || lane == 2)) {
    double m8[8], v9[9];
    scatter_metrics(lane == 1 ? scanT : mapT, lane == 1 ? sumNs : sumNm, eps, m8, v9);  // map: same matrix as T_map (:572-587)
    if (do_rotation) {
      const int om = lane == 1 ? GCL_EV_SCAN_METRICS : GCL_EV_MAP_METRICS, ov = lane == 1 ? GCL_EV_SCAN_EIGVECS : GCL_EV_MAP_EIGVECS;
      for (int i = 0; i < 8; ++i) ev[om + i] = m8[i];
      for (int i = 0; i < 9; ++i) ev[ov + i] = v9[i];
    }
    map_eval[0] = m8[0]; map_eval[2] = m8[2];
    if (role == 2 && lane == 2) { s_l13[0] = m8[0]; s_l13[1] = m8[2]; }
  }
  if (role == 2) {
    TR(14);
    asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
    return;
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
  if (role == 1) TR(4);
  double l1m = __shfl_sync(0xffffffffu, map_eval[0], 2), l3m = __shfl_sync(0xffffffffu, map_eval[2], 2);
  if (role == 1) {
    asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
    l1m = s_l13[0]; l3m = s_l13[1];
  }
  if (R_hat_override)
    for (int i = 0; i < 9; ++i) R_hat[i] = R_hat_override[i];
  if (!do_translation) return;

  double Lf[9], hf[3], NeffT, tw[3];
  if (role == 1) TR(5);
  if (pre_bar) asm volatile("bar.sync %0, %1;" ::"r"(pre_bar), "r"(kChainThreads) : "memory");
  if (role == 1) TR(6);
  translation_sums_warp(lane, B, R_hat, p_bar, Sig_p, sN, m_cen, m_Sig, map_N_pos, eps, Lf, hf, NeffT);
  translation_solve(Lf, hf, eps, tw);
  if (role == 1) TR(7);
  if (iw_out) {  // residual sufficient statistics for the measurement-noise IW update (uniform branch)
    double rr[9], ws;
    residual_sums_warp(lane, B, R_hat, p_bar, sN, m_cen, map_N_pos, tw, eps, rr, ws);
    if (lane == 0) {
      for (int i = 0; i < 9; ++i) iw_out[i] = rr[i];
      iw_out[9] = ws;
    }
  }
  if (iw_bar) {   // this warp is done reading the map and has left the residual statistics: IW update / map update may start
    __threadfence_block();
    __syncwarp();
    asm volatile("bar.arrive %0, %1;" ::"r"(iw_bar), "r"(iw_bar_count) : "memory");
  }
  if (role == 1) TR(8);
  if (lane == 0) translation_finish(Lf, tw, NeffT, l1m, l3m, pose, eps_psd, eps, ev);
  if (role == 1) TR(9);
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
// softplus(x) = max(x,0) + log1p(exp(-|x|)); beyond |x| = 40 the second term (< 4.3e-18) is below half an ulp of the first
// whenever the first is non-zero, and for x < -40 it equals exp(x) to rounding: the shortcuts return the same doubles
__device__ __forceinline__ double softplus64(double x) {
  if (x > 40.0) return x;
  return (x > 0.0 ? x : 0.0) + log1p(exp(-fabs(x)));
}

// IW mode Sigma = PSD(Psi / (nu + p + 1)), p = 3   (measurement_noise_iw_jax.py:37-57)
__device__ inline void iw_sigma_mode(const double* iw /* [nu, Psi(9)] */, double eps_psd, double* Sigma) {
  double S[9];
  const double inv = 1.0 / (iw[0] + 3.0 + 1.0);   // one reciprocal (nine divisions were 0.4 us of the sequential stream core per scan)
  for (int i = 0; i < 9; ++i) S[i] = iw[1 + i] * inv;
  psd_project3_fast(S, eps_psd, Sigma);
}

// suffstats from residuals (:104-127) then the forgetful update with the smooth nu projection (:60-100)
// nu' = nu_max - softplus(nu_max - (nu_min + softplus(rho nu + 1 - nu_min))): the degrees of freedom follow a recurrence of
// their own, independent of the data
__device__ inline double iw_nu_next(double nu, double rho, double nu_weak_add, double nu_max) {
  const double nu_raw = rho * nu + 1.0;
  const double nu_min = 3.0 + 1.0 + nu_weak_add;
  const double nu_floor = nu_min + softplus64(nu_raw - nu_min);
  return nu_max - softplus64(nu_max - nu_floor);
}
__device__ inline void iw_update(double* iw, const double* rr_w /* [sum w r r^T (9), sum w] */, double rho, double nu_weak_add,
                                 double nu_max, double eps_mass, double eps_psd, bool update_nu = true) {
  double rrT[9], dPsi[9];
  const double iw_ = 1.0 / (rr_w[9] + eps_mass);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) rrT[i * 3 + j] = 0.5 * (rr_w[i * 3 + j] + rr_w[j * 3 + i]) * iw_;
  psd_project3_fast(rrT, eps_psd, dPsi);
  double Praw[9], Ppsd[9];
  for (int i = 0; i < 9; ++i) Praw[i] = rho * iw[1 + i] + dPsi[i];
  psd_project3_fast(Praw, eps_psd, Ppsd);  // symmetrises first
  if (update_nu) iw[0] = iw_nu_next(iw[0], rho, nu_weak_add, nu_max);
  for (int i = 0; i < 9; ++i) iw[1 + i] = Ppsd[i];
}

// ---- the chain over a stream of scans ---------------------------------------------------------------------------------
// per scan s:  [Sigma_p += Sigma_meas]  ->  map-derived stats  ->  pose evidence (H hypotheses, one warp each)
//              ->  measurement-noise IW update  ->  map update with pose_update[s]
// (a) chain_kernel: one CTA walks the scans in order -- used for a single scan (the per-scan host call).
// (b) staged stream (n_scans >= kStagedMinScans): only what truly depends on the previous scan stays sequential.
//     The map recurrence m <- gamma m + inc(s) does not depend on the evidence (the update poses are inputs), so the
//     increments are computed in parallel and the recurrence runs elementwise; map-derived statistics and the rotation
//     evidence then run in parallel over scans.  With use_sigma_meas the IW state feeds every bin covariance of the next
//     scan, so ONE warp walks the scans doing just  Sigma_meas -> bin covariances -> t_wls -> residuals -> IW update
//     (inputs of scan s+1 are staged into shared memory by a helper warp meanwhile); the planar mask / PSD projection /
//     certificates of the translation evidence are finished in parallel afterwards.  Same arithmetic, same order within
//     a scan as (a): results agree to the last ulps (to re-association with Sigma_meas fed back, see S4b).
constexpr int kStagedMinScans = 4;

__device__ __forceinline__ void sigma_p_with_meas(const double* st, const StatsLayout& SL, int b, const double* sigma,
                                                  double eps_mass, double eps_psd, double* Spsd) {
  // C_i = Sigma_meas for every point  =>  sum_C / N = Sigma_meas N/(N+eps)   (binning.py:173-182)
  const double N = st[SL.N() + b];
  const double invN = 1.0 / (N + eps_mass + kF64Eps);
  const double* pb = st + SL.p_bar() + b * 3;
  const double* spp = st + SL.sum_ppT() + b * 9;
  double Sraw[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) Sraw[i * 3 + j] = spp[i * 3 + j] * invN - pb[i] * pb[j] + sigma[i * 3 + j] * (N * invN);
  psd_project3_fast(Sraw, eps_psd, Spsd);
}

// CTA-wide copy of n2 16-byte words, eight loads of a thread in flight before its first store (a dependent load / store pair
// per iteration would pay the full L2 latency every time)
__device__ __forceinline__ void copy_batched(double2* __restrict__ dst, const double2* __restrict__ src, int n2, int tid,
                                             int nthreads = kChainThreads) {
  for (int base = 0; base < n2; base += 8 * nthreads) {
    double2 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int i = base + k * nthreads + tid;
      v[k] = i < n2 ? src[i] : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int i = base + k * nthreads + tid;
      if (i < n2) dst[i] = v[k];
    }
  }
}

__global__ void __launch_bounds__(kChainThreads) chain_kernel(const __grid_constant__ ChainParams P) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int B = P.n_bins, H = P.n_hyp;
  const StatsLayout SL(B);
  __shared__ double s_iw[10];
  __shared__ double s_sigma[9];
  __shared__ double s_l13[kChainThreads / 64][2];
  extern __shared__ __align__(16) double s_mder_dyn[];
  double* mder = P.chain_smem ? s_mder_dyn : P.workspace;   // map-derived statistics stay on chip when they fit
  // poses that travel as kernel arguments (host call, few hypotheses): read straight out of the parameter block
  const double* pose_pred = P.pose_inline_n ? P.pose_inline : P.pose_pred;
  const double* pose_update = P.pose_inline_n ? (P.pose_inline_update ? P.pose_inline + (size_t)H * 6 : nullptr) : P.pose_update;
  if (tid == 0) TR(0);
  if (P.stats_src) {
    // pose / map / IW out of page-locked host memory (nothing at all when the poses came as arguments and the device still holds
    // the map and the IW state the previous call left -- see gcl_scan_evidence_begin_host) (16-byte loads, all of a thread's loads in flight before its first store:
    // a PCIe read takes ~1.5 us), and the stage-1 records into the working copy that the chain step rewrites
    const double2* src = reinterpret_cast<const double2*>(P.stage_in_src);
    double2* dst = reinterpret_cast<double2*>(P.stage_in_dst);
    const int n2 = P.stage_in_doubles >> 1;
    double2 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {   // first batch of PCIe reads goes out ...
      const int i = k * kChainThreads + tid;
      v[k] = i < n2 ? __ldcv(src + i) : make_double2(0.0, 0.0);
    }
    copy_batched(reinterpret_cast<double2*>(P.stats), reinterpret_cast<const double2*>(P.stats_src), P.stats_doubles >> 1, tid);   // ... while the records are copied
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int i = k * kChainThreads + tid;
      if (i < n2) dst[i] = v[k];
    }
    for (int base = 8 * kChainThreads; base < n2; base += 8 * kChainThreads) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int i = base + k * kChainThreads + tid;
        v[k] = i < n2 ? __ldcv(src + i) : make_double2(0.0, 0.0);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int i = base + k * kChainThreads + tid;
        if (i < n2) dst[i] = v[k];
      }
    }
    __syncthreads();
  }
  for (int s = 0; s < P.n_scans; ++s) {
    double* st0 = P.stats + (size_t)s * H * P.stats_stride;
    // Warp roles: even warps with a hypothesis run the rotation chain at once (it needs neither the bin covariances nor the
    // map-derived statistics) and wait on barrier 8 only before their translation part; all other warps first prepare
    //   1. the measurement noise in every bin covariance (Sigma_meas on one thread, then bins x hypotheses in parallel),
    //   2. the map-derived statistics,
    // arrive on barrier 8, and then the odd ones take the scatter-metric half of their hypothesis; warp 1 finally runs
    //   4. the IW update as soon as hypothesis 0's residual statistics are there (barrier 9), next to the translation finish.
    if (tid == 0) TR(1);
    const bool role1 = (warp & 1) == 0 && (warp >> 1) < H;
    const bool sigma_step = P.iw_state && P.use_sigma_meas;
    // One hypothesis: warps 2 and 3 have no evidence role, so they also take the map update (as soon as warp 0 has read the map
    // for the last time: barrier 9) and, when the kernel stages its own outputs, carry the finished blocks to the host while
    // warp 0 is still busy -- statistics after step 1, the map after its update; warp 1 the IW state, warp 0 the evidence.
    const bool solo = H == 1;
    const bool solo_out = solo && P.stage_out_dst != nullptr;
    const bool early_map = solo && pose_update != nullptr;
    const int bar9 = 32 + (P.iw_state ? 32 : 0) + (early_map ? 64 : 0);   // 32: nobody waits, barrier not used
    if (role1) {
      bool first = true;
      for (int h = warp >> 1; h < H; h += kChainThreads / 64) {
        pose_evidence_warp(lane, st0 + (size_t)h * P.stats_stride, P.map_S_dir, P.map_S_dir_scatter, P.map_N_dir, P.map_N_pos,
                           mder, pose_pred + ((size_t)s * H + h) * 6, nullptr, B, P.eps_psd, P.eps_mass, 1, 1,
                           P.evidence + ((size_t)s * H + h) * GCL_EVID_DOUBLES, (h == 0 && P.iw_state) ? s_iw : nullptr, 1,
                           s_l13[warp >> 1], 1 + (warp >> 1), first ? 8 : 0, (h == 0 && bar9 > 32) ? 9 : 0, bar9);
        first = false;
      }
      if (solo_out) {   // the evidence record is complete (warp 1 wrote its part before the pair barrier)
        __syncwarp();
        copy_batched(reinterpret_cast<double2*>(P.stage_out_dst + P.out_off_evid), reinterpret_cast<const double2*>(P.evidence),
                     GCL_EVID_DOUBLES >> 1, lane, 32);
      }
    } else {
      // rank among the preparing threads: with H = 1 warps 1, 2, 3 (96 threads), otherwise the odd warps (64 threads)
      const int n_role1 = H < kChainThreads / 64 ? H : kChainThreads / 64;
      const int n_pre = kChainThreads - 32 * n_role1;
      const int pw = (warp & 1) ? (warp >> 1) : (kChainThreads / 64 + (warp >> 1) - n_role1);   // odd warps first
      const int p = pw * 32 + lane;
      if (sigma_step && p == 0) iw_sigma_mode(P.iw_state, P.eps_psd, s_sigma);
      for (int b = p - 32; b >= 0 && b < B; b += n_pre - 32)   // the first warp is busy with Sigma_meas
        map_derived_bin(P.map_S_dir, P.map_N_dir, P.map_N_pos, P.map_sum_p, P.map_sum_ppT, B, b, P.eps_mass, P.eps_psd, mder);
      if (p == 0) TR(10);
      if (p == 32) TR(11);
      if (sigma_step) {
        asm volatile("bar.sync 10, %0;" ::"r"(n_pre) : "memory");
        for (int idx = p; idx < B * H; idx += n_pre) {
          const int h = idx / B, b = idx - h * B;
          double* st = st0 + (size_t)h * P.stats_stride;
          double Spsd[9];
          sigma_p_with_meas(st, SL, b, s_sigma, P.eps_mass, P.eps_psd, Spsd);
          for (int i = 0; i < 9; ++i) st[SL.Sigma_p() + b * 9 + i] = Spsd[i];
        }
      }
      __threadfence_block();
      if (p == 0) TR(12);
      asm volatile("bar.arrive 8, %0;" ::"r"(kChainThreads) : "memory");
      if (solo_out && sigma_step) asm volatile("bar.sync 10, %0;" ::"r"(n_pre) : "memory");   // every bin covariance is in place
      if (solo && warp >= 2) {   // warps 2 and 3: no evidence role with one hypothesis
        const int q = p - 32;      // rank among these 64 threads
        if (solo_out) {
          copy_batched(reinterpret_cast<double2*>(P.stage_out_dst + P.out_off_stats), reinterpret_cast<const double2*>(st0),
                       P.stats_doubles >> 1, q, 64);
        }
        if (early_map) {
          asm volatile("bar.sync 9, %0;" ::"r"(bar9) : "memory");
          double R[9];
          so3_exp(pose_update + (size_t)s * 6 + 3, R);
          for (int b = q; b < B; b += 64)
            map_update_bin(st0, R, pose_update + (size_t)s * 6, B, b, P.forgetting, P.map_S_dir, P.map_S_dir_scatter, P.map_N_dir,
                           P.map_N_pos, P.map_sum_p, P.map_sum_ppT);
          if (solo_out) {
            __threadfence_block();
            asm volatile("bar.sync 11, 64;" ::: "memory");
            copy_batched(reinterpret_cast<double2*>(P.stage_out_dst), reinterpret_cast<const double2*>(P.stage_out_src), (26 * B) >> 1, q, 64);
          }
        }
      }
      if ((warp & 1) && (warp >> 1) < H) {
        for (int h = warp >> 1; h < H; h += kChainThreads / 64)
          pose_evidence_warp(lane, st0 + (size_t)h * P.stats_stride, P.map_S_dir, P.map_S_dir_scatter, P.map_N_dir, P.map_N_pos,
                             mder, pose_pred + ((size_t)s * H + h) * 6, nullptr, B, P.eps_psd, P.eps_mass, 1, 1,
                             P.evidence + ((size_t)s * H + h) * GCL_EVID_DOUBLES, nullptr, 2, s_l13[warp >> 1], 1 + (warp >> 1));
        if (warp == 1 && P.iw_state) {
          asm volatile("bar.sync 9, %0;" ::"r"(bar9) : "memory");
          TR(15);
          if (lane == 0) iw_update(P.iw_state, s_iw, P.iw_rho, P.iw_nu_weak_add, P.iw_nu_max, P.eps_mass, P.eps_psd);
          TR(16);
          if (solo_out) {
            __syncwarp();
            if (lane < 10) P.stage_out_dst[P.out_off_iw + lane] = P.iw_state[lane];
          }
        }
      }
    }
    __syncthreads();
    if (tid == 0) TR(17);
    // 5. map update (hypothesis 0's statistics, pose_update[s]) -- already done by warps 2, 3 when there is one hypothesis
    if (pose_update && !early_map) {
      double R[9];
      so3_exp(pose_update + (size_t)s * 6 + 3, R);
      for (int b = tid; b < B; b += kChainThreads)
        map_update_bin(st0, R, pose_update + (size_t)s * 6, B, b, P.forgetting, P.map_S_dir, P.map_S_dir_scatter, P.map_N_dir,
                       P.map_N_pos, P.map_sum_p, P.map_sum_ppT);
    }
    __syncthreads();
  }
  if (tid == 0) TR(18);
  if (P.stage_out_dst) {
    // map / IW / stats / evidence back to page-locked host memory, then the completion ticket (posted writes stay in order)
    const double2* src = reinterpret_cast<const double2*>(P.stage_out_src);
    double2* dst = reinterpret_cast<double2*>(P.stage_out_dst);
    const int n2 = P.stage_out_doubles >> 1;
    if (P.n_hyp != 1) {
      copy_batched(dst, src, n2, tid);
    } else if (!pose_update) {
      copy_batched(dst, src, (26 * B) >> 1, tid);   // the map was not updated: it rides along unchanged
    }
    __threadfence_system();
    __syncthreads();
    if (tid == 0) {
      *reinterpret_cast<volatile unsigned long long*>(P.done_flag) = P.done_value;
      TR(19);
    }
  }
}
#ifdef GCL_CHAIN_TRACE
extern "C" int gcl_debug_chain_trace(long long* out) { return (int)cudaMemcpyFromSymbol(out, g_chain_trace, sizeof(long long) * 32); }
#endif

// ---- staged stream ----------------------------------------------------------------------------------------------------
// workspace carve (doubles): [single-scan map-derived 16B][map_pre: S x 26B][mder: S x 16B][sigma: S x 16][tpart: S*H x 32]
//                           [pre: S x (32 + 17B, rounded up to even)]
struct StreamWs {
  double *map_pre, *mder, *sigma, *tpart, *pre;
  __host__ __device__ StreamWs(double* ws, int S, int B, int H) {
    map_pre = ws + 16 * (size_t)B;
    mder = map_pre + (size_t)S * 26 * B;
    sigma = mder + (size_t)S * 16 * B;
    tpart = sigma + (size_t)S * 16;
    pre = tpart + (size_t)S * H * 32;
  }
};
// 16 values per bin at a pitch of 17 doubles: the sequential core reads "its" bin with one lane per bin, and a pitch of 16
// doubles (128 bytes) put all 32 lanes on the same shared-memory bank (every LDS.64 a 32-way conflict: 1.1 of its 2.9 us per scan)
constexpr int kPrePitch = 17;
__host__ __device__ inline size_t pre_block_doubles(int B) { return (32 + kPrePitch * (size_t)B + 1) & ~(size_t)1; }
size_t chain_workspace_doubles(int n_scans, int n_hyp, int n_bins) {
  return 16 * (size_t)n_bins + (size_t)n_scans * (26 * (size_t)n_bins + 16 * (size_t)n_bins + 16) + (size_t)n_scans * n_hyp * 32 +
         (size_t)n_scans * pre_block_doubles(n_bins);
}
// pre block of one scan (hypothesis 0), everything of the translation step that does NOT depend on Sigma_meas:
//   header: R_hat[0..9) | S_wtt[9..18) = sum_b w_b t_b t_b^T | S_wt[18..21) | S_w[21]
//   per bin (16): M_b[0..9) = Sigma_c,b + R A_b R^T + eps I | c_b = N/(N+eps) | t_b[10..13) | w_b | flag (A_b provably >= eps_psd)
// with A_b = sum_ppT/N - p_bar p_bar^T, so that  Sigma_c,b + R Sigma_p,b R^T + eps I = M_b + c_b (R Sigma_meas R^T)  whenever the
// PSD projection of A_b + c_b Sigma_meas is the identity (flag), and
//   sum_b w_b r_b r_b^T = S_wtt - S_wt t^T - t S_wt^T + S_w t t^T   for r_b = t_b - t.
// tpart record: Lf[0..9) tw[9..12) NeffT[12] rr[13..22) ws[22]

// S1: world-frame increments of every scan's bin sums (same expressions as map_update_bin, without the recurrence);
//     layout of one scan's block = the six map arrays back to back [S_dir 3B | S_sc 9B | N_dir B | N_pos B | sum_p 3B | sum_ppT 9B]
__global__ void stream_increments_kernel(const __grid_constant__ ChainParams P) {
  const int s = blockIdx.x, B = P.n_bins;
  const StatsLayout SL(B);
  const StreamWs W(P.workspace, P.n_scans, B, P.n_hyp);
  const double* stats = P.stats + (size_t)s * P.n_hyp * P.stats_stride;
  const double* pu = P.pose_update + (size_t)s * 6;
  double R[9];
  so3_exp(pu + 3, R);
  const double t[3] = {pu[0], pu[1], 0.0};
  double* o = W.map_pre + (size_t)s * 26 * B;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const double N = stats[SL.N() + b];
    const double* sd = stats + SL.s_dir() + b * 3;
    const double* S = stats + SL.S_scatter() + b * 9;
    const double* sp = stats + SL.sum_p() + b * 3;
    const double* spp = stats + SL.sum_ppT() + b * 9;
    double Rs[3], Rp[3], T1[9], RSR[9], RPR[9];
    for (int i = 0; i < 3; ++i) {
      Rs[i] = R[i * 3] * sd[0] + R[i * 3 + 1] * sd[1] + R[i * 3 + 2] * sd[2];
      Rp[i] = R[i * 3] * sp[0] + R[i * 3 + 1] * sp[1] + R[i * 3 + 2] * sp[2];
    }
    mat3_mul(R, S, T1); mat3_mul_nt(T1, R, RSR);
    mat3_mul(R, spp, T1); mat3_mul_nt(T1, R, RPR);
    for (int i = 0; i < 3; ++i) {
      o[b * 3 + i] = Rs[i];
      o[14 * B + b * 3 + i] = Rp[i] + N * t[i];
      for (int j = 0; j < 3; ++j) {
        o[3 * B + b * 9 + i * 3 + j] = RSR[i * 3 + j];
        o[17 * B + b * 9 + i * 3 + j] = RPR[i * 3 + j] + Rp[i] * t[j] + t[i] * Rp[j] + N * t[i] * t[j];
      }
    }
    o[12 * B + b] = N;
    o[13 * B + b] = N;
  }
}

// S2: m <- gamma m + inc(s), elementwise; the slot of scan s is overwritten with the map state that scan s SEES
__global__ void stream_map_recurrence_kernel(const __grid_constant__ ChainParams P) {
  const int B = P.n_bins, e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= 26 * B) return;
  const StreamWs W(P.workspace, P.n_scans, B, P.n_hyp);
  double* arr = e < 3 * B ? P.map_S_dir + e : e < 12 * B ? P.map_S_dir_scatter + (e - 3 * B) : e < 13 * B ? P.map_N_dir + (e - 12 * B)
              : e < 14 * B ? P.map_N_pos + (e - 13 * B) : e < 17 * B ? P.map_sum_p + (e - 14 * B) : P.map_sum_ppT + (e - 17 * B);
  double m = *arr;
  const double g = P.forgetting;
  double* slot = W.map_pre + e;
  const size_t pitch = 26 * (size_t)B;
  // increments are read eight scans ahead of the stores (the compiler cannot prove that slot[s] and slot[s+1] differ, and a
  // load that waits for the previous store makes every step an L2 round trip)
  for (int s0 = 0; s0 < P.n_scans; s0 += 8) {
    double inc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) inc[j] = (s0 + j < P.n_scans) ? slot[(size_t)(s0 + j) * pitch] : 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (s0 + j < P.n_scans) {
        slot[(size_t)(s0 + j) * pitch] = m;
        m = g * m + inc[j];
      }
    }
  }
  *arr = m;
}

struct MapView { const double *S_dir, *S_sc, *N_dir, *N_pos, *sum_p, *sum_ppT; };
__device__ __forceinline__ MapView map_view(const ChainParams& P, const StreamWs& W, int s) {
  MapView v;
  if (P.pose_update) {
    const double* m = W.map_pre + (size_t)s * 26 * P.n_bins;
    const int B = P.n_bins;
    v.S_dir = m; v.S_sc = m + 3 * B; v.N_dir = m + 12 * B; v.N_pos = m + 13 * B; v.sum_p = m + 14 * B; v.sum_ppT = m + 17 * B;
  } else {
    v.S_dir = P.map_S_dir; v.S_sc = P.map_S_dir_scatter; v.N_dir = P.map_N_dir; v.N_pos = P.map_N_pos; v.sum_p = P.map_sum_p;
    v.sum_ppT = P.map_sum_ppT;
  }
  return v;
}

// S3: one CTA per scan: map-derived statistics of the map state that scan sees, then one warp per hypothesis.
//     With use_sigma_meas only the rotation part can run here (the bin covariances are not known yet).
__global__ void __launch_bounds__(kChainThreads) stream_evidence_kernel(const __grid_constant__ ChainParams P) {
  const int s = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, B = P.n_bins, H = P.n_hyp;
  const StreamWs W(P.workspace, P.n_scans, B, P.n_hyp);
  const MapView mv = map_view(P, W, s);
  double* mder = W.mder + (size_t)s * 16 * B;
  for (int b = tid; b < B; b += kChainThreads)
    map_derived_bin(mv.S_dir, mv.N_dir, mv.N_pos, mv.sum_p, mv.sum_ppT, B, b, P.eps_mass, P.eps_psd, mder);
  __syncthreads();
  const bool sig = P.iw_state && P.use_sigma_meas;
  for (int h = warp; h < H; h += kChainThreads / 32)
    pose_evidence_warp(lane, P.stats + ((size_t)s * H + h) * P.stats_stride, mv.S_dir, mv.S_sc, mv.N_dir, mv.N_pos, mder,
                       P.pose_pred + ((size_t)s * H + h) * 6, nullptr, B, P.eps_psd, P.eps_mass, 1, sig ? 0 : 1,
                       P.evidence + ((size_t)s * H + h) * GCL_EVID_DOUBLES,
                       (h == 0 && P.iw_state && !sig) ? W.tpart + ((size_t)s * H) * 32 + 13 : nullptr);
  if (!sig) return;
  // everything of hypothesis 0's translation step that does not depend on Sigma_meas (see the pre block above)
  __syncthreads();
  const StatsLayout SL(B);
  const double* st = P.stats + ((size_t)s * H) * P.stats_stride;
  const double* ev = P.evidence + ((size_t)s * H) * GCL_EVID_DOUBLES;
  double* pre = W.pre + (size_t)s * pre_block_doubles(B);
  double R[9];
  for (int i = 0; i < 9; ++i) R[i] = ev[GCL_EV_R_MF + i];
  for (int b = tid; b < B; b += kChainThreads) {
    const double N = st[SL.N() + b];
    const double invN = 1.0 / (N + P.eps_mass + kF64Eps);
    const double* pb = st + SL.p_bar() + b * 3;
    const double* spp = st + SL.sum_ppT() + b * 9;
    double A[9], As[9], RA[9], RAR[9];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) A[i * 3 + j] = spp[i * 3 + j] * invN - pb[i] * pb[j];
    const double a = A[0], d = A[4], f = A[8], bb = 0.5 * (A[1] + A[3]), c = 0.5 * (A[2] + A[6]), e2 = 0.5 * (A[5] + A[7]);
    const double m2 = a * d - bb * bb, det = a * (d * f - e2 * e2) - bb * (bb * f - e2 * c) + c * (bb * e2 - d * c), tr = a + d + f;
    const bool pd = a > 0.0 && m2 > 0.0 && det > 0.0 && 4.0 * det > 2.0 * P.eps_psd * tr * tr;   // as psd_project3_fast
    As[0] = a; As[1] = bb; As[2] = c; As[3] = bb; As[4] = d; As[5] = e2; As[6] = c; As[7] = e2; As[8] = f;
    mat3_mul(R, As, RA);
    mat3_mul_nt(RA, R, RAR);
    double* q = pre + 32 + kPrePitch * (size_t)b;
    for (int i = 0; i < 9; ++i) q[i] = mder[7 * B + b * 9 + i] + RAR[i] + ((i % 4 == 0) ? P.eps_mass : 0.0);
    q[9] = N * invN;
    for (int i = 0; i < 3; ++i) q[10 + i] = mder[4 * B + b * 3 + i] - (R[i * 3] * pb[0] + R[i * 3 + 1] * pb[1] + R[i * 3 + 2] * pb[2]);
    q[13] = sqrt(N * mv.N_pos[b] + P.eps_mass);
    q[14] = pd ? 1.0 : 0.0;
    q[15] = 0.0;
  }
  __syncthreads();
  if (warp == 0) {
    double acc[13] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int b = lane; b < B; b += 32) {
      const double* q = pre + 32 + kPrePitch * (size_t)b;
      const double w = q[13];
      for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) acc[i * 3 + j] += w * q[10 + i] * q[10 + j];
        acc[9 + i] += w * q[10 + i];
      }
      acc[12] += w;
    }
    for (int i = 0; i < 13; ++i) acc[i] = warp_sum(acc[i]);
    if (lane < 9) pre[lane] = R[lane];
    if (lane < 13) pre[9 + lane] = acc[lane];
  }
}

// S4a: IW recurrence alone (the residual statistics do not depend on the IW state when Sigma_meas is not fed back)
__global__ void stream_iw_recurrence_kernel(const __grid_constant__ ChainParams P) {
  const StreamWs W(P.workspace, P.n_scans, P.n_bins, P.n_hyp);
  __shared__ double s_rr[32][10];   // residual statistics of the next 32 scans (one global round trip per 32 scans)
  double iw[10];
  for (int i = 0; i < 10; ++i) iw[i] = P.iw_state[i];
  for (int s0 = 0; s0 < P.n_scans; s0 += 32) {
    const int n = P.n_scans - s0 < 32 ? P.n_scans - s0 : 32;
    for (int i = threadIdx.x; i < n * 10; i += 32) s_rr[i / 10][i % 10] = W.tpart[((size_t)(s0 + i / 10) * P.n_hyp) * 32 + 13 + i % 10];
    __syncwarp();
    for (int s = 0; s < n; ++s) iw_update(iw, s_rr[s], P.iw_rho, P.iw_nu_weak_add, P.iw_nu_max, P.eps_mass, P.eps_psd);  // every lane, same values
    __syncwarp();
  }
  if (threadIdx.x == 0)
    for (int i = 0; i < 10; ++i) P.iw_state[i] = iw[i];
}

// S4b: the sequential core with Sigma_meas fed back.  Warp 0 computes, the other warps stage the next scan's pre block in
//      shared memory.  Per scan:  Sigma_meas -> S = R Sigma_meas R^T -> per bin inv(M_b + c_b S) -> L, h -> t_wls -> residual
//      statistics from the pre-computed sums -> IW update.  Bins whose A_b is not provably >= eps_psd (flag 0: e.g. empty
//      bins) take the full route: PSD projection of A_b + c_b Sigma_meas, R . R^T, map covariance.
__device__ __forceinline__ size_t seq_block_doubles(int B) { return pre_block_doubles(B); }
constexpr int kSeqHelpers = 3;  // warps that stage the next scan's inputs while warp 0 computes
__global__ void __launch_bounds__(32 * (1 + kSeqHelpers)) stream_sigma_chain_kernel(const __grid_constant__ ChainParams P, int use_smem) {
  extern __shared__ __align__(16) double s_stage[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, B = P.n_bins, H = P.n_hyp, S = P.n_scans;
  const StatsLayout SL(B);
  const StreamWs W(P.workspace, S, B, P.n_hyp);
  const size_t blk = seq_block_doubles(B);
  const int hl = tid - 32, HN = 32 * kSeqHelpers;  // helper-thread index / count
  auto stage = [&](int s, double* dst) {  // all helper threads; 16-byte copies
    const double2* src = reinterpret_cast<const double2*>(W.pre + (size_t)s * blk);
    double2* d2 = reinterpret_cast<double2*>(dst);
#pragma unroll 4
    for (int i = hl; i < (int)(blk / 2); i += HN) d2[i] = src[i];
  };
  if (use_smem) {
    if (warp >= 1) stage(0, s_stage);
    __syncthreads();
  }
  // the degrees of freedom nu_s do not depend on the data: a helper thread runs their recurrence one scan ahead (two softplus
  // evaluations off the critical path); the compute warp picks nu_s up after the barrier
  __shared__ double s_nu[2];
  double iw[10];
  for (int i = 0; i < 10; ++i) iw[i] = P.iw_state[i];
  double nu_h = iw[0];
#ifdef GCL_CHAIN_TRACE
  long long tr_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tr_t = clock64();
#define SQ(i) do { const long long n_ = clock64(); tr_acc[i] += n_ - tr_t; tr_t = n_; } while (0)
#else
#define SQ(i) do { } while (0)
#endif
  for (int s = 0; s < S; ++s) {
    if (warp >= 1) {
      if (hl == 0) { nu_h = iw_nu_next(nu_h, P.iw_rho, P.iw_nu_weak_add, P.iw_nu_max); s_nu[(s + 1) & 1] = nu_h; }
      if (use_smem && s + 1 < S) stage(s + 1, s_stage + ((s + 1) & 1) * blk);
    } else {
      if (s > 0) iw[0] = s_nu[s & 1];
      const double* in = use_smem ? s_stage + (s & 1) * blk : W.pre + (size_t)s * blk;
      double sigma[9], R[9], RS[9], Sc[9];
      iw_sigma_mode(iw, P.eps_psd, sigma);
      for (int i = 0; i < 9; ++i) R[i] = in[i];
      mat3_mul(R, sigma, RS);
      mat3_mul_nt(RS, R, Sc);   // S = R Sigma_meas R^T
      SQ(0);
      double Lf[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, hf[3] = {0, 0, 0};
      auto bin_matrix = [&](int b, const double* q, double* Sg) {
        if (q[14] != 0.0) {
          for (int i = 0; i < 9; ++i) Sg[i] = q[i] + q[9] * Sc[i];
        } else {   // full route for this bin
          const double* st0 = P.stats + (size_t)s * H * P.stats_stride;
          const double* md = W.mder + (size_t)s * 16 * B;
          double Sp[9], T1[9], RSR[9];
          sigma_p_with_meas(st0, SL, b, sigma, P.eps_mass, P.eps_psd, Sp);
          mat3_mul(R, Sp, T1);
          mat3_mul_nt(T1, R, RSR);
          for (int i = 0; i < 9; ++i) Sg[i] = md[7 * B + b * 9 + i] + RSR[i] + ((i % 4 == 0) ? P.eps_mass : 0.0);
        }
      };
      // two bins of a lane at a time: their inverses (a division each, the longest dependent chain of the loop) overlap; the
      // sums still take the lane's bins in increasing order
      for (int b0 = lane; b0 < B; b0 += 64) {
        const int b1 = b0 + 32;
        const bool two = b1 < B;
        const double* q0 = in + 32 + kPrePitch * (size_t)b0;
        const double* q1 = in + 32 + kPrePitch * (size_t)(two ? b1 : b0);
        double Sg0[9], Sg1[9], Si0[9], Si1[9];
        bin_matrix(b0, q0, Sg0);
        bin_matrix(two ? b1 : b0, q1, Sg1);
        mat3_inv_adj(Sg0, Si0);
        mat3_inv_adj(Sg1, Si1);
        const double wb0 = q0[13], wb1 = two ? q1[13] : 0.0;
        for (int i = 0; i < 9; ++i) { Si0[i] *= wb0; Lf[i] += Si0[i]; }
        for (int i = 0; i < 3; ++i) hf[i] += Si0[i * 3] * q0[10] + Si0[i * 3 + 1] * q0[11] + Si0[i * 3 + 2] * q0[12];
        if (two) {
          for (int i = 0; i < 9; ++i) { Si1[i] *= wb1; Lf[i] += Si1[i]; }
          for (int i = 0; i < 3; ++i) hf[i] += Si1[i * 3] * q1[10] + Si1[i * 3 + 1] * q1[11] + Si1[i * 3 + 2] * q1[12];
        }
      }
      SQ(1);
      for (int i = 0; i < 9; ++i) Lf[i] = warp_sum(Lf[i]);
      for (int i = 0; i < 3; ++i) hf[i] = warp_sum(hf[i]);
      SQ(2);
      double tw[3], rr_w[10];
      translation_solve(Lf, hf, P.eps_mass, tw);
      const double* Swtt = in + 9; const double* Swt = in + 18; const double Sw = in[21];
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) rr_w[i * 3 + j] = Swtt[i * 3 + j] - Swt[i] * tw[j] - tw[i] * Swt[j] + Sw * tw[i] * tw[j];
      rr_w[9] = Sw;
      SQ(3);
      double* tp = W.tpart + ((size_t)s * H) * 32;
      if (lane < 9) { tp[lane] = Lf[lane]; W.sigma[(size_t)s * 16 + lane] = sigma[lane]; }
      if (lane == 0) { tp[9] = tw[0]; tp[10] = tw[1]; tp[11] = tw[2]; tp[12] = Sw; }
      SQ(4);
      iw_update(iw, rr_w, P.iw_rho, P.iw_nu_weak_add, P.iw_nu_max, P.eps_mass, P.eps_psd, false);  // identical on every lane
      SQ(5);
    }
    __syncthreads();
    if (warp == 0) SQ(6);
  }
#ifdef GCL_CHAIN_TRACE
  if (tid == 0) for (int i = 0; i < 7; ++i) g_chain_trace[20 + i] = tr_acc[i];
#endif
#undef SQ
  if (tid == 0) {
    P.iw_state[0] = s_nu[S & 1];
    for (int i = 1; i < 10; ++i) P.iw_state[i] = iw[i];
  }
}

// S5: finish the translation evidence in parallel: hypothesis 0 from the stored sums, the others from scratch with the
//     stored Sigma_meas of their scan
__global__ void __launch_bounds__(32) stream_translation_finish_kernel(const __grid_constant__ ChainParams P) {
  const int job = blockIdx.x, lane = threadIdx.x, B = P.n_bins, H = P.n_hyp;
  const int s = job / H, h = job - s * H;
  const StatsLayout SL(B);
  const StreamWs W(P.workspace, P.n_scans, B, P.n_hyp);
  double* st = P.stats + (size_t)job * P.stats_stride;
  double* ev = P.evidence + (size_t)job * GCL_EVID_DOUBLES;
  const double* pose = P.pose_pred + (size_t)job * 6;
  const double l1m = ev[GCL_EV_MAP_METRICS + 0], l3m = ev[GCL_EV_MAP_METRICS + 2];
  double Lf[9], tw[3], NeffT;
  const double* sigma = W.sigma + (size_t)s * 16;
  for (int b = lane; b < B; b += 32) {   // the bin covariances with the measurement noise are an output of the chain
    double Sp[9];
    sigma_p_with_meas(st, SL, b, sigma, P.eps_mass, P.eps_psd, Sp);
    for (int i = 0; i < 9; ++i) st[SL.Sigma_p() + b * 9 + i] = Sp[i];
  }
  __syncwarp();
  if (h == 0) {
    const double* tp = W.tpart + ((size_t)s * H) * 32;
    for (int i = 0; i < 9; ++i) Lf[i] = tp[i];
    tw[0] = tp[9]; tw[1] = tp[10]; tw[2] = tp[11]; NeffT = tp[12];
  } else {
    const double* md = W.mder + (size_t)s * 16 * B;
    double R_hat[9], hf[3];
    for (int i = 0; i < 9; ++i) R_hat[i] = ev[GCL_EV_R_MF + i];
    translation_sums_warp(lane, B, R_hat, st + SL.p_bar(), st + SL.Sigma_p(), st + SL.N(), md + 4 * B, md + 7 * B,
                          map_view(P, W, s).N_pos, P.eps_mass, Lf, hf, NeffT);
    translation_solve(Lf, hf, P.eps_mass, tw);
  }
  if (lane == 0) translation_finish(Lf, tw, NeffT, l1m, l3m, pose, P.eps_psd, P.eps_mass, ev);
}

cudaError_t launch_chain(const ChainParams& P, cudaStream_t stream, int* launches) {
  *launches = 0;
  // phase 0: everything.  Phases 1 / 2 split a staged stream for pipelining across batches: 1 = what depends only on the map of
  // the previous batch (increments, map recurrence, map-derived statistics, rotation evidence, pre blocks), 2 = what depends on
  // its IW state (the sequential measurement-noise core and the translation finish).  A short chain runs whole in phase 2.
  if (P.n_scans < kStagedMinScans) {
    if (P.phase == 1) return cudaSuccess;
    if (P.n_scans == 1 && P.n_hyp >= 8 && !P.iw_state && !P.pose_update && !P.stats_src) {
      // a batch of hypotheses against a read-only map: nothing is carried, so the hypotheses go one warp each over the whole
      // GPU instead of two at a time through the single-CTA chain (64 hypotheses: 0.19 ms -> two short launches)
      cudaError_t e = launch_map_derived(P.map_S_dir, P.map_N_dir, P.map_N_pos, P.map_sum_p, P.map_sum_ppT, P.n_bins, P.eps_mass,
                                         P.eps_psd, P.workspace, stream);
      if (e != cudaSuccess) return e;
      EvidenceParams E;
      E.stats = P.stats; E.stats_stride = P.stats_stride;
      E.map_S_dir = P.map_S_dir; E.map_S_dir_scatter = P.map_S_dir_scatter; E.map_N_dir = P.map_N_dir; E.map_N_pos = P.map_N_pos;
      E.map_derived = P.workspace; E.pose_pred = P.pose_pred; E.R_hat_override = nullptr; E.n_bins = P.n_bins;
      E.eps_psd = P.eps_psd; E.eps_mass = P.eps_mass; E.do_rotation = 1; E.do_translation = 1; E.evidence = P.evidence;
      *launches = 2;
      return launch_pose_evidence(E, P.n_hyp, stream);
    }
    ChainParams Q = P;
    const size_t smem = 16 * (size_t)P.n_bins * sizeof(double);
    Q.chain_smem = smem <= 40 * 1024;
    chain_kernel<<<1, kChainThreads, Q.chain_smem ? smem : 0, stream>>>(Q);
    *launches = 1;
    return cudaGetLastError();
  }
  const int B = P.n_bins, S = P.n_scans;
  if (P.phase != 2) {
    if (P.pose_update) {
      stream_increments_kernel<<<S, B < 128 ? ((B + 31) / 32) * 32 : 128, 0, stream>>>(P);
      stream_map_recurrence_kernel<<<(26 * B + 127) / 128, 128, 0, stream>>>(P);
      *launches += 2;
    }
    stream_evidence_kernel<<<S, kChainThreads, 0, stream>>>(P);
    *launches += 1;
  }
  if (P.phase == 1) return cudaGetLastError();
  if (P.iw_state && P.use_sigma_meas) {
    const size_t smem = 2 * pre_block_doubles(B) * sizeof(double);
    const int use_smem = smem <= 200 * 1024;
    static std::atomic<uint64_t> optin{0};
    if (cudaError_t e = optin_dynamic_smem(stream_sigma_chain_kernel, 200 * 1024, optin)) return e;
    stream_sigma_chain_kernel<<<1, 32 * (1 + kSeqHelpers), use_smem ? smem : 0, stream>>>(P, use_smem);
    stream_translation_finish_kernel<<<S * P.n_hyp, 32, 0, stream>>>(P);
    *launches += 2;
  } else if (P.iw_state) {
    stream_iw_recurrence_kernel<<<1, 32, 0, stream>>>(P);
    *launches += 1;
  }
  return cudaGetLastError();
}

}  // namespace gcl
