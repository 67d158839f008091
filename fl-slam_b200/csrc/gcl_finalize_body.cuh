// gcl_finalize_body.cuh -- the body of K2 (fixed-order sum of partial records + per-bin non-linear part) for one job and one
// CTA of kFinThreads threads.  Included by gcl_finalize.cu (reduce_finalize_kernel) and gcl_exchange.cu (the last CTA of
// sharded_reduce_finalize_kernel runs it over the records gathered from the peer GPUs).
#pragma once
#include "gcl_internal.cuh"

namespace gcl {

__device__ __forceinline__ static double block_reduce(double v, bool is_max, double* s_buf) {
  // fixed shuffle tree inside a warp, then warps in index order
  for (int o = 16; o > 0; o >>= 1) {
    const double other = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmax(v, other) : v + other;
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_buf[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = s_buf[0];
  for (int w = 1; w < kFinThreads / 32; ++w) r = is_max ? fmax(r, s_buf[w]) : r + s_buf[w];
  return r;
}

// One job, one CTA of kFinThreads threads: step 1 + (if P.finalize) the per-bin part.  Shared by reduce_finalize_kernel and by the
// last CTA of sharded_reduce_finalize_kernel (gcl_exchange.cu), which calls it with the gathered per-rank records as the parts.
__device__ __forceinline__ static void reduce_finalize_job(const FinalizeParams& P, const int job, FinalizeShared& sh) {
  double* const s_buf = sh.buf;
  double* const s_pad = sh.pad;
  double* const s_redrec = sh.redrec;
  const int tid = threadIdx.x;
  const int B = P.n_bins, NMOM = P.nmom;
  const int rec = NMOM * B + GCL_NCERT;
  const double* src = P.parts + (size_t)job * P.job_stride;
  double* red = P.reduced + (size_t)job * rec;

  // ---- padded rows of the soft-assign certificate (all identical: p = 0, t = 0, w = 0): the LAST warp evaluates one such row
  //      with warp-level reductions while the other warps run step 1 -------------------------------------------------------
  const gcl_job& J = P.jobs ? P.jobs[job] : P.job;
  const int64_t cap = J.n_points_cap > 0 ? J.n_points_cap : 1;
  const int64_t stride0 = (J.n_raw + cap - 1) / cap;
  const int64_t stride = stride0 > 1 ? stride0 : 1;
  const int64_t n_sel = (J.n_raw + stride - 1) / stride;
  const int64_t n_pad = cap > n_sel ? cap - n_sel : 0;
  const bool pad_warp = P.finalize && !P.from_resp && n_pad > 0 && (tid >> 5) == kFinThreads / 32 - 1;
  if (pad_warp) {
    const int lane = tid & 31;
    // deskew of the padded row exactly as the reference evaluates it (epoch-scale alpha)
    const double denomt = fmax(J.scan_end_time - J.scan_start_time, 1e-12);
    const double alpha = (0.0 - J.scan_start_time) / denomt;
    const double pz[3] = {0.0, 0.0, 0.0};
    double p0[3];
    deskew_point64(J.xi_body, alpha, pz, p0);
    const double r0 = p0[0] - P.origin[0], r1 = p0[1] - P.origin[1], r2 = p0[2] - P.origin[2];
    const double inv = 1.0 / (sqrt(r0 * r0 + r1 * r1 + r2 * r2) + P.eps_mass);
    const double d[3] = {r0 * inv, r1 * inv, r2 * inv};
    double mx = -1e300;
    for (int b = lane; b < B; b += 32)
      mx = fmax(mx, (d[0] * P.bin_dirs[b * 3] + d[1] * P.bin_dirs[b * 3 + 1] + d[2] * P.bin_dirs[b * 3 + 2]) / P.tau);
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    double z = 0;
    for (int b = lane; b < B; b += 32)
      z += exp((d[0] * P.bin_dirs[b * 3] + d[1] * P.bin_dirs[b * 3 + 1] + d[2] * P.bin_dirs[b * 3 + 2]) / P.tau - mx);
    for (int o = 16; o > 0; o >>= 1) z += __shfl_xor_sync(0xffffffffu, z, o);
    double h = 0, rm = 0;
    for (int b = lane; b < B; b += 32) {
      const double r = exp((d[0] * P.bin_dirs[b * 3] + d[1] * P.bin_dirs[b * 3 + 1] + d[2] * P.bin_dirs[b * 3 + 2]) / P.tau - mx) / z;
      h -= r * log(r + P.eps_mass);
      rm = fmax(rm, r);
    }
    for (int o = 16; o > 0; o >>= 1) { h += __shfl_xor_sync(0xffffffffu, h, o); rm = fmax(rm, __shfl_xor_sync(0xffffffffu, rm, o)); }
    if (lane == 0) { s_pad[0] = h; s_pad[1] = rm; }
  }

  // ---- step 1: fixed-order sum over parts (the reduced record also stays in shared memory when it fits) -----------------
  const int nred = pad_warp ? 0 : ((P.finalize && !P.from_resp && n_pad > 0) ? kFinThreads - 32 : kFinThreads);
  const bool red_smem = rec <= kRedSmem;
  for (int e = tid; nred > 0 && e < rec; e += nred) {
    const bool is_max = (e == NMOM * B + GCL_CS_MAX_RESP);
    double s = __ldcg(src + e);
    // eight loads in flight per thread (one load per trip pays the L2 latency n_parts times); the order of the additions is unchanged
    for (int p0 = 1; p0 < P.n_parts; p0 += 8) {
      double v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = (p0 + j < P.n_parts) ? __ldcg(src + (size_t)(p0 + j) * P.part_stride + e) : 0.0;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (p0 + j < P.n_parts) s = is_max ? fmax(s, v[j]) : s + v[j];
    }
    red[e] = s;
    if (red_smem) s_redrec[e] = s;
  }
  if (!P.finalize) return;
  __syncthreads();

  const StatsLayout SL(B);
  double* st = P.stats + (size_t)job * SL.total();
  const double eps = P.eps_mass;
  const double* redv = red_smem ? s_redrec : red;   // step 1's result, read back from shared memory when it fits
  const double* cert = redv + (size_t)NMOM * B;

  // ---- scalars of the resample / deskew steps -------------------------------------------------------------------
  const double W_raw = cert[GCL_CS_SUM_W_RAW], W_sel = cert[GCL_CS_SUM_W_SEL];
  const double mass_scale = P.from_resp ? 1.0 : W_raw / (W_sel + eps);  // point_budget.py:80

  // ---- per bin --------------------------------------------------------------------------------------------------
  double a_N = 0, a_N2 = 0, a_sf = 0, a_psd = 0, a_ratio = 0;
  for (int b = tid; b < B; b += kFinThreads) {
    double m[28];
    for (int k = 0; k < NMOM; ++k) m[k] = mass_scale * redv[(size_t)k * B + b];
    const double N = m[0];
    const double denom = N + eps + kF64Eps;  // inv_mass_core, primitives.py:205-212
    const double invN = 1.0 / denom;
    const double ratio = eps / denom;
    const double pb[3] = {m[10] * invN, m[11] * invN, m[12] * invN};
    const double spp[9] = {m[13], m[14], m[15], m[14], m[16], m[17], m[15], m[17], m[18]};
    double Sraw[9];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        double v = spp[i * 3 + j] * invN - pb[i] * pb[j];
        if (NMOM > 19) v += m[19 + i * 3 + j] * invN;  // meas_cov, binning.py:180
        Sraw[i * 3 + j] = v;
      }
    double Spsd[9], pc[6];
    pc[0] = psd_project3_fast(Sraw, P.eps_psd, Spsd);
    const double sn = sqrt(m[1] * m[1] + m[2] * m[2] + m[3] * m[3]);
    const double kap = kappa_from_resultant(sn * invN, 1e-6, 3.0, 0.8, 0.03);  // GC_EPS_R, blend r0/tau: constants.py:73,:95-96
    st[SL.N() + b] = N;
    for (int i = 0; i < 3; ++i) st[SL.s_dir() + b * 3 + i] = m[1 + i];
    const double S9[9] = {m[4], m[5], m[6], m[5], m[7], m[8], m[6], m[8], m[9]};
    for (int i = 0; i < 9; ++i) st[SL.S_scatter() + b * 9 + i] = S9[i];
    for (int i = 0; i < 3; ++i) st[SL.p_bar() + b * 3 + i] = pb[i];
    for (int i = 0; i < 9; ++i) st[SL.Sigma_p() + b * 9 + i] = Spsd[i];
    st[SL.kappa() + b] = kap;
    for (int i = 0; i < 3; ++i) st[SL.sum_p() + b * 3 + i] = m[10 + i];
    for (int i = 0; i < 9; ++i) st[SL.sum_ppT() + b * 9 + i] = spp[i];
    a_N += N; a_N2 += N * N; a_sf += N / (N + eps); a_psd += pc[0];
    a_ratio = fmax(a_ratio, ratio);
  }
  const double tN = block_reduce(a_N, false, s_buf);
  const double tN2 = block_reduce(a_N2, false, s_buf);
  const double tsf = block_reduce(a_sf, false, s_buf);
  const double tpsd = block_reduce(a_psd, false, s_buf);
  const double tratio = block_reduce(a_ratio, true, s_buf);

  // ---- soft-assign certificate: add the padded rows (evaluated by the last warp above; visible after block_reduce's barriers)
  double ent_total = cert[GCL_CS_SUM_ENTROPY], max_resp = cert[GCL_CS_MAX_RESP];
  if (!P.from_resp && n_pad > 0) {
    ent_total += (double)n_pad * s_pad[0];
    max_resp = fmax(max_resp, s_pad[1]);
  }

  if (tid == 0) {
    double* sc = st + SL.scalars();
    for (int i = 0; i < GCL_STATS_SCALARS; ++i) sc[i] = 0.0;
    const double M = W_raw;
    sc[GCL_SS_TOTAL_MASS_IN] = W_raw;
    sc[GCL_SS_TOTAL_MASS_SEL] = W_sel;
    sc[GCL_SS_MASS_SCALE] = mass_scale;
    const double wn2 = mass_scale * mass_scale * cert[GCL_CS_SUM_W_SEL_SQ] / ((M + eps) * (M + eps));
    sc[GCL_SS_PB_ESS] = 1.0 / (wn2 + (double)cap * eps);  // eps inside the sum over all cap rows, point_budget.py:95-96
    sc[GCL_SS_N_SELECTED] = (double)n_sel;
    sc[GCL_SS_STRIDE] = (double)stride;
    const double w_out = mass_scale * cert[GCL_CS_SUM_W_DESKEW];
    sc[GCL_SS_SUM_W_OUT] = w_out;
    sc[GCL_SS_DESKEW_RETAINED] = w_out / (mass_scale * W_sel + eps);  // deskew_constant_twist.py:102-103
    sc[GCL_SS_AVG_ENTROPY] = ent_total / ((double)cap + eps);         // binning.py:75 with n_points = cap rows
    sc[GCL_SS_MAX_RESP] = max_resp;
    sc[GCL_SS_ENTROPY_VALID_ROWS] = cert[GCL_CS_SUM_ENTROPY] / (cert[GCL_CS_ROWS] + eps);
    sc[GCL_SS_MAX_RESP_VALID_ROWS] = cert[GCL_CS_MAX_RESP];
    sc[GCL_SS_MM_ESS] = tN * tN / (tN2 + eps);                        // binning.py:201-202
    sc[GCL_SS_MM_SUPPORT_FRAC] = tsf / (double)B;
    sc[GCL_SS_MM_PSD_DELTA] = tpsd;
    sc[GCL_SS_MM_MAX_EPS_RATIO] = tratio;
  }
}

}  // namespace gcl
