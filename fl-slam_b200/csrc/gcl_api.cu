// gcl_api.cu -- the C ABI declared in include/gcl_b200.h: argument checks, launch planning, error text.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "gcl_internal.cuh"

using namespace gcl;

static std::string g_create_error;

// The submit thread and the caller's thread both run argument checks and CUDA calls on one handle: each writes its OWN error
// string (the thread's text reaches the caller through Slot::task_err at the next call on the slot).
static thread_local bool t_submit_thread = false;
static inline std::string& err_of(gcl_handle h) { return t_submit_thread ? h->err_submit : h->err; }

#define GCL_REQUIRE(h, cond, msg)                  \
  do {                                             \
    if (!(cond)) {                                 \
      err_of(h) = std::string(__func__) + ": " + (msg); \
      return GCL_ERR_ARG;                          \
    }                                              \
  } while (0)

static int cuda_fail(gcl_handle h, const char* fn, cudaError_t e) {
  err_of(h) = std::string(fn) + ": CUDA error: " + cudaGetErrorString(e);
  return GCL_ERR_CUDA;
}
#define GCL_CUDA(h, expr)                                   \
  do {                                                      \
    cudaError_t e__ = (expr);                               \
    if (e__ != cudaSuccess) return cuda_fail(h, __func__, e__); \
  } while (0)

static bool fixed_shift_ok(double tau, int precision) {
  // exp arguments lie in [-2/tau, 0] with the shift 1/tau; keep them far from fp32 / fp64 underflow
  return (2.0 / tau) <= (precision == GCL_F64 ? 600.0 : 60.0);
}

extern "C" {

int gcl_abi_version(void) { return GCL_ABI_VERSION; }

int gcl_create(int device, gcl_handle* out) {
  if (!out) { g_create_error = "gcl_create: out is NULL"; return GCL_ERR_ARG; }
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    g_create_error = std::string("gcl_create: no CUDA device: ") + cudaGetErrorString(e);
    return GCL_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= count) { g_create_error = "gcl_create: device index out of range"; return GCL_ERR_ARG; }
  e = cudaSetDevice(device);
  if (e != cudaSuccess) { g_create_error = std::string("gcl_create: cudaSetDevice: ") + cudaGetErrorString(e); return GCL_ERR_CUDA; }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) { g_create_error = std::string("gcl_create: ") + cudaGetErrorString(e); return GCL_ERR_CUDA; }
  if (prop.major < 10) {
    g_create_error = "gcl_create: this library is built for sm_100a only (found compute capability " +
                     std::to_string(prop.major) + "." + std::to_string(prop.minor) + ")";
    return GCL_ERR_UNSUPPORTED;
  }
  gcl_handle h = new gcl_handle_s();
  h->device = device;
  h->sm_count = prop.multiProcessorCount;
  *out = h;
  return GCL_OK;
}

void gcl_destroy(gcl_handle h) {
  if (!h) return;
  if (h->worker.joinable()) {
    { std::lock_guard<std::mutex> lk(h->mu); h->stop = true; }
    h->cv_task.notify_all();
    h->worker.join();
  }
  cudaSetDevice(h->device);
  if (h->stream) { cudaStreamSynchronize(h->stream); cudaStreamDestroy(h->stream); }
  if (h->stream_copy) { cudaStreamSynchronize(h->stream_copy); cudaStreamDestroy(h->stream_copy); }
  for (int i = 0; i < GCL_HOST_SLOTS; ++i) {
    if (h->slot[i].stream1) { cudaStreamSynchronize(h->slot[i].stream1); cudaStreamDestroy(h->slot[i].stream1); }
    if (h->slot[i].ev_copy) cudaEventDestroy(h->slot[i].ev_copy);
    if (h->slot[i].ev_stage1) cudaEventDestroy(h->slot[i].ev_stage1);
    if (h->slot[i].d_buf) cudaFree(h->slot[i].d_buf);
    if (h->slot[i].h_buf) cudaFreeHost(h->slot[i].h_buf);
  }
  if (h->d_buf) cudaFree(h->d_buf);
  if (h->h_buf) cudaFreeHost(h->h_buf);
  if (h->done_flag) cudaFreeHost(h->done_flag);
  delete h;
}

const char* gcl_last_error(gcl_handle h) { return h ? h->err.c_str() : g_create_error.c_str(); }
int gcl_device_sm_count(gcl_handle h) { return h ? h->sm_count : 0; }
int64_t gcl_launch_count(gcl_handle h) { return h ? h->launches.load() : (int64_t)0; }

size_t gcl_partial_record_doubles(int n_bins) { return (size_t)GCL_NM * n_bins + GCL_NCERT; }
size_t gcl_stats_record_doubles(int n_bins) { return (size_t)StatsLayout(n_bins).total(); }

int gcl_plan_ctas_per_job(gcl_handle h, int n_jobs, int64_t max_sel_rows, int n_bins, int precision) {
  // resident CTA slots: SMs x CTAs per SM of the kernel variant this (n_bins, precision) selects
  const int64_t sms = (int64_t)(h ? h->sm_count : 148) * fused_ctas_per_sm(n_bins < 1 ? 1 : n_bins, precision);
  if (n_jobs < 1) n_jobs = 1;
  if (max_sel_rows < 1) max_sel_rows = 1;
  // at least one 256-row tile of work per CTA; when that would leave SMs idle (one small scan: the latency case),
  // go down to 96 rows per CTA: shorter CTAs, but not so many partial records that reduce_finalize pays for them
  // (measured at 7500 rows x 48 bins: 118 / 74 / 37 CTAs -> scan_moments 13.1 / 14.4 / 17.0 us, reduce_finalize 22.7 / 18.7 / 14.7 us)
  int64_t by_rows = (max_sel_rows + 255) / 256;
  if ((int64_t)n_jobs * by_rows < sms) by_rows = (max_sel_rows + 95) / 96;
  // one job: one CTA per SM (equal slices, no tail)
  if (n_jobs == 1) return (int)(sms < by_rows ? sms : by_rows);
  // several jobs: the grid runs in waves of `sms` CTAs and the last wave is rarely full (200 jobs x 3 CTAs = 4.05 waves cost
  // 5).  Pick the CTAs per job that minimise  waves x (rows per CTA + fixed cost per CTA), the fixed cost (record zeroing,
  // atlas set-up, flush, certificate reduction: ~6 us) being worth about 128 rows of a 48-bin fp64 pass.
  int64_t cmax = by_rows < 64 ? by_rows : 64, best_c = 1;
  double best = 1e300;
  for (int64_t c = 1; c <= cmax; ++c) {
    const int64_t waves = (n_jobs * c + sms - 1) / sms;
    const double cost = (double)waves * ((double)((max_sel_rows + c - 1) / c) + 128.0);
    if (cost < best) { best = cost; best_c = c; }
  }
  return (int)best_c;
}

int gcl_scan_moments(gcl_handle h, void* stream, const gcl_job* jobs_dev, const gcl_job* job_host, int n_jobs,
                     int ctas_per_job, int layout, int precision, const double* bin_dirs, int n_bins, double tau,
                     const double* origin_host, double eps_mass, double* partials) {
  if (!h) return GCL_ERR_ARG;
  GCL_REQUIRE(h, n_jobs >= 1, "n_jobs must be >= 1");
  GCL_REQUIRE(h, (jobs_dev != nullptr) != (job_host != nullptr), "exactly one of jobs_dev / job_by_value_host must be given");
  GCL_REQUIRE(h, jobs_dev || n_jobs == 1, "a by-value job implies n_jobs == 1");
  GCL_REQUIRE(h, n_bins >= 1 && n_bins <= GCL_MAX_BINS, "n_bins must be in [1, 1024]");
  GCL_REQUIRE(h, tau > 0.0, "tau must be > 0 (the reference ships no default, SURVEY.md 0.4)");
  GCL_REQUIRE(h, precision == GCL_F64 || precision == GCL_F32C, "unknown precision mode");
  GCL_REQUIRE(h, layout == GCL_LAYOUT_REF_F64 || layout == GCL_LAYOUT_PACKED_F32, "unknown layout");
  GCL_REQUIRE(h, !(precision == GCL_F64 && layout == GCL_LAYOUT_PACKED_F32), "fp64 mode needs the fp64 reference layout");
  GCL_REQUIRE(h, ctas_per_job >= 1 && ctas_per_job <= 65535, "ctas_per_job out of range");
  GCL_REQUIRE(h, bin_dirs && origin_host && partials, "NULL pointer");
  if (job_host) {
    GCL_REQUIRE(h, job_host->n_raw >= 0 && job_host->n_points_cap >= 1, "job: n_raw >= 0 and n_points_cap >= 1 required");
    GCL_REQUIRE(h, job_host->n_raw == 0 || (job_host->points && job_host->weights), "job: NULL points / weights");
  }
  ScanParams P;
  memset(&P, 0, sizeof(P));
  P.jobs = jobs_dev;
  if (job_host) P.job = *job_host;
  P.c.origin[0] = origin_host[0]; P.c.origin[1] = origin_host[1]; P.c.origin[2] = origin_host[2];
  P.c.eps_mass = eps_mass;
  P.bin_dirs = bin_dirs;
  P.partials = partials;
  P.rec_doubles = (int64_t)gcl_partial_record_doubles(n_bins);
  P.tau = tau;
  P.n_bins = n_bins;
  P.layout = layout;
  P.fixed_shift = fixed_shift_ok(tau, precision) ? 1 : 0;
  cudaSetDevice(h->device);
  GCL_CUDA(h, launch_scan_moments_fused(P, n_jobs, ctas_per_job, precision, (cudaStream_t)stream));
  h->launches += 1;
  return GCL_OK;
}

int gcl_reduce_finalize(gcl_handle h, void* stream, const double* parts, int n_parts, int64_t part_stride,
                        int64_t job_stride, const gcl_job* jobs_dev, const gcl_job* job_host, int n_jobs, int n_bins,
                        const double* bin_dirs, double tau, const double* origin_host, double eps_psd, double eps_mass,
                        int finalize, double* reduced, double* stats) {
  if (!h) return GCL_ERR_ARG;
  GCL_REQUIRE(h, n_jobs >= 1 && n_parts >= 1, "n_jobs and n_parts must be >= 1");
  GCL_REQUIRE(h, n_bins >= 1 && n_bins <= GCL_MAX_BINS, "n_bins must be in [1, 1024]");
  GCL_REQUIRE(h, parts && reduced, "parts and reduced must be given");
  GCL_REQUIRE(h, !finalize || (stats && bin_dirs && origin_host && tau > 0.0), "finalize needs stats, bin_dirs, origin and tau > 0");
  GCL_REQUIRE(h, !finalize || ((jobs_dev != nullptr) != (job_host != nullptr)), "finalize needs the jobs (device array or one by value)");
  FinalizeParams P;
  memset(&P, 0, sizeof(P));
  P.parts = parts; P.n_parts = n_parts; P.part_stride = part_stride; P.job_stride = job_stride;
  P.jobs = jobs_dev;
  if (job_host) P.job = *job_host;
  P.n_bins = n_bins; P.nmom = GCL_NM; P.bin_dirs = bin_dirs; P.tau = tau;
  if (origin_host) { P.origin[0] = origin_host[0]; P.origin[1] = origin_host[1]; P.origin[2] = origin_host[2]; }
  P.eps_psd = eps_psd; P.eps_mass = eps_mass; P.finalize = finalize; P.from_resp = 0;
  P.reduced = reduced; P.stats = stats;
  cudaSetDevice(h->device);
  int l = 0;
  GCL_CUDA(h, launch_reduce_finalize(P, n_jobs, (cudaStream_t)stream, &l));
  h->launches += l - 1;
  h->launches += 1;
  return GCL_OK;
}

int gcl_map_derived_stats(gcl_handle h, void* stream, const double* S_dir, const double* N_dir, const double* N_pos,
                          const double* sum_p, const double* sum_ppT, int n_bins, double eps_mass, double eps_psd, double* out) {
  if (!h) return GCL_ERR_ARG;
  GCL_REQUIRE(h, n_bins >= 1, "n_bins must be >= 1");
  GCL_REQUIRE(h, S_dir && N_dir && N_pos && sum_p && sum_ppT && out, "NULL pointer");
  cudaSetDevice(h->device);
  GCL_CUDA(h, launch_map_derived(S_dir, N_dir, N_pos, sum_p, sum_ppT, n_bins, eps_mass, eps_psd, out, (cudaStream_t)stream));
  h->launches += 1;
  return GCL_OK;
}

int gcl_pose_evidence(gcl_handle h, void* stream, const double* stats, int64_t stats_stride, const double* map_S_dir,
                      const double* map_S_dir_scatter, const double* map_N_dir, const double* map_N_pos,
                      const double* map_derived, const double* pose_pred, const double* R_hat_override, int n_jobs, int n_bins,
                      double eps_psd, double eps_mass, int do_rotation, int do_translation, double* evidence) {
  if (!h) return GCL_ERR_ARG;
  GCL_REQUIRE(h, n_jobs >= 1 && n_bins >= 1, "n_jobs and n_bins must be >= 1");
  GCL_REQUIRE(h, stats && map_S_dir && map_S_dir_scatter && map_N_dir && pose_pred && evidence, "NULL pointer");
  GCL_REQUIRE(h, !do_translation || (map_N_pos && map_derived), "translation needs map_N_pos and map_derived");
  GCL_REQUIRE(h, do_rotation || !do_translation || R_hat_override, "translation without rotation needs R_hat_override");
  EvidenceParams P;
  P.stats = stats; P.stats_stride = stats_stride;
  P.map_S_dir = map_S_dir; P.map_S_dir_scatter = map_S_dir_scatter; P.map_N_dir = map_N_dir; P.map_N_pos = map_N_pos;
  P.map_derived = map_derived; P.pose_pred = pose_pred; P.R_hat_override = R_hat_override; P.n_bins = n_bins;
  P.eps_psd = eps_psd; P.eps_mass = eps_mass; P.do_rotation = do_rotation; P.do_translation = do_translation;
  P.evidence = evidence;
  cudaSetDevice(h->device);
  GCL_CUDA(h, launch_pose_evidence(P, n_jobs, (cudaStream_t)stream));
  h->launches += 1;
  return GCL_OK;
}

int gcl_point_budget_resample(gcl_handle h, void* stream, const double* points, const double* timestamps, const double* weights,
                              const uint8_t* ring, const uint8_t* tag, int64_t n_input, int32_t cap, double eps_mass,
                              double* points_out, double* timestamps_out, double* weights_out, uint8_t* ring_out,
                              uint8_t* tag_out, double* scalars_out, double* workspace) {
  if (!h) return GCL_ERR_ARG;
  GCL_REQUIRE(h, n_input >= 0 && cap >= 1, "n_input >= 0 and n_points_cap >= 1 required");
  GCL_REQUIRE(h, n_input == 0 || (points && timestamps && weights), "NULL input");
  GCL_REQUIRE(h, points_out && timestamps_out && weights_out && ring_out && tag_out && scalars_out && workspace, "NULL output");
  int l = 0;
  cudaSetDevice(h->device);
  GCL_CUDA(h, op_point_budget(points, timestamps, weights, ring, tag, n_input, cap, eps_mass, points_out, timestamps_out,
                              weights_out, ring_out, tag_out, scalars_out, workspace, (cudaStream_t)stream, &l));
  h->launches += l;
  return GCL_OK;
}

int gcl_deskew_constant_twist(gcl_handle h, void* stream, const double* points, const double* timestamps, const double* weights,
                              int64_t n, double t0, double t1, const double* xi_host, double eps_mass, double* points_out,
                              double* weights_out, double* scalars_out, double* workspace) {
  if (!h) return GCL_ERR_ARG;
  GCL_REQUIRE(h, n >= 0 && xi_host, "n >= 0 and xi_body required");
  GCL_REQUIRE(h, n == 0 || (points && timestamps && weights && points_out && weights_out), "NULL array");
  GCL_REQUIRE(h, scalars_out && workspace, "NULL scalars / workspace");
  int l = 0;
  cudaSetDevice(h->device);
  GCL_CUDA(h, op_deskew(points, timestamps, weights, n, t0, t1, xi_host, eps_mass, points_out, weights_out, scalars_out,
                        workspace, (cudaStream_t)stream, &l));
  h->launches += l;
  return GCL_OK;
}

int gcl_ray_directions(gcl_handle h, void* stream, const double* points, int64_t n, const double* origin_host, double eps_mass,
                       double* dirs_out) {
  if (!h) return GCL_ERR_ARG;
  GCL_REQUIRE(h, n >= 0 && origin_host && (n == 0 || (points && dirs_out)), "bad arguments");
  int l = 0;
  cudaSetDevice(h->device);
  GCL_CUDA(h, op_ray_dirs(points, n, origin_host, eps_mass, dirs_out, (cudaStream_t)stream, &l));
  h->launches += l;
  return GCL_OK;
}

int gcl_bin_soft_assign(gcl_handle h, void* stream, const double* point_dirs, int64_t n, const double* bin_dirs, int n_bins,
                        double tau, double eps_mass, double* responsibilities, double* scalars_out, double* workspace) {
  if (!h) return GCL_ERR_ARG;
  GCL_REQUIRE(h, n >= 0 && n_bins >= 1 && n_bins <= 4096, "n >= 0 and 1 <= n_bins <= 4096 required");
  GCL_REQUIRE(h, tau > 0.0, "tau must be > 0");
  GCL_REQUIRE(h, bin_dirs && scalars_out && workspace && (n == 0 || (point_dirs && responsibilities)), "NULL pointer");
  int l = 0;
  cudaSetDevice(h->device);
  GCL_CUDA(h, op_soft_assign(point_dirs, n, bin_dirs, n_bins, tau, eps_mass, responsibilities, scalars_out, workspace,
                             (cudaStream_t)stream, &l));
  h->launches += l;
  return GCL_OK;
}

static int mm_ctas(gcl_handle h, int64_t n) {
  int64_t c = (n + 255) / 256;
  if (c < 1) c = 1;
  if (c > 2 * h->sm_count) c = 2 * h->sm_count;
  return (int)c;
}

size_t gcl_moment_match_workspace_doubles(gcl_handle h, int64_t n, int n_bins) {
  if (!h) return 0;
  const size_t rec = (size_t)28 * n_bins + GCL_NCERT;
  return rec * (size_t)mm_ctas(h, n) + rec;
}

int gcl_scan_bin_moment_match(gcl_handle h, void* stream, const double* points, const double* point_cov, const double* weights,
                              const double* resp, const double* point_lambda, const double* origin_host, int64_t n, int n_bins,
                              double eps_psd, double eps_mass, double* stats, double* workspace) {
  if (!h) return GCL_ERR_ARG;
  GCL_REQUIRE(h, n >= 0 && n_bins >= 1 && n_bins <= GCL_MAX_BINS, "n >= 0 and 1 <= n_bins <= 1024 required");
  GCL_REQUIRE(h, n < (int64_t)1 << 31, "n too large for the materialised-responsibility operator");
  GCL_REQUIRE(h, origin_host && stats && workspace && (n == 0 || (points && weights && resp)), "NULL pointer");
  const int ctas = mm_ctas(h, n);
  const int64_t rec = (int64_t)28 * n_bins + GCL_NCERT;
  ScanParams P;
  memset(&P, 0, sizeof(P));
  P.job.points = points; P.job.weights = weights; P.job.point_lambda = point_lambda;
  P.job.n_raw = n; P.job.n_points_cap = (int32_t)(n > 0 ? n : 1); P.job.sel_begin = 0; P.job.sel_end = -1;
  P.c.origin[0] = origin_host[0]; P.c.origin[1] = origin_host[1]; P.c.origin[2] = origin_host[2];
  P.c.eps_mass = eps_mass;
  P.resp = resp; P.point_cov = point_cov;
  P.partials = workspace;
  P.rec_doubles = rec;
  P.tau = 1.0; P.n_bins = n_bins;
  int G, L, chunks;
  plan_bins(n_bins, GCL_F64, &G, &L, &chunks);
  P.lanes = L; P.n_chunks = chunks; P.layout = GCL_LAYOUT_REF_F64; P.fixed_shift = 1;
  cudaSetDevice(h->device);
  GCL_CUDA(h, launch_moments_from_resp(P, ctas, G, (cudaStream_t)stream));
  FinalizeParams F;
  memset(&F, 0, sizeof(F));
  F.parts = workspace; F.n_parts = ctas; F.part_stride = rec; F.job_stride = 0;
  F.job = P.job; F.n_bins = n_bins; F.nmom = 28; F.tau = 1.0;
  F.eps_psd = eps_psd; F.eps_mass = eps_mass; F.finalize = 1; F.from_resp = 1;
  F.reduced = workspace + (size_t)rec * ctas; F.stats = stats;
  int lf = 0;
  GCL_CUDA(h, launch_reduce_finalize(F, 1, (cudaStream_t)stream, &lf));
  h->launches += lf - 1;
  h->launches += 2;
  return GCL_OK;
}

int gcl_kappa_from_resultant(gcl_handle h, void* stream, const double* R_bar, int64_t n, double eps_r, int d, double r0,
                             double tau, double* kappa_out) {
  if (!h) return GCL_ERR_ARG;
  GCL_REQUIRE(h, n >= 0 && (n == 0 || (R_bar && kappa_out)), "bad arguments");
  int l = 0;
  cudaSetDevice(h->device);
  GCL_CUDA(h, op_kappa(R_bar, n, eps_r, d, r0, tau, kappa_out, (cudaStream_t)stream, &l));
  h->launches += l;
  return GCL_OK;
}

int gcl_fibonacci_atlas(gcl_handle h, void* stream, int n_bins, double eps_mass, double* dirs_out) {
  if (!h) return GCL_ERR_ARG;
  GCL_REQUIRE(h, n_bins >= 1 && dirs_out, "n_bins >= 1 and an output buffer required");
  int l = 0;
  cudaSetDevice(h->device);
  GCL_CUDA(h, op_atlas(n_bins, eps_mass, dirs_out, (cudaStream_t)stream, &l));
  h->launches += l;
  return GCL_OK;
}

int gcl_map_update(gcl_handle h, void* stream, const double* stats, const double* R_world, const double* t_world, int n_bins,
                   double forgetting, double* S_dir, double* S_dir_scatter, double* N_dir, double* N_pos, double* sum_p,
                   double* sum_ppT) {
  if (!h) return GCL_ERR_ARG;
  GCL_REQUIRE(h, n_bins >= 1 && stats && R_world && t_world && S_dir && S_dir_scatter && N_dir && N_pos && sum_p && sum_ppT,
              "bad arguments");
  GCL_REQUIRE(h, forgetting > 0.0 && forgetting <= 1.0, "forgetting factor must be in (0, 1]");
  cudaSetDevice(h->device);
  GCL_CUDA(h, launch_map_update(stats, R_world, t_world, n_bins, forgetting, S_dir, S_dir_scatter, N_dir, N_pos, sum_p,
                                sum_ppT, (cudaStream_t)stream));
  h->launches += 1;
  return GCL_OK;
}

size_t gcl_scan_chain_workspace_doubles(int n_scans, int n_hyp, int n_bins) {
  return chain_workspace_doubles(n_scans < 1 ? 1 : n_scans, n_hyp < 1 ? 1 : n_hyp, n_bins < 1 ? 1 : n_bins);
}

int gcl_scan_chain(gcl_handle h, void* stream, double* stats, int64_t stats_stride, int n_scans, int n_hyp, int n_bins,
                   double* mS_dir, double* mS_sc, double* mN_dir, double* mN_pos, double* msum_p, double* msum_ppT,
                   double* map_derived, const double* pose_pred, const double* pose_update, double forgetting, double* iw_state,
                   int use_sigma_meas, double iw_rho, double iw_nu_weak_add, double iw_nu_max, double eps_psd, double eps_mass,
                   double* evidence) {
  return gcl_scan_chain_phase(h, stream, 0, stats, stats_stride, n_scans, n_hyp, n_bins, mS_dir, mS_sc, mN_dir, mN_pos, msum_p, msum_ppT,
                              map_derived, pose_pred, pose_update, forgetting, iw_state, use_sigma_meas, iw_rho, iw_nu_weak_add, iw_nu_max,
                              eps_psd, eps_mass, evidence);
}

int gcl_scan_chain_phase(gcl_handle h, void* stream, int phase, double* stats, int64_t stats_stride, int n_scans, int n_hyp, int n_bins,
                         double* mS_dir, double* mS_sc, double* mN_dir, double* mN_pos, double* msum_p, double* msum_ppT,
                         double* map_derived, const double* pose_pred, const double* pose_update, double forgetting, double* iw_state,
                         int use_sigma_meas, double iw_rho, double iw_nu_weak_add, double iw_nu_max, double eps_psd, double eps_mass,
                         double* evidence) {
  if (!h) return GCL_ERR_ARG;
  GCL_REQUIRE(h, phase >= 0 && phase <= 2, "phase must be 0 (all), 1 (map-dependent part) or 2 (IW-dependent part)");
  GCL_REQUIRE(h, n_scans >= 1 && n_hyp >= 1 && n_bins >= 1, "n_scans, n_hyp, n_bins must be >= 1");
  GCL_REQUIRE(h, stats && mS_dir && mS_sc && mN_dir && mN_pos && msum_p && msum_ppT && map_derived && pose_pred && evidence,
              "NULL pointer");
  GCL_REQUIRE(h, !pose_update || (forgetting > 0.0 && forgetting <= 1.0), "forgetting factor must be in (0, 1]");
  GCL_REQUIRE(h, !use_sigma_meas || iw_state, "use_sigma_meas needs iw_state");
  ChainParams P;
  P.stats = stats; P.stats_stride = stats_stride;
  P.map_S_dir = mS_dir; P.map_S_dir_scatter = mS_sc; P.map_N_dir = mN_dir; P.map_N_pos = mN_pos; P.map_sum_p = msum_p;
  P.map_sum_ppT = msum_ppT; P.workspace = map_derived; P.pose_pred = pose_pred; P.pose_update = pose_update;
  P.iw_state = iw_state; P.evidence = evidence; P.n_scans = n_scans; P.n_hyp = n_hyp; P.n_bins = n_bins;
  P.use_sigma_meas = use_sigma_meas; P.forgetting = forgetting; P.iw_rho = iw_rho; P.iw_nu_weak_add = iw_nu_weak_add;
  P.iw_nu_max = iw_nu_max; P.eps_psd = eps_psd; P.eps_mass = eps_mass; P.chain_smem = 0; P.phase = phase;
  cudaSetDevice(h->device);
  int l = 0;
  GCL_CUDA(h, launch_chain(P, (cudaStream_t)stream, &l));
  h->launches += l;
  return GCL_OK;
}

int gcl_pointcloud2_ingest(gcl_handle h, void* stream, const uint8_t* data, int64_t n_points, int32_t point_step,
                           const int32_t* off, const int32_t* dt, double header_stamp, const double* R_host, const double* t_host,
                           double* points_out, double* timestamps_out, double* weights_out, uint8_t* ring_out, uint8_t* tag_out,
                           void* packed_xyzt, float* packed_w, double packed_t0, double* workspace) {
  if (!h) return GCL_ERR_ARG;
  GCL_REQUIRE(h, n_points >= 0 && point_step >= 1 && point_step <= 512, "n_points >= 0 and 1 <= point_step <= 512 required");
  GCL_REQUIRE(h, off && dt && workspace && (n_points == 0 || data), "NULL pointer");
  GCL_REQUIRE(h, (R_host == nullptr) == (t_host == nullptr), "R_base_lidar and t_base_lidar go together");
  GCL_REQUIRE(h, (packed_xyzt == nullptr) == (packed_w == nullptr), "packed_xyzt and packed_w go together");
  static const int width[9] = {0, 1, 1, 2, 2, 4, 4, 4, 8};
  IngestParams P;
  memset(&P, 0, sizeof(P));
  for (int k = 0; k < 5; ++k) {
    const bool optional = (k == 4) && off[k] < 0;
    GCL_REQUIRE(h, optional || (dt[k] >= 1 && dt[k] <= 8), "Unsupported PointField datatype");   // backend_node.py:374
    GCL_REQUIRE(h, optional || (off[k] >= 0 && off[k] + width[dt[k]] <= point_step), "field does not fit in point_step");
    P.off[k] = off[k]; P.dtype[k] = optional ? 0 : dt[k];
  }
  P.data = data; P.n_points = n_points; P.point_step = point_step; P.header_stamp = header_stamp;
  P.sentinel = 1e6; P.sigma = 0.25; P.min_r = 0.5; P.max_r = 50.0; P.weight_floor = 1e-12;   // constants.py:256-262
  P.transform = R_host != nullptr;
  if (R_host) { for (int k = 0; k < 9; ++k) P.R[k] = R_host[k]; for (int k = 0; k < 3; ++k) P.t[k] = t_host[k]; }
  P.points_out = points_out; P.timestamps_out = timestamps_out; P.weights_out = weights_out; P.ring_out = ring_out; P.tag_out = tag_out;
  P.packed_xyzt = static_cast<float4*>(packed_xyzt); P.packed_w = packed_w; P.packed_t0 = packed_t0;
  P.flag = reinterpret_cast<int*>(workspace);
  int l = 0;
  cudaSetDevice(h->device);
  GCL_CUDA(h, launch_ingest(P, (cudaStream_t)stream, &l));
  h->launches += l;
  return GCL_OK;
}

int gcl_imu_preintegrate(gcl_handle h, void* stream, const double* stamps, const double* gyro, const double* accel,
                         const double* weights, int n_samples, double t0, double t1, double sigma, const double* rotvec_start,
                         const double* gyro_bias, const double* accel_bias, const double* gravity_host, int rotation_only,
                         int n_hyp, double* out, double* weights_out, gcl_job* jobs_out) {
  if (!h) return GCL_ERR_ARG;
  GCL_REQUIRE(h, n_samples >= 1 && n_hyp >= 1, "n_samples and n_hyp must be >= 1");
  GCL_REQUIRE(h, stamps && gyro && accel && rotvec_start && gyro_bias && accel_bias && gravity_host && out, "NULL pointer");
  ImuParams P;
  P.stamps = stamps; P.gyro = gyro; P.accel = accel; P.weights = weights; P.M = n_samples;
  P.t0 = t0; P.t1 = t1; P.sigma = sigma; P.bounds_dev = nullptr;
  P.rotvec_start = rotvec_start; P.gyro_bias = gyro_bias; P.accel_bias = accel_bias;
  P.gravity[0] = gravity_host[0]; P.gravity[1] = gravity_host[1]; P.gravity[2] = gravity_host[2];
  P.trans_scale = rotation_only ? 0.0 : 1.0;
  P.out = out; P.weights_out = weights_out; P.jobs_out = jobs_out;
  cudaSetDevice(h->device);
  GCL_CUDA(h, launch_imu_preintegrate(P, n_hyp, (cudaStream_t)stream));
  h->launches += 1;
  return GCL_OK;
}

int gcl_pack_points_f32(gcl_handle h, void* stream, const double* points, const double* timestamps, const double* weights,
                        int64_t n, double t0, void* packed_xyzt, float* packed_w) {
  if (!h) return GCL_ERR_ARG;
  GCL_REQUIRE(h, n >= 0 && (n == 0 || (points && timestamps && weights && packed_xyzt && packed_w)), "bad arguments");
  int l = 0;
  cudaSetDevice(h->device);
  GCL_CUDA(h, op_pack(points, timestamps, weights, n, t0, packed_xyzt, packed_w, (cudaStream_t)stream, &l));
  h->launches += l;
  return GCL_OK;
}

// ---- reference-facing calls with HOST buffers ---------------------------------------------------------------------------------
// Two stages per scan.  Stage 1 (submit) needs only the scan and the twist: H2D of the scan on the copy stream, then
// scan_moments + reduce_finalize on the stage-1 stream; it returns without waiting.  Stage 2 (evidence) needs the pose
// prediction and the map: it waits (on the GPU) for stage 1 of the same slot, runs the chain step, copies every output back
// in ONE D2H and synchronises.  GCL_HOST_SLOTS slots let the uploads and moment passes of the next scans overlap the evidence
// step (and the caller's own fusion work) of scan k; gcl_lidar_scan_evidence_host is submit + evidence on slot 0.
static inline size_t al(size_t x) { return (x + 255) & ~(size_t)255; }

static int grow(gcl_handle h, void** d, size_t* dcap, void** hp, size_t* hcap, size_t dbytes, size_t hbytes) {
  if (dbytes > *dcap) {
    if (*d) cudaFree(*d);
    *d = nullptr; *dcap = 0;
    GCL_CUDA(h, cudaMalloc(d, dbytes));
    *dcap = dbytes;
  }
  if (hbytes > *hcap) {
    if (*hp) cudaFreeHost(*hp);
    *hp = nullptr; *hcap = 0;
    GCL_CUDA(h, cudaMallocHost(hp, hbytes));
    *hcap = hbytes;
  }
  return GCL_OK;
}

static int ensure_streams(gcl_handle h) {
  if (h->stream) return GCL_OK;
  GCL_CUDA(h, cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  GCL_CUDA(h, cudaStreamCreateWithFlags(&h->stream_copy, cudaStreamNonBlocking));
  for (int i = 0; i < GCL_HOST_SLOTS; ++i) {
    GCL_CUDA(h, cudaStreamCreateWithFlags(&h->slot[i].stream1, cudaStreamNonBlocking));
    GCL_CUDA(h, cudaEventCreateWithFlags(&h->slot[i].ev_copy, cudaEventDisableTiming));
    GCL_CUDA(h, cudaEventCreateWithFlags(&h->slot[i].ev_stage1, cudaEventDisableTiming));
  }
  return GCL_OK;
}

// GCL_HOST_TIMING=1: per-stage wall clock of the blocking host entry (the streams are synchronised at each mark, so this mode
// serialises the stages -- for diagnosis only), averaged and printed every 64 calls
struct HostTiming {
  bool on; int n; double acc[8]; std::chrono::steady_clock::time_point t;
  HostTiming() : on(getenv("GCL_HOST_TIMING") != nullptr), n(0) { for (double& a : acc) a = 0; }
  void start() { if (on) t = std::chrono::steady_clock::now(); }
  void mark(int i, gcl_handle h, bool sync = true) {
    if (!on) return;
    if (sync) {
      cudaStreamSynchronize(h->stream_copy);
      for (int i = 0; i < GCL_HOST_SLOTS; ++i) cudaStreamSynchronize(h->slot[i].stream1);
      cudaStreamSynchronize(h->stream);
    }
    const auto now = std::chrono::steady_clock::now();
    acc[i] += std::chrono::duration<double, std::micro>(now - t).count();
    t = now;
  }
  void done() {
    if (!on || ++n < 64) return;
    fprintf(stderr, "[gcl host timing, us/call] stage+enqueue H2D %.1f | H2D done %.1f | scan_moments %.1f | reduce_finalize %.1f | chain %.1f | D2H+sync %.1f | copy-out %.1f\n",
            acc[0] / n, acc[1] / n, acc[2] / n, acc[3] / n, acc[4] / n, acc[5] / n, acc[6] / n);
    // the same figures under the reference's stage labels (pkg/backend/pipeline.py:411 point_budget_ms, :581 deskew_ms; binning and the
    // evidence operators sit under its map branch): resample, deskew, soft assign and the moment sums are ONE kernel here, so they share a label
    fprintf(stderr, "[gcl host timing, reference labels, ms/call] upload_ms %.4f | point_budget_ms+deskew_ms+bin_soft_assign_ms+moment_sums_ms (fused scan_moments) %.4f | "
                    "moment_match_finalize_ms %.4f | map_derived+matrix_fisher+planar_translation+iw+map_update_ms (chain) %.4f | readback_ms %.4f | total_ms %.4f\n",
            (acc[0] + acc[1]) / n * 1e-3, acc[2] / n * 1e-3, acc[3] / n * 1e-3, acc[4] / n * 1e-3, (acc[5] + acc[6]) / n * 1e-3,
            (acc[0] + acc[1] + acc[2] + acc[3] + acc[4] + acc[5] + acc[6]) / n * 1e-3);
    n = 0; for (double& a : acc) a = 0;
  }
};
static HostTiming g_ht;

// GCL_STAGE2_TRACE=1: host-side timeline of the single-launch evidence step, averaged over 64 calls (begin: enqueue;
// gap: between begin returning and end being called; wait: polling in end; out: copying the results to the caller) and the
// chain kernel's own duration from a pair of events
struct Stage2Trace {
  bool on; int n; double acc[5]; std::chrono::steady_clock::time_point t0, t1; cudaEvent_t e0 = nullptr, e1 = nullptr;
  Stage2Trace() : on(getenv("GCL_STAGE2_TRACE") != nullptr), n(0) { for (double& a : acc) a = 0; }
  static double us(std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
    return std::chrono::duration<double, std::micro>(b - a).count();
  }
};
static Stage2Trace g_s2;

// Is this host pointer page-locked?  The answer only selects between copying straight from the caller's array and staging
// through the pinned block -- both are correct whatever the truth is -- so the last few answers are remembered per handle
// (cudaPointerGetAttributes costs ~1 us a call; a scan stream reuses the same buffers).
static bool is_pinned(gcl_handle h, const void* p) {
  for (int i = 0; i < 8; ++i)
    if (h->pin_ptr[i] == p) return h->pin_val[i];
  cudaPointerAttributes a;
  bool v = false;
  if (cudaPointerGetAttributes(&a, p) == cudaSuccess) v = a.type == cudaMemoryTypeHost; else cudaGetLastError();
  h->pin_ptr[h->pin_next] = p; h->pin_val[h->pin_next] = v; h->pin_next = (h->pin_next + 1) & 7;
  return v;
}

// Ranges registered with gcl_host_register: the caller vouches that they stay page-locked until unregistered, so the zero-copy
// decision needs no driver query for arrays inside them.
static bool in_registered(gcl_handle h, const void* p, size_t bytes) {
  const char* c = static_cast<const char*>(p);
  std::lock_guard<std::mutex> lk(h->mu);
  for (const auto& r : h->pinned_ranges)
    if (c >= r.first && c + bytes <= r.first + r.second) return true;
  return false;
}

static bool is_pinned_now(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) == cudaSuccess) return a.type == cudaMemoryTypeHost;
  cudaGetLastError();
  return false;
}

static int submit_arrays(gcl_handle h, int slot_index, const double* points_host, const double* timestamps_host,
                         const double* weights_host, int64_t n_raw, int32_t cap, double t0, double t1, const double* xi_host,
                         int n_hyp, const double* origin_host, const double* bins_host, int n_bins, double tau, int precision,
                         double eps_psd, double eps_mass, bool latency_call) {
  if (!h) return GCL_ERR_ARG;
  GCL_REQUIRE(h, slot_index >= 0 && slot_index < GCL_HOST_SLOTS, "slot must be in [0, GCL_HOST_SLOTS)");
  GCL_REQUIRE(h, n_raw >= 0 && cap >= 1 && n_hyp >= 1 && n_hyp <= 4096, "n_raw >= 0, cap >= 1, 1 <= n_hyp <= 4096 required");
  GCL_REQUIRE(h, n_bins >= 1 && n_bins <= GCL_MAX_BINS && tau > 0.0, "1 <= n_bins <= 1024 and tau > 0 required");
  GCL_REQUIRE(h, n_raw == 0 || (points_host && timestamps_host && weights_host), "NULL scan arrays");
  GCL_REQUIRE(h, xi_host && origin_host && bins_host, "NULL pointer");
  GCL_REQUIRE(h, precision == GCL_F64 || precision == GCL_F32C, "unknown precision mode");
  cudaSetDevice(h->device);
  g_ht.start();
  int rc = ensure_streams(h);
  if (rc != GCL_OK) return rc;
  gcl_handle_s::Slot& S = h->slot[slot_index];
  if (S.pending) {  // a scan that was submitted and never collected is dropped
    cudaStreamSynchronize(S.stream1);
    S.pending = false;
  }
  const size_t B = n_bins, H = n_hyp, n = (size_t)n_raw;
  int64_t stride = (n_raw + cap - 1) / cap; if (stride < 1) stride = 1;
  const int64_t n_sel = (n_raw + stride - 1) / stride;
  const int ctas = gcl_plan_ctas_per_job(h, n_hyp, n_sel, n_bins, precision);
  const size_t rec = gcl_partial_record_doubles(n_bins), srec = gcl_stats_record_doubles(n_bins);
  // slot block (device, mirrored in pinned memory up to o_parts): [bins][jobs] | [points][timestamps][weights] | work
  // (With stride > 1 only the selected rows of points / timestamps are needed; gathering them on the host first was
  // measured: the strided CPU reads of a cold 1.2 MB scan cost 33 us against 13 us of PCIe time saved.  A caller whose
  // producer already emits compact rows can use the device-pointer API with GCL_JOB_COMPACT_ROWS.)
  // (the atlas and the jobs come first so that their device addresses do not move with the scan size)
  const size_t o_bins = 0, o_jobs = al(o_bins + B * 24), o_pts = al(o_jobs + H * sizeof(gcl_job));
  const size_t o_ts = al(o_pts + n * 24), o_w = al(o_ts + n * 8), in_end = al(o_w + n * 8);
  const size_t o_parts = in_end, o_red = al(o_parts + H * ctas * rec * 8), o_stats = al(o_red + H * rec * 8);
  const size_t d_end = al(o_stats + H * srec * 8);
  const size_t d_bytes_before = S.d_bytes;
  rc = grow(h, &S.d_buf, &S.d_bytes, &S.h_buf, &S.h_bytes, d_end, in_end);
  if (rc != GCL_OK) return rc;
  if (S.d_bytes != d_bytes_before) {   // the device block was reallocated (possibly at the same address): its atlas is gone
    S.atlas_bins = 0; S.atlas_dev = nullptr; S.atlas_copy.clear();
  }
  unsigned char* hb = static_cast<unsigned char*>(S.h_buf);
  unsigned char* db = static_cast<unsigned char*>(S.d_buf);
  // Scan arrays that already live in page-locked memory are copied straight from the caller's buffers (they must stay
  // untouched until the matching evidence call returns); pageable arrays go through the pinned block.
  const bool direct = n > 0 && is_pinned(h, points_host) && is_pinned(h, timestamps_host) && is_pinned(h, weights_host);
  // One hypothesis over page-locked arrays: no copy at all -- scan_moments reads the selected rows (and every raw weight) straight
  // out of host memory (unified addressing; each value is read exactly once, prefetched a tile ahead with cp.async).  With
  // stride 4 that is 0.54 MB over PCIe instead of 1.2 MB through the copy engine, and three cudaMemcpyAsync calls plus the
  // copy-stream event hand-off less on the host.  GCL_HOST_ZERO_COPY=0 restores the copies.
  // For the copy engine a stale "page-locked" answer from the cache is harmless; a kernel reading the array is not, so the
  // zero-copy decision asks the driver afresh (~2 us per array).  That is worth it where the latency of stage 1 counts -- the
  // blocking call: 121 -> 105 us -- and where the queries do not cost the caller's thread: submissions issued by the submit
  // thread (52 -> 50 us per scan: 0.54 MB instead of 1.2 MB competing with the evidence step's outputs for PCIe).  On the
  // caller's thread in the two-stage form (GCL_HOST_ASYNC=0) they cost 6 us per scan for nothing, so that route copies.
  static const bool zero_copy_ok = !(getenv("GCL_HOST_ZERO_COPY") && atoi(getenv("GCL_HOST_ZERO_COPY")) == 0);
  const bool zero_copy = latency_call && direct && n_hyp == 1 && zero_copy_ok && !g_ht.on &&
                         (in_registered(h, points_host, n * 24) || is_pinned_now(points_host)) &&
                         (in_registered(h, timestamps_host, n * 8) || is_pinned_now(timestamps_host)) &&
                         (in_registered(h, weights_host, n * 8) || is_pinned_now(weights_host));
  if (!direct && n > 0) {
    memcpy(hb + o_pts, points_host, n * 24);
    memcpy(hb + o_ts, timestamps_host, n * 8);
    memcpy(hb + o_w, weights_host, n * 8);
  }
  // the atlas rarely changes: it is re-uploaded only when its bytes differ from what this slot already holds on the device;
  // a single hypothesis travels as a kernel argument (by-value job), so the common call has no small H2D at all
  const bool atlas_same = S.atlas_bins == n_bins && S.atlas_copy.size() == B * 24 && memcmp(S.atlas_copy.data(), bins_host, B * 24) == 0
                          && S.atlas_dev == static_cast<void*>(db + o_bins);
  if (!atlas_same) { S.atlas_copy.assign(reinterpret_cast<const char*>(bins_host), reinterpret_cast<const char*>(bins_host) + B * 24); }
  memcpy(hb + o_bins, bins_host, B * 24);
  gcl_job* hj = reinterpret_cast<gcl_job*>(hb + o_jobs);
  for (int i = 0; i < n_hyp; ++i) {
    gcl_job& j = hj[i];
    memset(&j, 0, sizeof(j));
    j.points = db + o_pts; j.timestamps = db + o_ts; j.weights = db + o_w; j.point_lambda = nullptr;
    if (zero_copy) { j.points = points_host; j.timestamps = timestamps_host; j.weights = weights_host; }
    j.n_raw = n_raw; j.n_points_cap = cap; j.sel_begin = 0; j.sel_end = -1;
    j.scan_start_time = t0; j.scan_end_time = t1;
    for (int k = 0; k < 6; ++k) j.xi_body[k] = xi_host[i * 6 + k];
  }
  cudaStream_t sc = h->stream_copy, s1 = S.stream1;
  if (zero_copy) {
    if (!atlas_same) GCL_CUDA(h, cudaMemcpyAsync(db + o_bins, hb + o_bins, B * 24, cudaMemcpyHostToDevice, s1));
  } else if (direct) {
    GCL_CUDA(h, cudaMemcpyAsync(db + o_pts, points_host, n * 24, cudaMemcpyHostToDevice, sc));
    GCL_CUDA(h, cudaMemcpyAsync(db + o_ts, timestamps_host, n * 8, cudaMemcpyHostToDevice, sc));
    GCL_CUDA(h, cudaMemcpyAsync(db + o_w, weights_host, n * 8, cudaMemcpyHostToDevice, sc));
    if (!atlas_same || n_hyp > 1)
      GCL_CUDA(h, cudaMemcpyAsync(db + o_bins, hb + o_bins, o_pts - o_bins, cudaMemcpyHostToDevice, sc));
  } else {
    GCL_CUDA(h, cudaMemcpyAsync(db, hb, in_end, cudaMemcpyHostToDevice, sc));
  }
  S.atlas_bins = n_bins; S.atlas_dev = db + o_bins;
  if (!zero_copy) {
    GCL_CUDA(h, cudaEventRecord(S.ev_copy, sc));
    GCL_CUDA(h, cudaStreamWaitEvent(s1, S.ev_copy, 0));
  }
  g_ht.mark(0, h, false);
  g_ht.mark(1, h);
  const double* d_bins = reinterpret_cast<const double*>(db + o_bins);
  const gcl_job* d_jobs = reinterpret_cast<const gcl_job*>(db + o_jobs);
  double* d_parts = reinterpret_cast<double*>(db + o_parts);
  double* d_red = reinterpret_cast<double*>(db + o_red);
  double* d_stats = reinterpret_cast<double*>(db + o_stats);
  const gcl_job* jd = n_hyp == 1 ? nullptr : d_jobs;   // one hypothesis: the job is a kernel argument
  const gcl_job* jv = n_hyp == 1 ? hj : nullptr;
  rc = gcl_scan_moments(h, s1, jd, jv, n_hyp, ctas, GCL_LAYOUT_REF_F64, precision, d_bins, n_bins, tau, origin_host, eps_mass, d_parts);
  if (rc != GCL_OK) return rc;
  g_ht.mark(2, h);
  rc = gcl_reduce_finalize(h, s1, d_parts, ctas, (int64_t)rec, (int64_t)rec * ctas, jd, jv, n_hyp, n_bins, d_bins, tau, origin_host,
                           eps_psd, eps_mass, 1, d_red, d_stats);
  if (rc != GCL_OK) return rc;
  GCL_CUDA(h, cudaEventRecord(S.ev_stage1, s1));
  g_ht.mark(3, h);
  S.n_hyp = n_hyp; S.n_bins = n_bins; S.d_stats = d_stats; S.pending = true;
  S.eps_psd = eps_psd; S.eps_mass = eps_mass;
  return GCL_OK;
}

// Stage 1 from the MESSAGE: the raw PointCloud2 bytes (point_step bytes per point, 22 for the Kimera VLP-16 layout against 40
// bytes of parsed float64) and, optionally, the IMU buffer cross PCIe; everything between the message and the bin statistics
// runs on the device without a host round trip:
//   pass 1 (time unit + extrema) -> scan bounds (backend_node.py:1776-1786) -> [IMU preintegration over those bounds ->
//   deskew twist into the jobs] -> decode + range weights + base-frame transform into COMPACT rows (only the rows
//   PointBudgetResample keeps; every raw weight) -> scan_moments -> reduce_finalize -> D2H of the info record.
// The reference does this with a NumPy parse on the host, five H2D copies, one D2H sync for the bounds and a 512-step lax.scan.
static int submit_msg_impl(gcl_handle h, int slot_index, const uint8_t* msg_host, int64_t n_points, int32_t point_step,
                             const int32_t* off, const int32_t* dt, double header_stamp, const double* R_host,
                             const double* t_host, int32_t cap, const double* xi_host, const double* imu_stamps_host,
                             const double* imu_gyro_host, const double* imu_accel_host, int imu_samples, double imu_sigma,
                             const double* rotvec_start_host, const double* gyro_bias_host, const double* accel_bias_host,
                             const double* gravity_host, int rotation_only, int n_hyp, const double* origin_host,
                             const double* bins_host, int n_bins, double tau, int precision, double eps_psd, double eps_mass) {
  if (!h) return GCL_ERR_ARG;
  GCL_REQUIRE(h, slot_index >= 0 && slot_index < GCL_HOST_SLOTS, "slot must be in [0, GCL_HOST_SLOTS)");
  GCL_REQUIRE(h, n_points >= 0 && cap >= 1 && n_hyp >= 1 && n_hyp <= 4096, "n_points >= 0, cap >= 1, 1 <= n_hyp <= 4096 required");
  GCL_REQUIRE(h, point_step >= 1 && point_step <= 512, "1 <= point_step <= 512 required");
  GCL_REQUIRE(h, n_bins >= 1 && n_bins <= GCL_MAX_BINS && tau > 0.0, "1 <= n_bins <= 1024 and tau > 0 required");
  GCL_REQUIRE(h, off && dt && origin_host && bins_host && (n_points == 0 || msg_host), "NULL pointer");
  GCL_REQUIRE(h, (R_host == nullptr) == (t_host == nullptr), "R_base_lidar and t_base_lidar go together");
  GCL_REQUIRE(h, precision == GCL_F64 || precision == GCL_F32C, "unknown precision mode");
  const bool imu = imu_stamps_host != nullptr;
  GCL_REQUIRE(h, imu != (xi_host != nullptr), "exactly one of xi_body_host and the IMU buffer must be given");
  GCL_REQUIRE(h, !imu || (imu_gyro_host && imu_accel_host && rotvec_start_host && gyro_bias_host && accel_bias_host && gravity_host
                          && imu_samples >= 1), "incomplete IMU buffer");
  static const int width[9] = {0, 1, 1, 2, 2, 4, 4, 4, 8};
  IngestParams IP;
  memset(&IP, 0, sizeof(IP));
  for (int k = 0; k < 5; ++k) {
    const bool optional = (k == 4) && off[k] < 0;
    GCL_REQUIRE(h, optional || (dt[k] >= 1 && dt[k] <= 8), "Unsupported PointField datatype");   // backend_node.py:374
    GCL_REQUIRE(h, optional || (off[k] >= 0 && off[k] + width[dt[k]] <= point_step), "field does not fit in point_step");
    IP.off[k] = off[k]; IP.dtype[k] = optional ? 0 : dt[k];
  }
  cudaSetDevice(h->device);
  int rc = ensure_streams(h);
  if (rc != GCL_OK) return rc;
  gcl_handle_s::Slot& S = h->slot[slot_index];
  if (S.pending) { cudaStreamSynchronize(S.stream1); S.pending = false; }
  const size_t B = n_bins, H = n_hyp, n = (size_t)n_points, M = imu ? (size_t)imu_samples : 0;
  int64_t stride = (n_points + cap - 1) / cap; if (stride < 1) stride = 1;
  const int64_t n_sel = (n_points + stride - 1) / stride;
  const int ctas = gcl_plan_ctas_per_job(h, n_hyp, n_sel, n_bins, precision);
  const size_t rec = gcl_partial_record_doubles(n_bins), srec = gcl_stats_record_doubles(n_bins);
  const size_t info_doubles = GCL_MSG_INFO_HEADER + H * GCL_IMU_DOUBLES;
  const int grid_ingest = (int)((n_points + 255) / 256);
  // slot block, mirrored in pinned memory up to in_end: [bins][jobs][info][imu stamps|gyro|accel|rotvec|bg|ba][message bytes]
  // device only: [bounds parts][flag][points][timestamps][weights] | partials | reduced | stats
  const size_t o_bins = 0, o_jobs = al(o_bins + B * 24), o_info = al(o_jobs + H * sizeof(gcl_job));
  const size_t o_flag = al(o_info + info_doubles * 8);   // time-unit flag + ticket: zero in the upload, so no memset
  const size_t o_imu = al(o_flag + 16), o_msg = al(o_imu + (7 * M + 9 * H) * 8), in_end = al(o_msg + n * point_step);
  const size_t o_bparts = in_end;
  const bool packed = precision == GCL_F32C;
  const size_t o_pts = al(o_bparts + (size_t)(grid_ingest > 0 ? grid_ingest : 1) * 16), o_ts = al(o_pts + (size_t)n_sel * (packed ? 16 : 24));
  const size_t o_w = al(o_ts + (packed ? 0 : (size_t)n_sel * 8)), o_parts = al(o_w + n * (packed ? 4 : 8));
  const size_t o_red = al(o_parts + H * ctas * rec * 8), o_stats = al(o_red + H * rec * 8), d_end = al(o_stats + H * srec * 8);
  rc = grow(h, &S.d_buf, &S.d_bytes, &S.h_buf, &S.h_bytes, d_end, in_end);
  if (rc != GCL_OK) return rc;
  unsigned char* hb = static_cast<unsigned char*>(S.h_buf);
  unsigned char* db = static_cast<unsigned char*>(S.d_buf);
  S.atlas_bins = 0; S.atlas_dev = nullptr; S.atlas_copy.clear();   // this path always uploads the head of the block
  memcpy(hb + o_bins, bins_host, B * 24);
  gcl_job* hj = reinterpret_cast<gcl_job*>(hb + o_jobs);
  for (int i = 0; i < n_hyp; ++i) {
    gcl_job& j = hj[i];
    memset(&j, 0, sizeof(j));
    j.points = db + o_pts; j.timestamps = db + o_ts; j.weights = db + o_w; j.point_lambda = nullptr;
    j.n_raw = n_points; j.n_points_cap = cap; j.sel_begin = 0; j.sel_end = -1;
    j.flags = GCL_JOB_COMPACT_ROWS;
    j.scan_start_time = header_stamp; j.scan_end_time = header_stamp;   // patched on the device by scan_bounds_kernel
    for (int k = 0; k < 6; ++k) j.xi_body[k] = xi_host ? xi_host[i * 6 + k] : 0.0;
  }
  double* hinfo = reinterpret_cast<double*>(hb + o_info);
  memset(hinfo, 0, info_doubles * 8);
  memset(hb + o_flag, 0, 16);
  hinfo[GCL_MSG_INFO_N_POINTS] = (double)n_points; hinfo[GCL_MSG_INFO_STRIDE] = (double)stride; hinfo[GCL_MSG_INFO_N_SEL] = (double)n_sel;
  if (xi_host)
    for (size_t i = 0; i < H; ++i)
      for (int k = 0; k < 6; ++k) hinfo[GCL_MSG_INFO_HEADER + i * GCL_IMU_DOUBLES + GCL_IMU_XI_BODY + k] = xi_host[i * 6 + k];
  double* himu = reinterpret_cast<double*>(hb + o_imu);
  if (imu) {
    memcpy(himu, imu_stamps_host, M * 8); memcpy(himu + M, imu_gyro_host, M * 24); memcpy(himu + 4 * M, imu_accel_host, M * 24);
    memcpy(himu + 7 * M, rotvec_start_host, H * 24); memcpy(himu + 7 * M + 3 * H, gyro_bias_host, H * 24);
    memcpy(himu + 7 * M + 6 * H, accel_bias_host, H * 24);
  }
  cudaStream_t sc = h->stream_copy, s1 = S.stream1;
  const bool direct = n > 0 && is_pinned(h, msg_host);
  if (direct) {
    GCL_CUDA(h, cudaMemcpyAsync(db, hb, o_msg, cudaMemcpyHostToDevice, sc));
    GCL_CUDA(h, cudaMemcpyAsync(db + o_msg, msg_host, n * point_step, cudaMemcpyHostToDevice, sc));
  } else {
    if (n > 0) memcpy(hb + o_msg, msg_host, n * point_step);
    GCL_CUDA(h, cudaMemcpyAsync(db, hb, n > 0 ? in_end : o_msg, cudaMemcpyHostToDevice, sc));
  }
  GCL_CUDA(h, cudaEventRecord(S.ev_copy, sc));
  GCL_CUDA(h, cudaStreamWaitEvent(s1, S.ev_copy, 0));
  gcl_job* d_jobs = reinterpret_cast<gcl_job*>(db + o_jobs);
  double* d_info = reinterpret_cast<double*>(db + o_info);
  IP.data = db + o_msg; IP.n_points = n_points; IP.point_step = point_step; IP.header_stamp = header_stamp;
  IP.sentinel = 1e6; IP.sigma = 0.25; IP.min_r = 0.5; IP.max_r = 50.0; IP.weight_floor = 1e-12;   // constants.py:256-262
  IP.transform = R_host != nullptr;
  if (R_host) { for (int k = 0; k < 9; ++k) IP.R[k] = R_host[k]; for (int k = 0; k < 3; ++k) IP.t[k] = t_host[k]; }
  if (packed) { IP.packed_xyzt = reinterpret_cast<float4*>(db + o_pts); IP.packed_w = reinterpret_cast<float*>(db + o_w); }
  else { IP.points_out = reinterpret_cast<double*>(db + o_pts); IP.timestamps_out = reinterpret_cast<double*>(db + o_ts);
         IP.weights_out = reinterpret_cast<double*>(db + o_w); }
  IP.flag = reinterpret_cast<int*>(db + o_flag); IP.flag_is_zero = 1;
  IP.compact_stride = (int)stride;
  IP.bounds_parts = reinterpret_cast<double*>(db + o_bparts);
  IP.bounds_out = d_info; IP.bounds_jobs = d_jobs; IP.bounds_n_jobs = n_hyp; IP.bounds_dev = d_info;
  IP.defer_decode = 1;
  int l = 0;
  GCL_CUDA(h, launch_ingest(IP, s1, &l));        // time unit + extrema, scan bounds into the info record and the jobs
  if (imu) {
    const double* dimu = reinterpret_cast<const double*>(db + o_imu);
    ImuParams P;
    P.stamps = dimu; P.gyro = dimu + M; P.accel = dimu + 4 * M; P.weights = nullptr; P.M = imu_samples;
    P.t0 = 0; P.t1 = 0; P.sigma = imu_sigma; P.bounds_dev = d_info;
    P.rotvec_start = dimu + 7 * M; P.gyro_bias = dimu + 7 * M + 3 * H; P.accel_bias = dimu + 7 * M + 6 * H;
    P.gravity[0] = gravity_host[0]; P.gravity[1] = gravity_host[1]; P.gravity[2] = gravity_host[2];
    P.trans_scale = rotation_only ? 0.0 : 1.0;
    P.out = d_info + GCL_MSG_INFO_HEADER; P.weights_out = nullptr; P.jobs_out = d_jobs;
    GCL_CUDA(h, launch_imu_preintegrate(P, n_hyp, s1));
    l += 1;
  }
  GCL_CUDA(h, launch_ingest_decode(IP, s1, &l));
  h->launches += l;
  const double* d_bins = reinterpret_cast<const double*>(db + o_bins);
  double* d_parts = reinterpret_cast<double*>(db + o_parts);
  double* d_red = reinterpret_cast<double*>(db + o_red);
  double* d_stats = reinterpret_cast<double*>(db + o_stats);
  rc = gcl_scan_moments(h, s1, d_jobs, nullptr, n_hyp, ctas, packed ? GCL_LAYOUT_PACKED_F32 : GCL_LAYOUT_REF_F64, precision, d_bins,
                        n_bins, tau, origin_host, eps_mass, d_parts);
  if (rc != GCL_OK) return rc;
  rc = gcl_reduce_finalize(h, s1, d_parts, ctas, (int64_t)rec, (int64_t)rec * ctas, d_jobs, nullptr, n_hyp, n_bins, d_bins, tau,
                           origin_host, eps_psd, eps_mass, 1, d_red, d_stats);
  if (rc != GCL_OK) return rc;
  GCL_CUDA(h, cudaMemcpyAsync(hb + o_info, d_info, info_doubles * 8, cudaMemcpyDeviceToHost, s1));
  GCL_CUDA(h, cudaEventRecord(S.ev_stage1, s1));
  S.n_hyp = n_hyp; S.n_bins = n_bins; S.d_stats = d_stats; S.pending = true;
  S.eps_psd = eps_psd; S.eps_mass = eps_mass;
  S.info_off = o_info; S.info_doubles = info_doubles;
  return GCL_OK;
}

// ---- the submit thread ----------------------------------------------------------------------------------------------------------
static int run_task(gcl_handle h, int slot) {
  const gcl_submit_task& T = h->slot[slot].task;
  if (T.kind == 1)
    return submit_arrays(h, slot, T.points, T.timestamps, T.weights, T.n, T.cap, T.t0, T.t1, T.xi.data(), T.n_hyp, T.origin,
                         T.bins.data(), T.n_bins, T.tau, T.precision, T.eps_psd, T.eps_mass, true);   // zero-copy: the driver queries cost this thread, not the caller
  return submit_msg_impl(h, slot, T.msg, T.n, T.point_step, T.off, T.dt, T.header_stamp, T.has_T ? T.R : nullptr, T.has_T ? T.t : nullptr,
                         T.cap, T.has_xi ? T.xi.data() : nullptr, T.imu_stamps, T.imu_gyro, T.imu_accel, T.imu_samples, T.imu_sigma,
                         T.has_xi ? nullptr : T.rotvec.data(), T.has_xi ? nullptr : T.bg.data(), T.has_xi ? nullptr : T.ba.data(),
                         T.has_xi ? nullptr : T.gravity, T.rotation_only, T.n_hyp, T.origin, T.bins.data(), T.n_bins, T.tau, T.precision,
                         T.eps_psd, T.eps_mass);
}

static void submit_thread_main(gcl_handle h) {
  t_submit_thread = true;
  cudaSetDevice(h->device);
  std::unique_lock<std::mutex> lk(h->mu);
  for (;;) {
    h->cv_task.wait(lk, [&] { return h->stop || !h->queue.empty(); });
    if (h->queue.empty()) return;   // stop requested and nothing left to do
    const int slot = h->queue.front();
    h->queue.pop_front();
    lk.unlock();
    const int rc = run_task(h, slot);
    lk.lock();
    gcl_handle_s::Slot& S = h->slot[slot];
    S.task_rc = rc;
    if (rc != GCL_OK) S.task_err = h->err_submit;
    S.busy = false;
    h->cv_done.notify_all();
  }
}

static bool use_submit_thread(gcl_handle h) {
  if (h->async_mode < 0) {
    const char* e = getenv("GCL_HOST_ASYNC");
    h->async_mode = (e ? atoi(e) != 0 : true) && !g_ht.on;   // the per-stage timing mode serialises everything
  }
  return h->async_mode == 1;
}

// wait until the submit thread has finished enqueuing the task of `slot` (if any); surfaces its error
static int wait_slot(gcl_handle h, int slot) {
  gcl_handle_s::Slot& S = h->slot[slot];
  std::unique_lock<std::mutex> lk(h->mu);
  h->cv_done.wait(lk, [&] { return !S.busy; });
  const int rc = S.task_rc;
  if (rc != GCL_OK) { h->err = S.task_err; S.task_rc = GCL_OK; }
  return rc;
}

// hand the filled task of `slot` to the submit thread (started on first use)
static int post_task(gcl_handle h, int slot) {
  std::lock_guard<std::mutex> lk(h->mu);
  if (!h->worker.joinable()) h->worker = std::thread(submit_thread_main, h);
  h->slot[slot].busy = true;
  h->queue.push_back(slot);
  h->cv_task.notify_one();
  return GCL_OK;
}

int gcl_scan_submit_host(gcl_handle h, int slot_index, const double* points_host, const double* timestamps_host,
                         const double* weights_host, int64_t n_raw, int32_t cap, double t0, double t1, const double* xi_host,
                         int n_hyp, const double* origin_host, const double* bins_host, int n_bins, double tau, int precision,
                         double eps_psd, double eps_mass) {
  if (!h) return GCL_ERR_ARG;
  GCL_REQUIRE(h, slot_index >= 0 && slot_index < GCL_HOST_SLOTS, "slot must be in [0, GCL_HOST_SLOTS)");
  GCL_REQUIRE(h, n_raw >= 0 && cap >= 1 && n_hyp >= 1 && n_hyp <= 4096, "n_raw >= 0, cap >= 1, 1 <= n_hyp <= 4096 required");
  GCL_REQUIRE(h, n_bins >= 1 && n_bins <= GCL_MAX_BINS && tau > 0.0, "1 <= n_bins <= 1024 and tau > 0 required");
  GCL_REQUIRE(h, n_raw == 0 || (points_host && timestamps_host && weights_host), "NULL scan arrays");
  GCL_REQUIRE(h, xi_host && origin_host && bins_host, "NULL pointer");
  GCL_REQUIRE(h, precision == GCL_F64 || precision == GCL_F32C, "unknown precision mode");
  GCL_REQUIRE(h, h->stage2_slot != slot_index, "the evidence step of this slot is in flight: call gcl_scan_evidence_end_host first");
  int rc = wait_slot(h, slot_index);   // an earlier submission to this slot goes first (its error, if any, is reported here)
  if (rc != GCL_OK) return rc;
  if (!use_submit_thread(h))
    return submit_arrays(h, slot_index, points_host, timestamps_host, weights_host, n_raw, cap, t0, t1, xi_host, n_hyp, origin_host,
                         bins_host, n_bins, tau, precision, eps_psd, eps_mass, false);
  gcl_submit_task& T = h->slot[slot_index].task;
  T.kind = 1;
  T.points = points_host; T.timestamps = timestamps_host; T.weights = weights_host; T.n = n_raw; T.cap = cap; T.t0 = t0; T.t1 = t1;
  T.xi.assign(xi_host, xi_host + (size_t)n_hyp * 6);
  T.bins.assign(bins_host, bins_host + (size_t)n_bins * 3);
  T.n_hyp = n_hyp; T.n_bins = n_bins; T.precision = precision; T.tau = tau; T.eps_psd = eps_psd; T.eps_mass = eps_mass;
  for (int k = 0; k < 3; ++k) T.origin[k] = origin_host[k];
  return post_task(h, slot_index);
}

int gcl_scan_submit_msg_host(gcl_handle h, int slot_index, const uint8_t* msg_host, int64_t n_points, int32_t point_step,
                             const int32_t* off, const int32_t* dt, double header_stamp, const double* R_host,
                             const double* t_host, int32_t cap, const double* xi_host, const double* imu_stamps_host,
                             const double* imu_gyro_host, const double* imu_accel_host, int imu_samples, double imu_sigma,
                             const double* rotvec_start_host, const double* gyro_bias_host, const double* accel_bias_host,
                             const double* gravity_host, int rotation_only, int n_hyp, const double* origin_host,
                             const double* bins_host, int n_bins, double tau, int precision, double eps_psd, double eps_mass) {
  if (!h) return GCL_ERR_ARG;
  GCL_REQUIRE(h, slot_index >= 0 && slot_index < GCL_HOST_SLOTS, "slot must be in [0, GCL_HOST_SLOTS)");
  GCL_REQUIRE(h, n_points >= 0 && cap >= 1 && n_hyp >= 1 && n_hyp <= 4096, "n_points >= 0, cap >= 1, 1 <= n_hyp <= 4096 required");
  GCL_REQUIRE(h, point_step >= 1 && point_step <= 512, "1 <= point_step <= 512 required");
  GCL_REQUIRE(h, n_bins >= 1 && n_bins <= GCL_MAX_BINS && tau > 0.0, "1 <= n_bins <= 1024 and tau > 0 required");
  GCL_REQUIRE(h, off && dt && origin_host && bins_host && (n_points == 0 || msg_host), "NULL pointer");
  GCL_REQUIRE(h, (R_host == nullptr) == (t_host == nullptr), "R_base_lidar and t_base_lidar go together");
  GCL_REQUIRE(h, precision == GCL_F64 || precision == GCL_F32C, "unknown precision mode");
  const bool imu = imu_stamps_host != nullptr;
  GCL_REQUIRE(h, imu != (xi_host != nullptr), "exactly one of xi_body_host and the IMU buffer must be given");
  GCL_REQUIRE(h, !imu || (imu_gyro_host && imu_accel_host && rotvec_start_host && gyro_bias_host && accel_bias_host && gravity_host
                          && imu_samples >= 1), "incomplete IMU buffer");
  static const int width[9] = {0, 1, 1, 2, 2, 4, 4, 4, 8};
  for (int k = 0; k < 5; ++k) {
    const bool optional = (k == 4) && off[k] < 0;
    GCL_REQUIRE(h, optional || (dt[k] >= 1 && dt[k] <= 8), "Unsupported PointField datatype");   // backend_node.py:374
    GCL_REQUIRE(h, optional || (off[k] >= 0 && off[k] + width[dt[k]] <= point_step), "field does not fit in point_step");
  }
  GCL_REQUIRE(h, h->stage2_slot != slot_index, "the evidence step of this slot is in flight: call gcl_scan_evidence_end_host first");
  int rc = wait_slot(h, slot_index);
  if (rc != GCL_OK) return rc;
  if (!use_submit_thread(h))
    return submit_msg_impl(h, slot_index, msg_host, n_points, point_step, off, dt, header_stamp, R_host, t_host, cap, xi_host,
                           imu_stamps_host, imu_gyro_host, imu_accel_host, imu_samples, imu_sigma, rotvec_start_host, gyro_bias_host,
                           accel_bias_host, gravity_host, rotation_only, n_hyp, origin_host, bins_host, n_bins, tau, precision, eps_psd,
                           eps_mass);
  gcl_submit_task& T = h->slot[slot_index].task;
  T.kind = 2;
  T.msg = msg_host; T.n = n_points; T.point_step = point_step; T.cap = cap; T.header_stamp = header_stamp;
  for (int k = 0; k < 5; ++k) { T.off[k] = off[k]; T.dt[k] = dt[k]; }
  T.has_T = R_host != nullptr;
  if (R_host) { for (int k = 0; k < 9; ++k) T.R[k] = R_host[k]; for (int k = 0; k < 3; ++k) T.t[k] = t_host[k]; }
  T.has_xi = xi_host != nullptr;
  if (xi_host) T.xi.assign(xi_host, xi_host + (size_t)n_hyp * 6);
  T.imu_stamps = imu_stamps_host; T.imu_gyro = imu_gyro_host; T.imu_accel = imu_accel_host; T.imu_samples = imu_samples;
  T.imu_sigma = imu_sigma; T.rotation_only = rotation_only;
  if (imu) {
    T.rotvec.assign(rotvec_start_host, rotvec_start_host + (size_t)n_hyp * 3);
    T.bg.assign(gyro_bias_host, gyro_bias_host + (size_t)n_hyp * 3);
    T.ba.assign(accel_bias_host, accel_bias_host + (size_t)n_hyp * 3);
    for (int k = 0; k < 3; ++k) T.gravity[k] = gravity_host[k];
  }
  T.bins.assign(bins_host, bins_host + (size_t)n_bins * 3);
  T.n_hyp = n_hyp; T.n_bins = n_bins; T.precision = precision; T.tau = tau; T.eps_psd = eps_psd; T.eps_mass = eps_mass;
  for (int k = 0; k < 3; ++k) T.origin[k] = origin_host[k];
  return post_task(h, slot_index);
}

// the info record of the last gcl_scan_submit_msg_host on `slot`: waits for stage 1 of that slot only
int gcl_scan_msg_info_host(gcl_handle h, int slot_index, double* info_host, int64_t n_doubles) {
  if (!h) return GCL_ERR_ARG;
  GCL_REQUIRE(h, slot_index >= 0 && slot_index < GCL_HOST_SLOTS, "slot must be in [0, GCL_HOST_SLOTS)");
  gcl_handle_s::Slot& S = h->slot[slot_index];
  { const int rcw = wait_slot(h, slot_index); if (rcw != GCL_OK) return rcw; }
  GCL_REQUIRE(h, S.info_doubles > 0 && S.h_buf, "no message was submitted to this slot");
  GCL_REQUIRE(h, info_host && n_doubles >= (int64_t)S.info_doubles, "info buffer too small (GCL_MSG_INFO_HEADER + n_hyp * GCL_IMU_DOUBLES doubles)");
  cudaSetDevice(h->device);
  GCL_CUDA(h, cudaEventSynchronize(S.ev_stage1));
  memcpy(info_host, static_cast<unsigned char*>(S.h_buf) + S.info_off, S.info_doubles * 8);
  return GCL_OK;
}

// stage 2, first half: everything is enqueued (H2D of pose / map / IW, chain step, the one D2H); no waiting
int gcl_scan_evidence_begin_host(gcl_handle h, int slot_index, const double* pose_host, const double* mS_dir, const double* mS_sc,
                                 const double* mN_dir, const double* mN_pos, const double* msum_p, const double* msum_ppT,
                                 const double* pose_update_host, double forgetting, const double* iw_state_host,
                                 int use_sigma_meas) {
  if (!h) return GCL_ERR_ARG;
  GCL_REQUIRE(h, slot_index >= 0 && slot_index < GCL_HOST_SLOTS, "slot must be in [0, GCL_HOST_SLOTS)");
  gcl_handle_s::Slot& S = h->slot[slot_index];
  { const int rcw = wait_slot(h, slot_index); if (rcw != GCL_OK) return rcw; }   // the submit thread is done with this slot
  GCL_REQUIRE(h, S.pending, "no scan was submitted to this slot");
  GCL_REQUIRE(h, h->stage2_slot < 0, "an evidence step is already in flight: call gcl_scan_evidence_end_host first");
  GCL_REQUIRE(h, pose_host, "NULL pointer");
  GCL_REQUIRE(h, mS_dir && mS_sc && mN_dir && mN_pos && msum_p && msum_ppT, "NULL map statistics");
  GCL_REQUIRE(h, !pose_update_host || (forgetting > 0.0 && forgetting <= 1.0), "forgetting factor must be in (0, 1]");
  GCL_REQUIRE(h, !use_sigma_meas || iw_state_host, "use_sigma_meas needs iw_state_host");
  cudaSetDevice(h->device);
  if (g_s2.on) g_s2.t0 = std::chrono::steady_clock::now();
  const size_t B = S.n_bins, H = S.n_hyp;
  const size_t srec = gcl_stats_record_doubles(S.n_bins);
  // stage-2 block, mirrored in pinned memory: [pose][pose_update][map][iw] | [stats][evidence] | map-derived scratch
  // ONE H2D carries pose .. iw, ONE D2H carries map .. evidence.
  const size_t o_pose = 0, o_pupd = al(o_pose + H * 48), o_map = al(o_pupd + 48);  // S_dir 3B, S_sc 9B, N_dir B, N_pos B, sum_p 3B, sum_ppT 9B
  const size_t o_iw = al(o_map + B * 26 * 8), o_stats = al(o_iw + 80);
  const size_t o_evid = al(o_stats + H * srec * 8), io_end = al(o_evid + H * GCL_EVID_DOUBLES * 8);
  const size_t d_end = al(io_end + B * 16 * 8);
  const size_t d_bytes_before = h->d_bytes;
  int rc = grow(h, &h->d_buf, &h->d_bytes, &h->h_buf, &h->h_bytes, d_end, io_end);
  if (rc != GCL_OK) return rc;
  if (h->d_bytes != d_bytes_before) h->map_shadow_bins = 0;   // a new device block (even at the old address) holds nothing yet
  if (!h->done_flag) {
    GCL_CUDA(h, cudaMallocHost(reinterpret_cast<void**>(&h->done_flag), 64));
    *h->done_flag = 0;
  }
  unsigned char* hb = static_cast<unsigned char*>(h->h_buf);
  unsigned char* db = static_cast<unsigned char*>(h->d_buf);
  // the blocking call runs stage 2 on the slot's own stage-1 stream: plain stream order instead of a cross-stream event wait
  cudaStream_t st = h->stage2_on_slot_stream ? S.stream1 : h->stream;
  h->stage2_stream = st;
  // small blocks (one or a few hypotheses): ONE launch, the chain kernel stages its own inputs and outputs across PCIe and the
  // host polls a ticket; large blocks (many hypotheses) go through the copy engine
  static const bool single_launch_ok = !(getenv("GCL_HOST_SINGLE_LAUNCH") && atoi(getenv("GCL_HOST_SINGLE_LAUNCH")) == 0);
  const bool single = single_launch_ok && !g_ht.on && (io_end - o_map) <= 96 * 1024;
  double* hm = reinterpret_cast<double*>(hb + o_map);
  memcpy(hm, mS_dir, B * 24); memcpy(hm + 3 * B, mS_sc, B * 72); memcpy(hm + 12 * B, mN_dir, B * 8);
  memcpy(hm + 13 * B, mN_pos, B * 8); memcpy(hm + 14 * B, msum_p, B * 24); memcpy(hm + 17 * B, msum_ppT, B * 72);
  memcpy(hb + o_pose, pose_host, H * 48);
  if (pose_update_host) memcpy(hb + o_pupd, pose_update_host, 48);
  if (iw_state_host) memcpy(hb + o_iw, iw_state_host, 80);
  if (!single) GCL_CUDA(h, cudaMemcpyAsync(db + o_pose, hb + o_pose, (o_iw + 80) - o_pose, cudaMemcpyHostToDevice, st));
  if (st != S.stream1) GCL_CUDA(h, cudaStreamWaitEvent(st, S.ev_stage1, 0));
  double* dm = reinterpret_cast<double*>(db + o_map);
  double* d_stats = reinterpret_cast<double*>(db + o_stats);
  double* d_evid = reinterpret_cast<double*>(db + o_evid);
  double* d_mder = reinterpret_cast<double*>(db + io_end);
  double* d_iw = iw_state_host ? reinterpret_cast<double*>(db + o_iw) : nullptr;
  // the chain step rewrites Sigma_p in the stats records: work on the copy that travels back
  if (single) {
    ChainParams P;
    P.stats = d_stats; P.stats_stride = (int64_t)srec;
    P.map_S_dir = dm; P.map_S_dir_scatter = dm + 3 * B; P.map_N_dir = dm + 12 * B; P.map_N_pos = dm + 13 * B; P.map_sum_p = dm + 14 * B;
    P.map_sum_ppT = dm + 17 * B; P.workspace = d_mder; P.pose_pred = reinterpret_cast<const double*>(db + o_pose);
    P.pose_update = pose_update_host ? reinterpret_cast<const double*>(db + o_pupd) : nullptr;
    P.iw_state = d_iw; P.evidence = d_evid; P.n_scans = 1; P.n_hyp = S.n_hyp; P.n_bins = S.n_bins;
    P.use_sigma_meas = use_sigma_meas; P.forgetting = forgetting; P.iw_rho = 0.99 /* GC_IW_RHO_MEAS_LIDAR, constants.py:281 */;
    P.iw_nu_weak_add = 0.5 /* GC_IW_NU_WEAK_ADD, :164 */; P.iw_nu_max = 1000.0; P.eps_psd = S.eps_psd; P.eps_mass = S.eps_mass;
    P.chain_smem = 0;
    // What has to cross PCIe for the kernel to start?  The poses travel as kernel arguments (few hypotheses).  The map and the
    // IW state usually are what the previous step left behind -- the caller hands back the arrays it was given -- and the device
    // still holds them: they are compared with the shadow of that output (10 KB, < 1 us) and fetched only if they differ.  A
    // kernel-side PCIe read costs ~2 us alone, but queues behind the 1.2 MB upload of the next scan when that is in flight.
    const bool inline_pose = H * 6 + 6 <= 30;
    if (inline_pose) {
      memcpy(P.pose_inline, pose_host, H * 48);
      if (pose_update_host) memcpy(P.pose_inline + H * 6, pose_update_host, 48);
      P.pose_inline_n = (int)(H * 6 + 6); P.pose_inline_update = pose_update_host != nullptr;
    }
    const size_t map_bytes = B * 26 * 8;
    const bool map_same = h->map_shadow_bins == (int)B && h->map_shadow_off == o_map && h->map_shadow.size() == map_bytes + 80 &&
                          memcmp(h->map_shadow.data(), hb + o_map, map_bytes) == 0 &&
                          (!iw_state_host || memcmp(h->map_shadow.data() + map_bytes, hb + o_iw, 80) == 0);
    const size_t in_lo = inline_pose ? o_map : o_pose, in_hi = map_same ? o_map : o_iw + 80;
    P.stage_in_src = reinterpret_cast<const double*>(hb + in_lo); P.stage_in_dst = reinterpret_cast<double*>(db + in_lo);
    P.stage_in_doubles = in_hi > in_lo ? (int)((in_hi - in_lo) / 8) : 0;
    h->map_shadow_bins = 0;   // until this step's output is in (gcl_scan_evidence_end_host)
    P.stats_src = S.d_stats; P.stats_doubles = (int)(H * srec);
    P.stage_out_src = reinterpret_cast<const double*>(db + o_map); P.stage_out_dst = reinterpret_cast<double*>(hb + o_map);
    P.stage_out_doubles = (int)((io_end - o_map) / 8);
    P.out_off_iw = (int)((o_iw - o_map) / 8); P.out_off_stats = (int)((o_stats - o_map) / 8); P.out_off_evid = (int)((o_evid - o_map) / 8);
    P.done_flag = h->done_flag; P.done_value = ++h->done_seq;
    int l = 0;
    if (g_s2.on) {
      if (!g_s2.e0) { cudaEventCreate(&g_s2.e0); cudaEventCreate(&g_s2.e1); }
      cudaEventRecord(g_s2.e0, st);
    }
    GCL_CUDA(h, launch_chain(P, st, &l));
    if (g_s2.on) { cudaEventRecord(g_s2.e1, st); g_s2.t1 = std::chrono::steady_clock::now(); g_s2.acc[0] += Stage2Trace::us(g_s2.t0, g_s2.t1); }
    h->launches += l;
    h->stage2_slot = slot_index;
    h->stage2_has_update = pose_update_host != nullptr;
    h->stage2_has_iw = iw_state_host != nullptr;
    h->stage2_single = true;
    return GCL_OK;
  }
  h->stage2_single = false;
  h->map_shadow_bins = 0;
  GCL_CUDA(h, cudaMemcpyAsync(d_stats, S.d_stats, H * srec * 8, cudaMemcpyDeviceToDevice, st));
  rc = gcl_scan_chain(h, st, d_stats, (int64_t)srec, 1, S.n_hyp, S.n_bins, dm, dm + 3 * B, dm + 12 * B, dm + 13 * B, dm + 14 * B,
                      dm + 17 * B, d_mder, reinterpret_cast<const double*>(db + o_pose),
                      pose_update_host ? reinterpret_cast<const double*>(db + o_pupd) : nullptr, forgetting, d_iw, use_sigma_meas,
                      0.99 /* GC_IW_RHO_MEAS_LIDAR, constants.py:281 */, 0.5 /* GC_IW_NU_WEAK_ADD, :164 */, 1000.0, S.eps_psd,
                      S.eps_mass, d_evid);
  if (rc != GCL_OK) return rc;
  g_ht.mark(4, h);
  // one D2H for every output: [map][iw][stats][evidence] (the map / IW blocks ride along even when they were not updated)
  GCL_CUDA(h, cudaMemcpyAsync(hb + o_map, db + o_map, io_end - o_map, cudaMemcpyDeviceToHost, st));
  h->stage2_slot = slot_index;
  h->stage2_has_update = pose_update_host != nullptr;
  h->stage2_has_iw = iw_state_host != nullptr;
  return GCL_OK;
}

// stage 2, second half: wait for the D2H and hand the results out
int gcl_scan_evidence_end_host(gcl_handle h, double* mS_dir, double* mS_sc, double* mN_dir, double* mN_pos, double* msum_p,
                               double* msum_ppT, double* iw_state_host, double* stats_host, double* evidence_host) {
  if (!h) return GCL_ERR_ARG;
  GCL_REQUIRE(h, h->stage2_slot >= 0, "no evidence step is in flight");
  GCL_REQUIRE(h, stats_host && evidence_host, "NULL pointer");
  gcl_handle_s::Slot& S = h->slot[h->stage2_slot];
  const bool pose_update_host = h->stage2_has_update;
  GCL_REQUIRE(h, !pose_update_host || (mS_dir && mS_sc && mN_dir && mN_pos && msum_p && msum_ppT), "NULL map statistics");
  GCL_REQUIRE(h, !h->stage2_has_iw || iw_state_host, "iw_state_host is NULL but the step carried an IW state");
  cudaSetDevice(h->device);
  const size_t B = S.n_bins, H = S.n_hyp;
  const size_t srec = gcl_stats_record_doubles(S.n_bins);
  const size_t o_pose = 0, o_pupd = al(o_pose + H * 48), o_map = al(o_pupd + 48);
  const size_t o_iw = al(o_map + B * 26 * 8), o_stats = al(o_iw + 80);
  const size_t o_evid = al(o_stats + H * srec * 8);
  unsigned char* hb = static_cast<unsigned char*>(h->h_buf);
  cudaStream_t st = h->stage2_stream ? h->stage2_stream : h->stream;
  std::chrono::steady_clock::time_point tr_a;
  if (g_s2.on && h->stage2_single) { tr_a = std::chrono::steady_clock::now(); g_s2.acc[1] += Stage2Trace::us(g_s2.t1, tr_a); }
  if (h->stage2_single) {
    // poll the ticket the chain kernel writes after its outputs (a stream synchronisation costs ~8 us of wake-up latency);
    // fall back to the stream if it does not arrive, which also surfaces a kernel fault
    const unsigned long long want = h->done_seq;
    const auto t_start = std::chrono::steady_clock::now();
    unsigned spins = 0;
    while (__atomic_load_n(h->done_flag, __ATOMIC_ACQUIRE) != want) {
      if ((++spins & 0x3ff) == 0 && std::chrono::steady_clock::now() - t_start > std::chrono::seconds(2)) {
        GCL_CUDA(h, cudaStreamSynchronize(st));
        break;
      }
#if defined(__x86_64__)
      __builtin_ia32_pause();
#endif
    }
  } else {
    GCL_CUDA(h, cudaStreamSynchronize(st));
  }
  std::chrono::steady_clock::time_point tr_b;
  if (g_s2.on && h->stage2_single) { tr_b = std::chrono::steady_clock::now(); g_s2.acc[2] += Stage2Trace::us(tr_a, tr_b); }
  if (h->stage2_single) {   // the page-locked block now mirrors the device's map and IW state
    h->map_shadow.assign(reinterpret_cast<const char*>(hb + o_map), B * 26 * 8);
    h->map_shadow.append(reinterpret_cast<const char*>(hb + o_iw), 80);
    h->map_shadow_bins = (int)B; h->map_shadow_off = o_map;
  }
  S.pending = false;
  h->stage2_slot = -1;
  g_ht.mark(5, h, false);
  memcpy(stats_host, hb + o_stats, H * srec * 8);
  memcpy(evidence_host, hb + o_evid, H * GCL_EVID_DOUBLES * 8);
  if (pose_update_host) {
    const double* om = reinterpret_cast<const double*>(hb + o_map);
    memcpy(mS_dir, om, B * 24); memcpy(mS_sc, om + 3 * B, B * 72); memcpy(mN_dir, om + 12 * B, B * 8);
    memcpy(mN_pos, om + 13 * B, B * 8); memcpy(msum_p, om + 14 * B, B * 24); memcpy(msum_ppT, om + 17 * B, B * 72);
  }
  if (h->stage2_has_iw) memcpy(iw_state_host, hb + o_iw, 80);
  g_ht.mark(6, h, false);
  g_ht.done();
  if (g_s2.on && h->stage2_single) {
    g_s2.acc[3] += Stage2Trace::us(tr_b, std::chrono::steady_clock::now());
    float ms = 0.f;
    cudaEventSynchronize(g_s2.e1);
    cudaEventElapsedTime(&ms, g_s2.e0, g_s2.e1);
    g_s2.acc[4] += 1e3 * ms;
    if (++g_s2.n == 64) {
      fprintf(stderr, "[gcl stage 2, us/call] begin %.1f | gap %.1f | wait %.1f | out %.1f | chain kernel (events) %.1f\n", g_s2.acc[0] / 64,
              g_s2.acc[1] / 64, g_s2.acc[2] / 64, g_s2.acc[3] / 64, g_s2.acc[4] / 64);
      g_s2.n = 0; for (double& a : g_s2.acc) a = 0;
    }
  }
  return GCL_OK;
}

int gcl_host_register(gcl_handle h, const void* ptr, size_t bytes) {
  if (!h) return GCL_ERR_ARG;
  GCL_REQUIRE(h, ptr && bytes > 0, "NULL pointer or empty range");
  cudaSetDevice(h->device);
  GCL_REQUIRE(h, is_pinned_now(ptr) && is_pinned_now(static_cast<const char*>(ptr) + bytes - 1),
              "the range is not page-locked host memory (cudaHostAlloc / cudaHostRegister / torch pin_memory)");
  std::lock_guard<std::mutex> lk(h->mu);
  for (auto& r : h->pinned_ranges)
    if (r.first == static_cast<const char*>(ptr)) { r.second = bytes; return GCL_OK; }
  h->pinned_ranges.emplace_back(static_cast<const char*>(ptr), bytes);
  return GCL_OK;
}

int gcl_host_unregister(gcl_handle h, const void* ptr) {
  if (!h) return GCL_ERR_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  for (size_t i = 0; i < h->pinned_ranges.size(); ++i)
    if (h->pinned_ranges[i].first == static_cast<const char*>(ptr)) {
      h->pinned_ranges.erase(h->pinned_ranges.begin() + i);
      return GCL_OK;
    }
  h->err = "gcl_host_unregister: range not registered";
  return GCL_ERR_ARG;
}

int gcl_scan_evidence_host(gcl_handle h, int slot_index, const double* pose_host, double* mS_dir, double* mS_sc, double* mN_dir,
                           double* mN_pos, double* msum_p, double* msum_ppT, const double* pose_update_host, double forgetting,
                           double* iw_state_host, int use_sigma_meas, double* stats_host, double* evidence_host) {
  if (!h) return GCL_ERR_ARG;
  GCL_REQUIRE(h, stats_host && evidence_host, "NULL pointer");
  int rc = gcl_scan_evidence_begin_host(h, slot_index, pose_host, mS_dir, mS_sc, mN_dir, mN_pos, msum_p, msum_ppT, pose_update_host,
                                        forgetting, iw_state_host, use_sigma_meas);
  if (rc != GCL_OK) return rc;
  return gcl_scan_evidence_end_host(h, mS_dir, mS_sc, mN_dir, mN_pos, msum_p, msum_ppT, iw_state_host, stats_host, evidence_host);
}

int gcl_lidar_scan_evidence_host(gcl_handle h, const double* points_host, const double* timestamps_host,
                                 const double* weights_host, int64_t n_raw, int32_t cap, double t0, double t1,
                                 const double* xi_host, const double* pose_host, int n_hyp, const double* origin_host,
                                 const double* bins_host, int n_bins, double tau, int precision, double* mS_dir, double* mS_sc,
                                 double* mN_dir, double* mN_pos, double* msum_p, double* msum_ppT, const double* pose_update_host,
                                 double forgetting, double* iw_state_host, int use_sigma_meas, double eps_psd, double eps_mass,
                                 double* stats_host, double* evidence_host) {
  if (!h) return GCL_ERR_ARG;
  GCL_REQUIRE(h, pose_host && stats_host && evidence_host, "NULL pointer");
  GCL_REQUIRE(h, mS_dir && mS_sc && mN_dir && mN_pos && msum_p && msum_ppT, "NULL map statistics");
  GCL_REQUIRE(h, !pose_update_host || (forgetting > 0.0 && forgetting <= 1.0), "forgetting factor must be in (0, 1]");
  GCL_REQUIRE(h, !use_sigma_meas || iw_state_host, "use_sigma_meas needs iw_state_host");
  int rc = wait_slot(h, 0);
  if (rc != GCL_OK) return rc;
  rc = submit_arrays(h, 0, points_host, timestamps_host, weights_host, n_raw, cap, t0, t1, xi_host, n_hyp, origin_host,
                     bins_host, n_bins, tau, precision, eps_psd, eps_mass, true);   // blocking call: no hop to the submit thread
  if (rc != GCL_OK) return rc;
  h->stage2_on_slot_stream = true;
  rc = gcl_scan_evidence_host(h, 0, pose_host, mS_dir, mS_sc, mN_dir, mN_pos, msum_p, msum_ppT, pose_update_host, forgetting,
                              iw_state_host, use_sigma_meas, stats_host, evidence_host);
  h->stage2_on_slot_stream = false;
  return rc;
}

}  // extern "C"
