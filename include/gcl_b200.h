/*
 * gcl_b200.h -- C ABI of the B200-native GC-SLAM v2 LiDAR evidence hot path.
 *
 * The reference (whabacivch/FL-SLAM) has no FFI for this path: its operators are Python functions over JAX
 * arrays (docs/OPERATOR_CONTRACTS.md:3).  Each entry point below therefore cites the reference *function*
 * it replaces (file:line; pkg/ = fl_ws/src/fl_slam_poc/fl_slam_poc/).  INTEGRATION.md shows the ctypes stub a
 * maintainer of the reference would add to call them.
 *
 * Conventions
 *   - every function returns 0 on success, a negative gcl_status otherwise; gcl_last_error(h) has the text.
 *     No CPU fallback exists anywhere behind this ABI (docs/GC_SLAM.md:128-132 "no GPU-else-CPU").
 *   - pointers are DEVICE pointers unless the parameter name ends in _host.
 *   - `stream` is a cudaStream_t passed as void*; all device-pointer entry points are asynchronous on it and
 *     perform no host synchronisation.  *_host entry points synchronise before returning.
 *   - all floating point results are IEEE binary64 unless stated; integer/byte outputs are bit-exact.
 *   - a handle is not thread-safe (the reference calls this path from one LiDAR worker thread,
 *     pkg/backend/backend_node.py:1367); use one handle per (device, thread).
 */
#ifndef GCL_B200_H_
#define GCL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GCL_ABI_VERSION 1
#define GCL_MAX_BINS 1024
#define GCL_NM 19            /* moments per bin: 1 + d(3) + ddT(6) + p(3) + ppT(6) */
#define GCL_NCERT 8          /* per-partial certificate scalars (see GCL_CS_*) */
#define GCL_STATS_SCALARS 24 /* per-job scalar block written by gcl_reduce_finalize (see GCL_SS_*) */
#define GCL_EVID_DOUBLES 640 /* per-job evidence record written by gcl_pose_evidence (see GCL_EV_*) */

typedef struct gcl_handle_s* gcl_handle;

enum gcl_status {
  GCL_OK = 0,
  GCL_ERR_ARG = -1,      /* shape / contract violation (reference: ValueError, binning.py:253-264) */
  GCL_ERR_CUDA = -2,     /* CUDA runtime error */
  GCL_ERR_NO_DEVICE = -3,
  GCL_ERR_UNSUPPORTED = -4
};

enum gcl_precision {
  GCL_F64 = 0,  /* all arithmetic binary64: parity mode (moments <= 1e-9 rel of the reference) */
  GCL_F32C = 1  /* binary32 per-point math + packed-fp32 tile accumulation carried in binary64 (<= 1e-5 rel) */
};

enum gcl_layout {
  GCL_LAYOUT_REF_F64 = 0, /* the reference's arrays: points (n,3) f64 row-major, timestamps (n,) f64, weights (n,) f64 */
  GCL_LAYOUT_PACKED_F32 = 1 /* float4 {x,y,z,t - scan_start} per point + float weight: 20 B/point */
};

/* One unit of work for the fused kernel: one scan under one hypothesis (twist).  Lives in device memory when a
 * batch is launched (scan stream, hypothesis batch), or is passed by value for a single scan. */
typedef struct gcl_job {
  const void* points;      /* layout-dependent, raw (pre-resample) */
  const void* timestamps;  /* f64 (n,) for REF_F64; unused for PACKED_F32 */
  const void* weights;     /* f64 (n,) or f32 (n,) */
  const double* point_lambda; /* optional (cap,) per-selected-row multiplier (binning.py:157), or NULL */
  int64_t n_raw;           /* raw points in this scan */
  int32_t n_points_cap;    /* PointBudget cap; stride = max(1, ceil(n_raw / cap)) (point_budget.py:160) */
  int32_t sel_begin;       /* this rank's slice of the selected-row index space [sel_begin, sel_end) ... */
  int32_t sel_end;         /* ... (0, n_sel) on one GPU; a contiguous sub-range when points are sharded */
  int32_t flags;           /* GCL_JOB_COMPACT_ROWS: points/timestamps hold only the SELECTED rows (row k = raw row k*stride),
                              already gathered by the caller; weights stay raw (all n_raw, needed for the mass sum) */
  double scan_start_time;
  double scan_end_time;
  double xi_body[6];       /* constant twist over the sweep, [rho, phi] (deskew_constant_twist.py:31-69) */
} gcl_job;

#define GCL_JOB_COMPACT_ROWS 1
#define GCL_HOST_SLOTS 4       /* scans that may be in flight through gcl_scan_submit_host */

/* ---- lifecycle ------------------------------------------------------------------------------------------ */
int gcl_abi_version(void);
int gcl_create(int device, gcl_handle* out);
void gcl_destroy(gcl_handle h);
const char* gcl_last_error(gcl_handle h); /* h may be NULL: error of the last failed gcl_create */
int gcl_device_sm_count(gcl_handle h);
/* number of kernels this handle has launched so far (bench.py reports it as gpu_launches) */
int64_t gcl_launch_count(gcl_handle h);

/* ---- fused hot path ------------------------------------------------------------------------------------- */
/* Layout of one partial record: GCL_NM * n_bins moments (moment-major: [k][bin]) then GCL_NCERT scalars. */
size_t gcl_partial_record_doubles(int n_bins);
/* CTAs per job the launcher will use for (n_jobs, largest job) -- size the partials buffer with it. */
int gcl_plan_ctas_per_job(gcl_handle h, int n_jobs, int64_t max_sel_rows, int n_bins, int precision);

/* Replaces, in one pass over the points and without materialising (N,B) or (N,3,3):
 *   point_budget_resample gather   pkg/backend/operators/point_budget.py:50-109
 *   deskew_constant_twist          pkg/backend/operators/deskew_constant_twist.py:31-69 (+ imu_preintegration.py:19-43)
 *   ray directions                 pkg/backend/pipeline.py:591-593
 *   bin_soft_assign                archive/legacy_operators/binning.py:56-76
 *   scan_bin_moment_match (sums)   archive/legacy_operators/binning.py:157-173
 * Writes partials[job][cta] records (fixed order => deterministic).  Exactly one of jobs_dev / job_by_value_host
 * is used: jobs_dev != NULL launches n_jobs jobs read from device memory; otherwise the single job is passed by
 * value from host memory. */
int gcl_scan_moments(gcl_handle h, void* stream, const gcl_job* jobs_dev, const gcl_job* job_by_value_host,
                     int n_jobs, int ctas_per_job, int layout, int precision, const double* bin_dirs /* (B,3) */,
                     int n_bins, double tau, const double* lidar_origin_host /* (3,) */, double eps_mass,
                     double* partials);

enum { /* indices into the GCL_NCERT tail of a partial record */
  GCL_CS_SUM_W_RAW = 0,   /* sum of all raw weights in this CTA's raw range           (point_budget.py:69) */
  GCL_CS_SUM_W_SEL = 1,   /* sum of selected raw weights                               (point_budget.py:79) */
  GCL_CS_SUM_W_SEL_SQ = 2,
  GCL_CS_SUM_W_DESKEW = 3,/* sum of w_sel * w_time                                     (deskew_constant_twist.py:102) */
  GCL_CS_SUM_ENTROPY = 4, /* sum of per-point assignment entropies                     (binning.py:70-74) */
  GCL_CS_MAX_RESP = 5,    /* max responsibility (reduced with max, not +)              (binning.py:69) */
  GCL_CS_ROWS = 6,        /* selected rows processed */
  GCL_CS_SPARE = 7
};

/* Fixed-order sum of `n_parts` partial records per job, then (if finalize != 0) the non-linear per-bin part:
 *   mass rescale                   point_budget.py:80-85
 *   inv mass, centroid, scatter    binning.py:175-182, pkg/common/primitives.py:195-212
 *   PSD projection per bin         pkg/common/primitives.py:80-123
 *   kappa                          pkg/backend/operators/kappa.py:130-169
 *   certificate scalars            point_budget.py:95-96,:202-213; deskew_constant_twist.py:102-109;
 *                                  binning.py:70-75,:118-125 (incl. the padded rows); binning.py:201-204
 * With finalize == 0 it only writes the reduced record to reduced[job] (the buffer that is exchanged between
 * GPUs when points are sharded); a second call with parts = gathered records and finalize = 1 completes it.
 * stats layout per job (doubles): N[B] s_dir[3B] S_dir_scatter[9B] p_bar[3B] Sigma_p[9B] kappa[B] sum_p[3B]
 * sum_ppT[9B] then GCL_STATS_SCALARS scalars. */
size_t gcl_stats_record_doubles(int n_bins);
int gcl_reduce_finalize(gcl_handle h, void* stream, const double* parts, int n_parts, int64_t part_stride_doubles,
                        int64_t job_stride_doubles, const gcl_job* jobs_dev, const gcl_job* job_by_value_host,
                        int n_jobs, int n_bins, const double* bin_dirs, double tau,
                        const double* lidar_origin_host, double eps_psd, double eps_mass, int finalize,
                        double* reduced /* [n_jobs][record] or NULL */, double* stats /* [n_jobs][stats record] */);

enum { /* indices into the scalar tail of a stats record */
  GCL_SS_TOTAL_MASS_IN = 0, GCL_SS_TOTAL_MASS_SEL = 1, GCL_SS_MASS_SCALE = 2, GCL_SS_PB_ESS = 3,
  GCL_SS_N_SELECTED = 4, GCL_SS_STRIDE = 5, GCL_SS_DESKEW_RETAINED = 6, GCL_SS_AVG_ENTROPY = 7,
  GCL_SS_MAX_RESP = 8, GCL_SS_MM_ESS = 9, GCL_SS_MM_SUPPORT_FRAC = 10, GCL_SS_MM_PSD_DELTA = 11,
  GCL_SS_MM_MAX_EPS_RATIO = 12, GCL_SS_ENTROPY_VALID_ROWS = 13, GCL_SS_MAX_RESP_VALID_ROWS = 14,
  GCL_SS_SUM_W_OUT = 15
};

/* Map-side derived statistics.  Replaces compute_map_derived_stats, archive/bin_atlas.py:166-224.
 * out layout: mu_dir[3B] kappa[B] centroid[3B] Sigma_c[9B]. */
int gcl_map_derived_stats(gcl_handle h, void* stream, const double* S_dir, const double* N_dir, const double* N_pos,
                          const double* sum_p, const double* sum_ppT, int n_bins, double eps_mass, double eps_psd,
                          double* out);

/* Matrix Fisher rotation + planar translation evidence + 22-D embedding, batched over hypotheses (one warp each).
 * Replaces matrix_fisher_rotation_evidence / planar_translation_evidence / build_combined_lidar_evidence_22d,
 * archive/legacy_operators/matrix_fisher_evidence.py:155-394, :413-671, :729-756.
 * scan stats: one stats record per job (stride stats_stride_doubles; 0 = all hypotheses share record 0).
 * pose_pred: (n_jobs,6) [t, rotvec] = belief_pred.mean_world_pose() (pkg/common/belief.py:408-434).
 * If R_hat_override != NULL (n_jobs,9) it is used as R_hat for the translation step instead of R_mf.
 * evidence: n_jobs records of GCL_EVID_DOUBLES doubles (see GCL_EV_*). */
int gcl_pose_evidence(gcl_handle h, void* stream, const double* stats, int64_t stats_stride_doubles,
                      const double* map_S_dir, const double* map_S_dir_scatter, const double* map_N_dir,
                      const double* map_N_pos, const double* map_derived, const double* pose_pred,
                      const double* R_hat_override, int n_jobs, int n_bins, double eps_psd, double eps_mass,
                      int do_rotation, int do_translation, double* evidence);

enum { /* offsets (doubles) into an evidence record */
  GCL_EV_R_MF = 0, GCL_EV_L_ROT = 9, GCL_EV_H_ROT = 18, GCL_EV_DELTA_ROT = 21, GCL_EV_SVD_S = 24,
  GCL_EV_ROT_CERT = 27,      /* [eig_min, eig_max, cond, near_null, nll_per_ess, directional_score, psd_delta,
                                 mass_eps_ratio, rot_nll, N_eff] */
  GCL_EV_SCAN_METRICS = 37,  /* [l1,l2,l3, linearity, planarity, sphericity, anisotropy, effective_rank] */
  GCL_EV_MAP_METRICS = 45,
  GCL_EV_SCAN_EIGVECS = 53, GCL_EV_MAP_EIGVECS = 62,
  GCL_EV_T_WLS = 71, GCL_EV_L_TRANS = 74, GCL_EV_H_TRANS = 83, GCL_EV_DELTA_TRANS = 86,
  GCL_EV_TRANS_CERT = 89,    /* [eig_min, eig_max, cond, near_null, nll_per_ess, psd_delta, mass_eps_ratio,
                                 trans_nll, N_eff, xy_info, z_info, z_precision_scale] */
  GCL_EV_H22 = 101, GCL_EV_L22 = 123 /* 22 + 484 doubles; ends at 607 <= GCL_EVID_DOUBLES */
};

/* ---- reference operators one by one (device pointers; used by the drop-in Python operators) --------------- */
/* point_budget_resample, pkg/backend/operators/point_budget.py:117-221.  Outputs are cap-sized and zero padded.
 * scalars_out (8 doubles): [n_input, n_selected, stride, total_mass_in, total_mass_selected, mass_scale, ess, 0]. */
int gcl_point_budget_resample(gcl_handle h, void* stream, const double* points, const double* timestamps,
                              const double* weights, const uint8_t* ring, const uint8_t* tag, int64_t n_input,
                              int32_t n_points_cap, double eps_mass, double* points_out, double* timestamps_out,
                              double* weights_out, uint8_t* ring_out, uint8_t* tag_out, double* scalars_out,
                              double* workspace /* >= 4096 doubles */);

/* deskew_constant_twist, pkg/backend/operators/deskew_constant_twist.py:72-117.
 * scalars_out (4 doubles): [sum_w_out, sum_w_in, retained_fraction, 0]. */
int gcl_deskew_constant_twist(gcl_handle h, void* stream, const double* points, const double* timestamps,
                              const double* weights, int64_t n, double scan_start_time, double scan_end_time,
                              const double* xi_body_host, double eps_mass, double* points_out, double* weights_out,
                              double* scalars_out, double* workspace /* >= 4096 doubles */);

/* ray directions, pkg/backend/pipeline.py:591-593. */
int gcl_ray_directions(gcl_handle h, void* stream, const double* points, int64_t n, const double* origin_host,
                       double eps_mass, double* dirs_out);

/* bin_soft_assign, archive/legacy_operators/binning.py:79-131.  responsibilities (N,B) row-major.
 * scalars_out (4 doubles): [avg_entropy, max_resp, exp(avg_entropy), 0]. */
int gcl_bin_soft_assign(gcl_handle h, void* stream, const double* point_dirs, int64_t n, const double* bin_dirs,
                        int n_bins, double tau, double eps_mass, double* responsibilities, double* scalars_out,
                        double* workspace /* >= 4096 doubles */);

/* scan_bin_moment_match with materialised responsibilities and optional per-point covariances,
 * archive/legacy_operators/binning.py:212-324.  stats: one stats record (gcl_stats_record_doubles). */
size_t gcl_moment_match_workspace_doubles(gcl_handle h, int64_t n, int n_bins);
int gcl_scan_bin_moment_match(gcl_handle h, void* stream, const double* points, const double* point_covariances,
                              const double* weights, const double* responsibilities, const double* point_lambda,
                              const double* origin_host, int64_t n, int n_bins, double eps_psd, double eps_mass,
                              double* stats, double* workspace);

/* kappa_from_resultant_batch, pkg/backend/operators/kappa.py:130-169. */
int gcl_kappa_from_resultant(gcl_handle h, void* stream, const double* R_bar, int64_t n, double eps_r, int d,
                             double r0, double tau, double* kappa_out);

/* create_fibonacci_atlas, archive/bin_atlas.py:40-75. */
int gcl_fibonacci_atlas(gcl_handle h, void* stream, int n_bins, double eps_mass, double* dirs_out);

/* Map maintenance ("next" row): world-frame increments of one scan's bin sums under pose (R, t) with t_z forced
 * to 0, forgetting and additive update in one kernel.  archive/bin_atlas.py:137-163, :232-257; the pushforward rule
 * itself is not in the reference tree -- declared in DESIGN.md.  Map arrays are updated in place. */
int gcl_map_update(gcl_handle h, void* stream, const double* stats /* one stats record */, const double* R_world,
                   const double* t_world, int n_bins, double forgetting, double* S_dir, double* S_dir_scatter,
                   double* N_dir, double* N_pos, double* sum_p, double* sum_ppT);

/* The sequential part of a scan stream in ONE launch: a single CTA walks n_scans scans in order and carries the map bin
 * statistics and the LiDAR measurement-noise inverse-Wishart block on the device.  Per scan: [bin covariances +=
 * Sigma_meas] -> map-derived statistics -> pose evidence (n_hyp hypotheses) -> IW update -> map update.
 * Replaces the per-scan host loop pkg/backend/backend_node.py:2036-2119 for this path; the IW rules are
 * pkg/backend/operators/measurement_noise_iw_jax.py:37-127 (mode, residual sufficient statistics, forgetful update
 * with the softplus nu projection), GC_IW_RHO_MEAS_LIDAR / GC_IW_NU_WEAK_ADD from pkg/common/constants.py:281,:164.
 * stats: n_scans*n_hyp records from gcl_reduce_finalize (scan-major); Sigma_p is rewritten in place when
 * use_sigma_meas != 0.  pose_update: (n_scans,6) pose applied by the map update (NULL = map is read-only).
 * iw_state: 10 doubles [nu, Psi (3x3)] updated in place (NULL = no measurement-noise adaptation).
 * workspace: gcl_scan_chain_workspace_doubles(n_scans, n_hyp, n_bins) doubles of scratch.
 * A single scan runs as one launch; a stream (n_scans >= 4) runs staged: map recurrence, map-derived statistics and
 * rotation evidence in parallel over scans, and only Sigma_meas -> t_wls -> residuals -> IW update sequentially. */
size_t gcl_scan_chain_workspace_doubles(int n_scans, int n_hyp, int n_bins);
int gcl_scan_chain(gcl_handle h, void* stream, double* stats, int64_t stats_stride_doubles, int n_scans, int n_hyp,
                   int n_bins, double* map_S_dir, double* map_S_dir_scatter, double* map_N_dir, double* map_N_pos,
                   double* map_sum_p, double* map_sum_ppT, double* workspace, const double* pose_pred,
                   const double* pose_update, double forgetting, double* iw_state, int use_sigma_meas, double iw_rho,
                   double iw_nu_weak_add, double iw_nu_max, double eps_psd, double eps_mass, double* evidence);

/* gcl_scan_chain in two halves, for pipelining consecutive batches of one stream on two CUDA streams: phase 1 launches what
 * depends only on the MAP left by the previous batch (world-frame increments, map recurrence, map-derived statistics, rotation
 * evidence), phase 2 what depends on its IW state (the sequential measurement-noise core, translation finish).  Phase 1 of batch
 * k+1 may run while phase 2 of batch k is still walking its scans (separate stats / workspace / evidence buffers per batch);
 * phase 0 = both, i.e. gcl_scan_chain.  A chain shorter than 4 scans runs whole in phase 2. */
int gcl_scan_chain_phase(gcl_handle h, void* stream, int phase, double* stats, int64_t stats_stride_doubles, int n_scans,
                         int n_hyp, int n_bins, double* map_S_dir, double* map_S_dir_scatter, double* map_N_dir,
                         double* map_N_pos, double* map_sum_p, double* map_sum_ppT, double* workspace,
                         const double* pose_pred, const double* pose_update, double forgetting, double* iw_state,
                         int use_sigma_meas, double iw_rho, double iw_nu_weak_add, double iw_nu_max, double eps_psd,
                         double eps_mass, double* evidence);

/* ---- multi-GPU: the one exchange step of the point-sharded path, over NVLink peer memory ---------------- */
/* The reference runs on one GPU in one process and has no collective; every raw bin sum of
 * scan_bin_moment_match is linear in the points (archive/legacy_operators/binning.py:166-173), so one scan's points can be
 * sharded over the GPUs of a node (one process per GPU) with ONE exchange of the per-bin record per job.
 *   gcl_exchange_create   allocates this rank's exchange block (two parities x world slots of slot_bytes, flag words) and
 *                         returns its CUDA IPC handle (GCL_EXCHANGE_HANDLE_BYTES bytes) for the caller to all-gather by any
 *                         means (torch.distributed in the Python layer);
 *   gcl_exchange_connect  maps the peers' blocks (all_handles: world x GCL_EXCHANGE_HANDLE_BYTES, rank-major);
 *   gcl_sharded_reduce_finalize
 *                         = gcl_reduce_finalize(finalize 0) + all-gather + gcl_reduce_finalize(finalize 1) in ONE launch: each
 *                         CTA sums this rank's partial records in a fixed order and stores the result straight into every
 *                         peer's block over NVLink; the last CTA publishes the step number in the peers' flag words, waits for
 *                         theirs, and runs the per-bin finalisation over the gathered records in rank order (bit-identical
 *                         statistics on every rank).  Collective: every rank must call it, in the same order, once per step.
 *                         A peer that never arrives trips a time limit (GCL_EXCHANGE_TIMEOUT_S, default 2 s): the kernel
 *                         returns without finalising and gcl_exchange_error_flag() becomes non-zero.
 *   gcl_exchange_all_gather
 *                         the same publish / wait protocol for a small device buffer (hypothesis-sharded evidence records):
 *                         src (bytes) lands in slot [rank] of every rank; dst (world x bytes) may be NULL, in which case the
 *                         data are read in place through gcl_exchange_gathered(x, src_rank). */
#define GCL_MAX_RANKS 8
#define GCL_EXCHANGE_HANDLE_BYTES 64
#define GCL_EXCHANGE_LANES 2
typedef struct gcl_exchange_s* gcl_exchange;
int gcl_exchange_create(gcl_handle h, int world, int rank, size_t slot_bytes, gcl_exchange* out,
                        unsigned char* ipc_handle_out /* GCL_EXCHANGE_HANDLE_BYTES */);
int gcl_exchange_connect(gcl_handle h, gcl_exchange x, const unsigned char* all_handles);
int gcl_exchange_disconnect(gcl_handle h, gcl_exchange x); /* unmap the peers' blocks; barrier; then gcl_exchange_destroy */
void gcl_exchange_destroy(gcl_handle h, gcl_exchange x);
const char* gcl_exchange_last_error(gcl_exchange x);
int gcl_exchange_error_flag(gcl_exchange x); /* synchronous read of the device-side error word: 0 = no time-out so far */
/* Cap the grid of gcl_sharded_reduce_finalize (0 = one CTA per 128 record elements).  A caller that overlaps the exchange of scan k
 * with the moment pass of scan k + 1 leaves a few SMs free for it (gcl_scan_moments with ctas_per_job < SM count) and caps the
 * exchange kernel's grid to that number, so that its CTAs do not queue behind the moment pass. */
int gcl_exchange_set_max_ctas(gcl_exchange x, int max_ctas);
/* A block holds GCL_EXCHANGE_LANES independent copies of the protocol state (slots, flag words, ticket, step counter).  Calls run on
 * the lane selected here (default 0).  Calls on ONE lane must be stream-ordered among themselves; calls on different lanes may be in
 * flight at the same time on different streams (the tails of two consecutive scans: gcl_scan_pipe).  Every rank must use the same
 * lane for the same step. */
int gcl_exchange_set_lane(gcl_exchange x, int lane);
int gcl_sharded_reduce_finalize(gcl_handle h, gcl_exchange x, void* stream, const double* parts, int n_parts,
                                int64_t part_stride_doubles, int64_t job_stride_doubles, const gcl_job* jobs_dev,
                                const gcl_job* job_by_value_host, int n_jobs, int n_bins, const double* bin_dirs, double tau,
                                const double* lidar_origin_host, double eps_psd, double eps_mass,
                                double* reduced /* [n_jobs][record]: the combined record */, double* stats);
int gcl_exchange_all_gather(gcl_handle h, gcl_exchange x, void* stream, const void* src, size_t bytes, void* dst);
const void* gcl_exchange_gathered(gcl_exchange x, int src_rank);

/* ---- consecutive scans (or hypothesis batches) in flight -------------------------------------------------------------------- */
/* The reference processes hypotheses and scans one after another on one stream (pkg/backend/backend_node.py:2036-2119).  Per
 * scan this library's path is three launches: gcl_scan_moments (O(N), the whole GPU), the exchange + finalisation, the evidence
 * chain (both O(B), a few CTAs, latency-bound).  A gcl_scan_pipe keeps the O(B) tail of scan k beside the moment pass of scan k + 1
 * (and, with n_tails = 2, beside the tail of scan k + 1 as well: scan k runs on exchange lane k mod 2) on streams of its own, so the
 * caller makes ONE call per scan.  The caller owns the buffers: n_sets sets of
 *   partials  [n_jobs][ctas_per_job][gcl_partial_record_doubles]      reduced [n_jobs][gcl_partial_record_doubles]
 *   stats     [n_jobs][gcl_stats_record_doubles]                      workspace gcl_scan_chain_workspace_doubles(1, n_jobs, n_bins)
 *   evidence  [n_jobs][GCL_EVID_DOUBLES]                              gathered [world][n_jobs][GCL_EVID_DOUBLES] (mode GATHER, world > 1)
 * and names the set a submission writes (`set`, any order; a set is rewritten only after the tail that read it has finished).
 * Results of a submission are valid on the caller's stream after gcl_scan_pipe_join; everything a submission reads (jobs, poses,
 * map, scan arrays) must stay valid until n_sets further submissions or a join + synchronisation.
 *   GCL_PIPE_SUM_RECORDS      points sharded over the ranks: tail = gcl_sharded_reduce_finalize -> chain step (needs the exchange;
 *                             world 1 runs the same kernel without peers)
 *   GCL_PIPE_GATHER_EVIDENCE  hypotheses sharded over the ranks: tail = gcl_reduce_finalize -> chain step -> gcl_exchange_all_gather
 *                             of the evidence records (no exchange needed at world 1)
 * Collective like the exchange: every rank submits the same sequence.  The map is read-only here (no pose update, no IW state). */
#define GCL_PIPE_MAX_SETS 8
enum { GCL_PIPE_SUM_RECORDS = 0, GCL_PIPE_GATHER_EVIDENCE = 1 };
typedef struct gcl_scan_pipe_s* gcl_scan_pipe;
int gcl_scan_pipe_create(gcl_handle h, gcl_exchange x, int world, int mode, int n_sets /* n_tails < n_sets <= GCL_PIPE_MAX_SETS */,
                         int n_tails /* 1 .. GCL_EXCHANGE_LANES */, gcl_scan_pipe* out);
void gcl_scan_pipe_destroy(gcl_scan_pipe p);
const char* gcl_scan_pipe_last_error(gcl_scan_pipe p);
int gcl_scan_pipe_submit(gcl_scan_pipe p, void* caller_stream, int wait_for_caller_stream, int set, const gcl_job* jobs_dev,
                         const gcl_job* job_by_value_host, int n_jobs, int ctas_per_job, int layout, int precision,
                         const double* bin_dirs, int n_bins, double tau, const double* lidar_origin_host, double eps_psd,
                         double eps_mass, const double* map_S_dir, const double* map_S_dir_scatter, const double* map_N_dir,
                         const double* map_N_pos, const double* map_sum_p, const double* map_sum_ppT, const double* pose_pred,
                         double* partials, double* reduced, double* stats, double* workspace, double* evidence, double* gathered);
int gcl_scan_pipe_join(gcl_scan_pipe p, void* caller_stream);
/* Scans that arrive in HOST memory (the reference hands every scan over as NumPy arrays: pkg/backend/backend_node.py:1761-1790, and
 * reads the results back on the host: :2106-2119).  gcl_scan_pipe_wait_set orders `stream` after the last submission that used `set`
 * (its moment pass and its tail have finished): the caller enqueues the H2D copies of the next scan into the staging buffers it ties
 * to that set on `stream` and then submits with caller_stream = `stream`, wait_for_caller_stream = 1 -- the upload of scan k + 1 runs
 * under the kernels of scan k.  gcl_scan_pipe_set_readback names page-locked host buffers ([n_jobs][gcl_stats_record_doubles],
 * [n_jobs][GCL_EVID_DOUBLES]; either may be NULL) that every later submission on `set` fills at the end of its tail; they are valid
 * after gcl_scan_pipe_join + a synchronisation of the caller's stream, until the set is submitted again. */
int gcl_scan_pipe_wait_set(gcl_scan_pipe p, int set, void* stream);
int gcl_scan_pipe_set_readback(gcl_scan_pipe p, int set, double* stats_host, double* evidence_host);

/* Convert the reference arrays to the packed 20 B/point layout (x,y,z,t-scan_start as float4; w as float). */
int gcl_pack_points_f32(gcl_handle h, void* stream, const double* points, const double* timestamps,
                        const double* weights, int64_t n, double scan_start_time, void* packed_xyzt /* float4[n] */,
                        float* packed_w);

/* ---- PointCloud2 ingest ("next" row) ------------------------------------------------------- */
/* parse_pointcloud2_vlp16 (pkg/backend/backend_node.py:377-468; datatype codes :351-374) + the base-frame transform
 * (:1675-1681) on the raw message bytes in DEVICE memory: n_points records of point_step bytes.
 *   field_offsets / field_datatypes  5 entries each: x, y, z, ring, per-point time ("t" else "time"; offset -1 if absent).
 *                                    Datatypes are sensor_msgs/PointField codes 1..8.
 *   R_base_lidar_host, t_base_lidar_host   (9,), (3,) or NULL/NULL for no transform (points stay in the LiDAR frame)
 *   outputs (any may be NULL): points (n,3) f64, timestamps (n,) f64, weights (n,) f64, ring (n,) u8, tag (n,) u8; and / or
 *   the packed 20 B/point layout (float4 {x,y,z,t - packed_t0} + float w) that the fp32c kernel reads.
 *   workspace: at least 1 double of device scratch.
 * Constants of the reference: sentinel 1e6 (GC_NONFINITE_SENTINEL), range window 0.25 / 0.5 / 50 m, weight floor 1e-12. */
int gcl_pointcloud2_ingest(gcl_handle h, void* stream, const uint8_t* data, int64_t n_points, int32_t point_step,
                           const int32_t* field_offsets_host, const int32_t* field_datatypes_host,
                           double header_stamp_sec, const double* R_base_lidar_host, const double* t_base_lidar_host,
                           double* points_out, double* timestamps_out, double* weights_out, uint8_t* ring_out,
                           uint8_t* tag_out, void* packed_xyzt, float* packed_w, double packed_t0, double* workspace);

/* ---- IMU preintegration -> deskew twist ("next" row) ---------------------------------------- */
/* smooth_window_weights + preintegrate_imu_relative_pose_jax + se3_log as called at pkg/backend/pipeline.py:437-483
 * (pkg/backend/operators/imu_preintegration.py:19-43, :46-147; pkg/common/geometry/se3_jax.py:179-256), batched over
 * n_hyp hypotheses (start orientation + biases), one warp each; the sequential lax.scan becomes a prefix product of
 * rotations and prefix sums.  weights == NULL: the window weights of (scan_start_time, scan_end_time, sigma) are used.
 * out: n_hyp records of GCL_IMU_DOUBLES doubles (offsets GCL_IMU_*).  jobs_out (optional): xi_body is also written into
 * n_hyp gcl_job records, so that the twist never leaves the device. */
#define GCL_IMU_DOUBLES 40
enum {
  GCL_IMU_DELTA_POSE = 0, GCL_IMU_DELTA_R = 6, GCL_IMU_DELTA_P = 15, GCL_IMU_DELTA_V = 18, GCL_IMU_ESS = 21,
  GCL_IMU_A_BODY_MEAN = 22, GCL_IMU_A_WORLD_NOG_MEAN = 25, GCL_IMU_A_WORLD_MEAN = 28, GCL_IMU_DT_SUM = 31,
  GCL_IMU_XI_BODY = 32
};
int gcl_imu_preintegrate(gcl_handle h, void* stream, const double* imu_stamps, const double* imu_gyro,
                         const double* imu_accel, const double* weights, int n_samples, double scan_start_time,
                         double scan_end_time, double sigma, const double* rotvec_start_WB, const double* gyro_bias,
                         const double* accel_bias, const double* gravity_W_host, int deskew_rotation_only, int n_hyp,
                         double* out, double* weights_out, gcl_job* jobs_out);

/* ---- reference-facing call with HOST buffers ------------------------------------------------------------- */
/* The whole chain for one scan and n_hyp hypotheses: H2D of the scan, fused moments, finalize, the chain step (map
 * derived statistics, pose evidence, optional IW + map update), D2H of the stats and evidence records; synchronises.
 * This is the entry a maintainer binds in place of the operator sequence at pkg/backend/pipeline.py:400-593 +
 * README.md:110-115 (and, with pose_update_host, the per-scan map / noise bookkeeping of
 * pkg/backend/backend_node.py:2106-2119).  map_*_host are the six MapBinStats arrays (archive/bin_atlas.py:83-101);
 * xi_body_host and pose_pred_host are (n_hyp,6).  Scan arrays that live in page-locked memory are copied directly.
 *   pose_update_host  (6,) or NULL: if given, the map arrays are updated IN PLACE (forgetting, then this scan's bin sums
 *                     pushed to the world frame) and written back to the caller's host arrays.
 *   iw_state_host     (10,) [nu, Psi(3x3)] or NULL: LiDAR measurement-noise IW block, updated in place.
 *   use_sigma_meas    add the IW mode to every bin covariance before the translation evidence. */
int gcl_lidar_scan_evidence_host(gcl_handle h, const double* points_host, const double* timestamps_host,
                                 const double* weights_host, int64_t n_raw, int32_t n_points_cap,
                                 double scan_start_time, double scan_end_time, const double* xi_body_host,
                                 const double* pose_pred_host, int n_hyp, const double* lidar_origin_host,
                                 const double* bin_dirs_host, int n_bins, double tau, int precision,
                                 double* map_S_dir_host, double* map_S_dir_scatter_host, double* map_N_dir_host,
                                 double* map_N_pos_host, double* map_sum_p_host, double* map_sum_ppT_host,
                                 const double* pose_update_host, double forgetting, double* iw_state_host,
                                 int use_sigma_meas, double eps_psd, double eps_mass,
                                 double* stats_host /* n_hyp stats records */,
                                 double* evidence_host /* n_hyp evidence records */);

/* The same call in two stages, for callers that want the upload and the moment pass of scan k+1 to overlap the evidence
 * step of scan k (and their own fusion work between the two).  Stage 1 needs only the scan and the IMU twist
 * (pkg/backend/pipeline.py:400-593 up to the moment match: nothing there depends on the pose prediction or the map);
 * stage 2 needs the pose prediction and the map (archive/legacy_operators/matrix_fisher_evidence.py:264-275, :502-516).
 *   gcl_scan_submit_host    asynchronous: H2D on a copy stream, scan_moments + reduce_finalize on a second stream; returns
 *                           at once.  Page-locked input arrays are read in place and must stay untouched until the matching
 *                           gcl_scan_evidence_host returns.  slot is in [0, GCL_HOST_SLOTS).
 *   gcl_scan_evidence_host  waits on the GPU for stage 1 of `slot`, runs the chain step, ONE D2H of map / IW / stats /
 *                           evidence, synchronises.  Arguments as in gcl_lidar_scan_evidence_host.
 * The CUDA calls of a submission (gcl_scan_submit_host, gcl_scan_submit_msg_host) are issued from a submit thread owned by the
 * handle: arguments are checked and the small arrays copied before the call returns, errors of the thread surface at the next
 * call on the same slot, and the caller's thread spends ~1 us in the call.  The scan arrays / message bytes / IMU samples are
 * read later: keep them valid and unchanged until the matching evidence call has returned.  GCL_HOST_ASYNC=0 issues everything
 * from the caller's thread instead. */
int gcl_scan_submit_host(gcl_handle h, int slot, const double* points_host, const double* timestamps_host,
                         const double* weights_host, int64_t n_raw, int32_t n_points_cap, double scan_start_time,
                         double scan_end_time, const double* xi_body_host, int n_hyp, const double* lidar_origin_host,
                         const double* bin_dirs_host, int n_bins, double tau, int precision, double eps_psd,
                         double eps_mass);
/* Stage 1 straight from the MESSAGE (the callers either side of the path, SURVEY.md 8f-3 / 8f-4 joined to stage 1): replaces
 * parse_pointcloud2_vlp16 (pkg/backend/backend_node.py:377-468), the base-frame transform (:1675-1681), the scan time bounds
 * (:1776-1786: |max - min| < 1e-9 ? [min, stamp] : [min(stamp, min), max(stamp, max)]) and, when the IMU buffer is given,
 * smooth_window_weights + preintegrate_imu_relative_pose_jax + se3_log over those bounds (pkg/backend/pipeline.py:437-483),
 * followed by the same moment pass as gcl_scan_submit_host.  The raw message bytes (point_step per point) are what crosses
 * PCIe; nothing returns to the host between the message and the bin statistics.
 *   field_offsets / field_datatypes   as in gcl_pointcloud2_ingest (x, y, z, ring, per-point time or -1)
 *   xi_body_host (n_hyp,6)            the deskew twist, OR (exactly one of the two)
 *   imu_*                             the IMU buffer (n_imu samples; stamps (M,), gyro (M,3), accel (M,3)), window width
 *                                     imu_sigma, and per hypothesis the start orientation and biases (n_hyp,3) each
 * The evidence step is the usual gcl_scan_evidence_begin_host / _end_host on the same slot.  gcl_scan_msg_info_host returns
 * GCL_MSG_INFO_HEADER doubles [scan start, scan end, min t, max t, time-unit flag (1 = nanoseconds), n_points, stride,
 * n_selected] followed by n_hyp IMU records of GCL_IMU_DOUBLES doubles (only XI_BODY is set when the twist was given). */
#define GCL_MSG_INFO_HEADER 8
enum {
  GCL_MSG_INFO_T0 = 0, GCL_MSG_INFO_T1 = 1, GCL_MSG_INFO_TMIN = 2, GCL_MSG_INFO_TMAX = 3, GCL_MSG_INFO_NS = 4,
  GCL_MSG_INFO_N_POINTS = 5, GCL_MSG_INFO_STRIDE = 6, GCL_MSG_INFO_N_SEL = 7
};
int gcl_scan_submit_msg_host(gcl_handle h, int slot, const uint8_t* msg_data_host, int64_t n_points, int32_t point_step,
                             const int32_t* field_offsets_host, const int32_t* field_datatypes_host, double header_stamp_sec,
                             const double* R_base_lidar_host, const double* t_base_lidar_host, int32_t n_points_cap,
                             const double* xi_body_host, const double* imu_stamps_host, const double* imu_gyro_host,
                             const double* imu_accel_host, int n_imu, double imu_sigma, const double* rotvec_start_WB_host,
                             const double* gyro_bias_host, const double* accel_bias_host, const double* gravity_W_host,
                             int deskew_rotation_only, int n_hyp, const double* lidar_origin_host, const double* bin_dirs_host,
                             int n_bins, double tau, int precision, double eps_psd, double eps_mass);
int gcl_scan_msg_info_host(gcl_handle h, int slot, double* info_host, int64_t n_doubles);

/* Page-locked scan buffers that are used scan after scan (a ring of receive buffers) can be registered once: the caller vouches
 * that [ptr, ptr + bytes) stays page-locked until gcl_host_unregister, and the entry points that read scan arrays in place
 * (zero-copy) skip their per-call driver query (~2 us per array) for arrays inside a registered range.  Registration itself
 * checks the range with the driver and fails for pageable memory. */
int gcl_host_register(gcl_handle h, const void* ptr, size_t bytes);
int gcl_host_unregister(gcl_handle h, const void* ptr);

/* gcl_scan_evidence_host in two halves, so that the caller can issue the next scan's gcl_scan_submit_host while the GPU runs
 * this scan's chain step: _begin enqueues everything (H2D of pose / map / IW, chain step, the one D2H) and returns; _end
 * waits for the D2H and hands out the results (map / IW arrays are written only if the step carried an update / an IW state).
 * One evidence step may be in flight at a time. */
int gcl_scan_evidence_begin_host(gcl_handle h, int slot, const double* pose_pred_host, const double* map_S_dir_host,
                                 const double* map_S_dir_scatter_host, const double* map_N_dir_host,
                                 const double* map_N_pos_host, const double* map_sum_p_host,
                                 const double* map_sum_ppT_host, const double* pose_update_host, double forgetting,
                                 const double* iw_state_host, int use_sigma_meas);
int gcl_scan_evidence_end_host(gcl_handle h, double* map_S_dir_host, double* map_S_dir_scatter_host, double* map_N_dir_host,
                               double* map_N_pos_host, double* map_sum_p_host, double* map_sum_ppT_host,
                               double* iw_state_host, double* stats_host, double* evidence_host);
int gcl_scan_evidence_host(gcl_handle h, int slot, const double* pose_pred_host, double* map_S_dir_host,
                           double* map_S_dir_scatter_host, double* map_N_dir_host, double* map_N_pos_host,
                           double* map_sum_p_host, double* map_sum_ppT_host, const double* pose_update_host,
                           double forgetting, double* iw_state_host, int use_sigma_meas, double* stats_host,
                           double* evidence_host);

#ifdef __cplusplus
}
#endif
#endif /* GCL_B200_H_ */
