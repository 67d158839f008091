// gcl_internal.cuh -- declarations shared by the translation units of libgcl_b200.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/gcl_b200.h"
#include "gcl_math.cuh"

#ifndef GCL_NT_F32
#define GCL_NT_F32 256  // threads per CTA of the fp32 scan_moments kernel
#endif
#ifndef GCL_GMAX_F32
#define GCL_GMAX_F32 4   // bins per thread (register tile) of the fp32 scan_moments kernel
#endif
#ifndef GCL_USE_FFMA2
#define GCL_USE_FFMA2 1  // packed fp32 FMA (fma.rn.f32x2, sm_100+); 0 = two scalar FFMAs (for A/B timing)
#endif

// One stage-1 submission, as handed to the submit thread: small arrays by value, the big ones (scan arrays, message bytes, IMU
// samples) by pointer -- the caller keeps those valid and unchanged until the matching evidence call has returned.
struct gcl_submit_task {
  int kind = 0;                           // 1 = parsed arrays (gcl_scan_submit_host), 2 = message (gcl_scan_submit_msg_host)
  const double *points = nullptr, *timestamps = nullptr, *weights = nullptr;
  int64_t n = 0; int32_t cap = 0; double t0 = 0, t1 = 0;
  const uint8_t* msg = nullptr; int32_t point_step = 0; int32_t off[5] = {0, 0, 0, 0, 0}, dt[5] = {0, 0, 0, 0, 0};
  double header_stamp = 0; bool has_T = false; double R[9] = {0}, t[3] = {0};
  bool has_xi = false; const double *imu_stamps = nullptr, *imu_gyro = nullptr, *imu_accel = nullptr;
  int imu_samples = 0; double imu_sigma = 0, gravity[3] = {0, 0, 0}; int rotation_only = 0;
  std::vector<double> xi, rotvec, bg, ba, bins;
  int n_hyp = 0, n_bins = 0, precision = 0; double origin[3] = {0, 0, 0}, tau = 0, eps_psd = 0, eps_mass = 0;
};

struct gcl_handle_s {
  int device = 0;
  int sm_count = 148;
  std::atomic<int64_t> launches{0};
  // submit thread: stage 1 of the two-stage host entry points is enqueued from a second host thread.  The ten-odd CUDA calls of a
  // submission cost 25-45 us of host time; the caller's thread meanwhile runs the evidence step of the previous scan, which
  // since it became one launch plus a polled ticket needs almost no driver time of its own.
  std::thread worker;
  std::mutex mu;
  std::condition_variable cv_task, cv_done;
  std::deque<int> queue;                 // slots with a task waiting
  bool stop = false;
  int async_mode = -1;                   // -1 = not decided yet (GCL_HOST_ASYNC, default 1)
  std::string err;                       // error text of calls made on the caller's thread
  std::string err_submit;                // ... and of the submit thread (handed over through Slot::task_err)
  // resources of the *_host entry points (allocated on first use, grown on demand)
  unsigned long long done_seq = 0;     // completion ticket of the single-launch evidence step (0 = the copy-engine path is in flight)
  std::string map_shadow;              // map (26 B doubles) + IW state (10) exactly as the last single-launch step left them on the DEVICE
  int map_shadow_bins = 0;             // 0 = the device copy is not known to match anything
  size_t map_shadow_off = 0;           // where in the stage-2 block that map lives
  unsigned long long* done_flag = nullptr;   // page-locked word the chain kernel writes the ticket to
  cudaStream_t stream = nullptr;       // stage 2: chain step + D2H
  cudaStream_t stream_copy = nullptr;  // H2D of the scans
  void* d_buf = nullptr; size_t d_bytes = 0;   // stage-2 block
  void* h_buf = nullptr; size_t h_bytes = 0;
  struct Slot {                         // one scan in flight
    void* d_buf = nullptr; size_t d_bytes = 0;
    void* h_buf = nullptr; size_t h_bytes = 0;
    cudaEvent_t ev_copy = nullptr, ev_stage1 = nullptr;
    cudaStream_t stream1 = nullptr;     // stage 1 of this slot (its own stream: the stage-1 chains of scans in flight overlap)
    double* d_stats = nullptr;
    int n_hyp = 0, n_bins = 0;
    double eps_psd = 0, eps_mass = 0;
    bool pending = false;
    std::string atlas_copy;             // bytes of the atlas this slot holds on the device
    int atlas_bins = 0; void* atlas_dev = nullptr;
    size_t info_off = 0, info_doubles = 0;   // info record of the last gcl_scan_submit_msg_host (pinned block)
    gcl_submit_task task;                    // the submission in flight on the submit thread (busy = not yet enqueued)
    bool busy = false; int task_rc = 0; std::string task_err;
  } slot[GCL_HOST_SLOTS];
  const void* pin_ptr[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  bool pin_val[8] = {false, false, false, false, false, false, false, false};
  int pin_next = 0;
  int stage2_slot = -1;                 // slot whose evidence step is enqueued and not yet collected
  bool stage2_has_update = false, stage2_has_iw = false, stage2_single = false, stage2_on_slot_stream = false;
  cudaStream_t stage2_stream = nullptr;   // the stream the evidence step in flight was enqueued on
  std::vector<std::pair<const char*, size_t>> pinned_ranges;   // gcl_host_register: ranges the caller vouches for as page-locked
};

namespace gcl {

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) applies PER DEVICE, and one process may hold handles on several devices
// (and enqueue from the caller's and the submit thread at once): one bit per device ordinal, set atomically after the opt-in.
template <typename K>
inline cudaError_t optin_dynamic_smem(K kern, int bytes, std::atomic<uint64_t>& done_mask) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const uint64_t bit = 1ull << (dev & 63);
  if (done_mask.load(std::memory_order_acquire) & bit) return cudaSuccess;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return e;
  done_mask.fetch_or(bit, std::memory_order_release);
  return cudaSuccess;
}

struct ScanConst {
  double origin[3];
  double eps_mass;
  float xi32[6];  // single-job fp32 copy of the twist (jobs array: converted per CTA instead)
};

struct ScanParams {
  const gcl_job* jobs;  // device array or nullptr
  gcl_job job;          // used when jobs == nullptr
  ScanConst c;
  const double* bin_dirs;  // (B,3) device
  const double* resp;      // (N,B) device, FROM_RESP variant only
  const double* point_cov; // (N,3,3) device, FROM_RESP variant only (may be nullptr => zeros)
  double* partials;
  int64_t rec_doubles;
  double tau;
  int n_bins;
  int lanes;        // L
  int n_chunks;     // bin chunks of G*L
  int tile_points;  // fused kernel: upper bound of points per tile (shared-memory bound)
  int layout;
  int fixed_shift;
  int stagger_ns;   // two-CTAs-per-SM variant of the fused kernel: the second half of the grid starts this much later (phases of the two CTAs of an SM apart)
};

void plan_bins(int n_bins, int precision, int* G_out, int* L_out, int* chunks_out);
int fused_ctas_per_sm(int n_bins, int precision);
cudaError_t launch_scan_moments_fused(ScanParams& P, int n_jobs, int ctas_per_job, int precision, cudaStream_t stream);
bool tc_path_supported(int n_bins, int precision, int fixed_shift);
cudaError_t launch_scan_moments_tc(ScanParams& P, int n_jobs, int ctas_per_job, cudaStream_t stream);
cudaError_t launch_moments_from_resp(const ScanParams& P, int ctas, int G, cudaStream_t stream);

struct FinalizeParams {
  const double* parts;
  int n_parts;
  int64_t part_stride, job_stride;
  const gcl_job* jobs;
  gcl_job job;
  int n_bins, nmom;
  const double* bin_dirs;
  double tau;
  double origin[3];
  double eps_psd, eps_mass;
  int finalize;
  int from_resp;  // standalone moment match: no resample / deskew / soft-assign scalars
  double* reduced;
  double* stats;
};
cudaError_t launch_reduce_finalize(const FinalizeParams& P, int n_jobs, cudaStream_t stream, int* launches);
constexpr int kFinThreads = 512;
constexpr int kRedSmem = 5120;   // doubles of the reduced record kept in shared memory (19 * 256 + 8 fits)
struct FinalizeShared { double buf[kFinThreads / 32]; double pad[2]; double redrec[kRedSmem]; };


struct EvidenceParams {
  const double* stats;
  int64_t stats_stride;
  const double *map_S_dir, *map_S_dir_scatter, *map_N_dir, *map_N_pos, *map_derived;
  const double* pose_pred;
  const double* R_hat_override;
  int n_bins;
  double eps_psd, eps_mass;
  int do_rotation, do_translation;
  double* evidence;
};
cudaError_t launch_pose_evidence(const EvidenceParams& P, int n_jobs, cudaStream_t stream);
struct ChainParams {
  double* stats;            // n_scans * n_hyp stats records (Sigma_p is rewritten when use_sigma_meas)
  int64_t stats_stride;
  double *map_S_dir, *map_S_dir_scatter, *map_N_dir, *map_N_pos, *map_sum_p, *map_sum_ppT;  // updated in place
  double* workspace;        // chain_workspace_doubles(n_scans, n_hyp, n_bins) doubles
  const double* pose_pred;  // (n_scans * n_hyp, 6)
  const double* pose_update;  // (n_scans, 6) pose used for the map update, or nullptr
  double* iw_state;         // [nu, Psi(9)] LiDAR measurement-noise block, or nullptr
  double* evidence;         // (n_scans * n_hyp) records
  int n_scans, n_hyp, n_bins, use_sigma_meas;
  int chain_smem;           // single-launch chain: map-derived statistics live in dynamic shared memory
  int phase = 0;            // staged stream: 0 = all, 1 = map-dependent part, 2 = IW-dependent part (launch_chain)
  double forgetting, iw_rho, iw_nu_weak_add, iw_nu_max, eps_psd, eps_mass;
  // single-launch host staging (one scan, gcl_scan_evidence_begin_host): the kernel itself carries the small blocks across PCIe
  // -- inputs out of page-locked host memory first, outputs back at the end, then a completion flag the host polls -- so that
  // the evidence step is ONE enqueue instead of H2D + D2D + kernel + D2H + stream synchronisation.  All zero: not used.
  const double* stage_in_src = nullptr; double* stage_in_dst = nullptr; int stage_in_doubles = 0;      // pinned host -> device
  const double* stats_src = nullptr; int stats_doubles = 0;                                            // stage-1 records -> stats
  const double* stage_out_src = nullptr; double* stage_out_dst = nullptr; int stage_out_doubles = 0;   // device -> pinned host
  int out_off_iw = 0, out_off_stats = 0, out_off_evid = 0;   // offsets (doubles) of the IW / stats / evidence blocks in the output
  double pose_inline[30] = {0};   // pose_pred (n_hyp x 6) [+ pose_update (6)] by value: no PCIe read for them (n_hyp <= 4)
  int pose_inline_n = 0, pose_inline_update = 0;
  unsigned long long* done_flag = nullptr; unsigned long long done_value = 0;                          // pinned host
};
cudaError_t launch_chain(const ChainParams& P, cudaStream_t stream, int* launches);
size_t chain_workspace_doubles(int n_scans, int n_hyp, int n_bins);
cudaError_t launch_map_derived(const double* S_dir, const double* N_dir, const double* N_pos, const double* sum_p,
                               const double* sum_ppT, int n_bins, double eps_mass, double eps_psd, double* out,
                               cudaStream_t stream);
cudaError_t launch_map_update(const double* stats, const double* R, const double* t, int n_bins, double forgetting,
                              double* S_dir, double* S_dir_scatter, double* N_dir, double* N_pos, double* sum_p,
                              double* sum_ppT, cudaStream_t stream);

struct IngestParams {
  const unsigned char* data;   // raw PointCloud2 bytes (device)
  int64_t n_points;
  int point_step;
  int off[5], dtype[5];        // x, y, z, ring, time (off[4] < 0: no per-point time)
  double header_stamp;
  double sentinel, sigma, min_r, max_r, weight_floor;
  int transform;
  double R[9], t[3];
  double* points_out; double* timestamps_out; double* weights_out;
  unsigned char* ring_out; unsigned char* tag_out;
  float4* packed_xyzt; float* packed_w; double packed_t0;
  int* flag;                   // device scratch, 2 ints: "per-point time is in nanoseconds", ticket counter of pass 1
  int flag_is_zero;            // the caller already zeroed flag[0..1] (e.g. as part of an upload): no memset
  // message -> evidence path (gcl_scan_submit_msg_host): everything below may stay zero for the stand-alone operator
  int compact_stride;          // > 1: points / times are written only for raw rows k * stride, at index k (weights: every row)
  double* bounds_parts;        // device scratch, 2 doubles per CTA of pass 1: extrema of the raw per-point time
  double* bounds_out;          // device, 5 doubles [scan start, scan end, min t, max t, unit flag] (backend_node.py:1776-1786)
  gcl_job* bounds_jobs; int bounds_n_jobs;   // job records whose time window is patched with the bounds
  const double* bounds_dev;    // decode pass: packed_t0 is read from here (the scan start computed on the device)
  int defer_decode;            // launch_ingest stops after the bounds; launch_ingest_decode runs the decode pass
};
cudaError_t launch_ingest(const IngestParams& P, cudaStream_t stream, int* launches);
cudaError_t launch_ingest_decode(const IngestParams& P, cudaStream_t stream, int* launches);

struct ImuParams {
  const double *stamps, *gyro, *accel, *weights;  // (M,), (M,3), (M,3), optional (M,) -- NULL: window weights computed here
  int M;
  double t0, t1, sigma;                           // window of smooth_window_weights (used when weights == NULL)
  const double* bounds_dev;                       // optional: [t0, t1] read from device memory instead (scan bounds of the ingest)
  const double *rotvec_start, *gyro_bias, *accel_bias;  // (H,3) each
  double gravity[3];
  double trans_scale;                             // 0 = rotation-only deskew, 1 = full twist
  double* out;                                    // (H, GCL_IMU_DOUBLES)
  double* weights_out;                            // optional (M,)
  gcl_job* jobs_out;                              // optional (H): xi_body is written into the jobs
};
cudaError_t launch_imu_preintegrate(const ImuParams& P, int n_hyp, cudaStream_t stream);

// standalone operators (gcl_ops.cu)
cudaError_t op_point_budget(const double* pts, const double* ts, const double* w, const uint8_t* ring, const uint8_t* tag,
                            int64_t n, int cap, double eps, double* pts_out, double* ts_out, double* w_out,
                            uint8_t* ring_out, uint8_t* tag_out, double* scalars, double* ws, cudaStream_t st, int* launches);
cudaError_t op_deskew(const double* pts, const double* ts, const double* w, int64_t n, double t0, double t1, const double* xi,
                      double eps, double* pts_out, double* w_out, double* scalars, double* ws, cudaStream_t st, int* launches);
cudaError_t op_ray_dirs(const double* pts, int64_t n, const double* o, double eps, double* out, cudaStream_t st, int* launches);
cudaError_t op_soft_assign(const double* dirs, int64_t n, const double* bins, int B, double tau, double eps, double* resp,
                           double* scalars, double* ws, cudaStream_t st, int* launches);
cudaError_t op_kappa(const double* R, int64_t n, double eps_r, int d, double r0, double tau, double* out, cudaStream_t st, int* launches);
cudaError_t op_atlas(int B, double eps, double* out, cudaStream_t st, int* launches);
cudaError_t op_pack(const double* pts, const double* ts, const double* w, int64_t n, double t0, void* xyzt, float* wf,
                    cudaStream_t st, int* launches);

// offsets inside a stats record
struct StatsLayout {
  int B;
  __host__ __device__ explicit StatsLayout(int b) : B(b) {}
  __host__ __device__ int N() const { return 0; }
  __host__ __device__ int s_dir() const { return B; }
  __host__ __device__ int S_scatter() const { return 4 * B; }
  __host__ __device__ int p_bar() const { return 13 * B; }
  __host__ __device__ int Sigma_p() const { return 16 * B; }
  __host__ __device__ int kappa() const { return 25 * B; }
  __host__ __device__ int sum_p() const { return 26 * B; }
  __host__ __device__ int sum_ppT() const { return 29 * B; }
  __host__ __device__ int scalars() const { return 38 * B; }
  __host__ __device__ int total() const { return 38 * B + GCL_STATS_SCALARS; }
};

}  // namespace gcl
