// gcl_pipe.cu -- consecutive scans (point-sharded) or hypothesis batches in flight: the ordering engine behind
// PointShardedEvidence.run_overlapped and the hypothesis-sharded batch.
//
// Per scan the path is three launches -- gcl_scan_moments (O(N), the whole GPU), the exchange + finalisation (O(B), a few CTAs, waits
// for the peers), the evidence chain (O(B), one CTA) -- and at 8 GPUs the O(B) tail (60-70 us of latency) is longer than the moment
// pass of a 1 M-point cloud's shard (43 us in fp32c).  One scan at a time the GPU idles through every tail; with the tail of scan k on
// a second stream beside the moment pass of scan k + 1 the step is bound by the tail; with TWO tail streams (scan k on lane k mod 2
// of the exchange block, see gcl_exchange.cu) it is bound by the moment pass again.  The host side matters as much: the Python layer
// spent 97 us per scan on stream / event bookkeeping (tools/tail_probe.py), more than any of the kernels at 8 GPUs; here a
// submission is one call and ~10 driver calls.
//
// The caller owns all buffers (n_sets sets of partials / reduced / stats / workspace / evidence [/ gathered]); the pipe owns the
// streams and the events that say when a set may be written again.  Memory handed to a submission (jobs, poses, map, scan) must stay
// valid until n_sets further submissions or a join.
#include <cstring>

#include "gcl_internal.cuh"

struct gcl_scan_pipe_s {
  gcl_handle h = nullptr;
  gcl_exchange x = nullptr;
  int mode = 0, n_sets = 2, n_tails = 1, world = 1;
  unsigned long long turn = 0;
  cudaStream_t s_k1 = nullptr, s_tail[GCL_EXCHANGE_LANES] = {nullptr, nullptr};
  cudaEvent_t ev_in = nullptr, ev_ready[GCL_PIPE_MAX_SETS], ev_free[GCL_PIPE_MAX_SETS];
  bool free_valid[GCL_PIPE_MAX_SETS];
  bool tail_used[GCL_EXCHANGE_LANES] = {false, false};
  cudaEvent_t ev_tail[GCL_EXCHANGE_LANES] = {nullptr, nullptr};   // end of the last tail enqueued on each lane
  // optional read-back of a set's results into page-locked host memory at the end of its tail (gcl_scan_pipe_set_readback)
  double* rb_stats[GCL_PIPE_MAX_SETS] = {nullptr};
  double* rb_evidence[GCL_PIPE_MAX_SETS] = {nullptr};
  std::string err;
};

extern "C" {

const char* gcl_scan_pipe_last_error(gcl_scan_pipe p) { return p ? p->err.c_str() : "gcl_scan_pipe: NULL pipe"; }

int gcl_scan_pipe_create(gcl_handle h, gcl_exchange x, int world, int mode, int n_sets, int n_tails, gcl_scan_pipe* out) {
  if (!h || !out || world < 1 || world > GCL_MAX_RANKS || (mode != GCL_PIPE_SUM_RECORDS && mode != GCL_PIPE_GATHER_EVIDENCE) || n_sets < 2 ||
      n_sets > GCL_PIPE_MAX_SETS || n_tails < 1 || n_tails > GCL_EXCHANGE_LANES || n_sets <= n_tails || (mode == GCL_PIPE_SUM_RECORDS && !x) ||
      (world > 1 && !x))
    return GCL_ERR_ARG;
  gcl_scan_pipe p = new gcl_scan_pipe_s();
  p->h = h; p->x = x; p->mode = mode; p->n_sets = n_sets; p->n_tails = n_tails; p->world = world;
  cudaSetDevice(h->device);
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);                        // hi = greatest priority (numerically lowest)
  cudaError_t e = cudaStreamCreateWithPriority(&p->s_k1, cudaStreamNonBlocking, lo);
  for (int t = 0; t < n_tails && e == cudaSuccess; ++t) {
    e = cudaStreamCreateWithPriority(&p->s_tail[t], cudaStreamNonBlocking, hi);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ev_tail[t], cudaEventDisableTiming);
  }
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ev_in, cudaEventDisableTiming);
  for (int i = 0; i < GCL_PIPE_MAX_SETS; ++i) { p->ev_ready[i] = p->ev_free[i] = nullptr; p->free_valid[i] = false; }
  for (int i = 0; i < n_sets && e == cudaSuccess; ++i) {
    e = cudaEventCreateWithFlags(&p->ev_ready[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ev_free[i], cudaEventDisableTiming);
  }
  if (e != cudaSuccess) {
    h->err = std::string("gcl_scan_pipe_create: ") + cudaGetErrorString(e);
    gcl_scan_pipe_destroy(p);
    return GCL_ERR_CUDA;
  }
  *out = p;
  return GCL_OK;
}

void gcl_scan_pipe_destroy(gcl_scan_pipe p) {
  if (!p) return;
  cudaSetDevice(p->h->device);
  if (p->s_k1) { cudaStreamSynchronize(p->s_k1); cudaStreamDestroy(p->s_k1); }
  for (int t = 0; t < GCL_EXCHANGE_LANES; ++t) {
    if (p->s_tail[t]) { cudaStreamSynchronize(p->s_tail[t]); cudaStreamDestroy(p->s_tail[t]); }
    if (p->ev_tail[t]) cudaEventDestroy(p->ev_tail[t]);
  }
  if (p->ev_in) cudaEventDestroy(p->ev_in);
  for (int i = 0; i < GCL_PIPE_MAX_SETS; ++i) {
    if (p->ev_ready[i]) cudaEventDestroy(p->ev_ready[i]);
    if (p->ev_free[i]) cudaEventDestroy(p->ev_free[i]);
  }
  if (p->x) gcl_exchange_set_lane(p->x, 0);
  delete p;
}

#define PIPE_CUDA(call)                                                                                  \
  do {                                                                                                   \
    const cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) { p->err = std::string("gcl_scan_pipe_submit: ") + cudaGetErrorString(e_); return GCL_ERR_CUDA; } \
  } while (0)

int gcl_scan_pipe_submit(gcl_scan_pipe p, void* caller_stream, int wait_for_caller, int set, const gcl_job* jobs_dev,
                         const gcl_job* job_host, int n_jobs, int ctas_per_job, int layout, int precision, const double* bin_dirs, int n_bins,
                         double tau, const double* origin_host, double eps_psd, double eps_mass, const double* map_S_dir,
                         const double* map_S_dir_scatter, const double* map_N_dir, const double* map_N_pos, const double* map_sum_p,
                         const double* map_sum_ppT, const double* pose_pred, double* partials, double* reduced, double* stats,
                         double* workspace, double* evidence, double* gathered) {
  if (!p) return GCL_ERR_ARG;
  if (set < 0 || set >= p->n_sets || !partials || !reduced || !stats || !workspace || !evidence || n_jobs < 1 || ctas_per_job < 1 ||
      (p->mode == GCL_PIPE_GATHER_EVIDENCE && p->world > 1 && !gathered)) {
    p->err = "gcl_scan_pipe_submit: bad arguments";
    return GCL_ERR_ARG;
  }
  gcl_handle h = p->h;
  cudaSetDevice(h->device);
  const int lane = (int)(p->turn % (unsigned long long)p->n_tails);
  cudaStream_t tail = p->s_tail[lane];
  if (wait_for_caller) {   // what the caller enqueued so far (uploads, poses, map) comes first
    PIPE_CUDA(cudaEventRecord(p->ev_in, (cudaStream_t)caller_stream));
    PIPE_CUDA(cudaStreamWaitEvent(p->s_k1, p->ev_in, 0));
    PIPE_CUDA(cudaStreamWaitEvent(tail, p->ev_in, 0));
  }
  if (p->free_valid[set]) PIPE_CUDA(cudaStreamWaitEvent(p->s_k1, p->ev_free[set], 0));   // the tail that read this set is done
  const size_t rec = gcl_partial_record_doubles(n_bins);
  int rc = gcl_scan_moments(h, p->s_k1, jobs_dev, job_host, n_jobs, ctas_per_job, layout, precision, bin_dirs, n_bins, tau, origin_host,
                            eps_mass, partials);
  if (rc != GCL_OK) { p->err = std::string("gcl_scan_pipe_submit: scan_moments: ") + h->err; return rc; }
  PIPE_CUDA(cudaEventRecord(p->ev_ready[set], p->s_k1));
  PIPE_CUDA(cudaStreamWaitEvent(tail, p->ev_ready[set], 0));
  if (p->x) gcl_exchange_set_lane(p->x, lane);
  if (p->mode == GCL_PIPE_SUM_RECORDS) {
    rc = gcl_sharded_reduce_finalize(h, p->x, tail, partials, ctas_per_job, (int64_t)rec, (int64_t)rec * ctas_per_job, jobs_dev, job_host,
                                     n_jobs, n_bins, bin_dirs, tau, origin_host, eps_psd, eps_mass, reduced, stats);
    if (rc != GCL_OK) { p->err = std::string("gcl_scan_pipe_submit: exchange: ") + gcl_exchange_last_error(p->x); return rc; }
  } else {
    rc = gcl_reduce_finalize(h, tail, partials, ctas_per_job, (int64_t)rec, (int64_t)rec * ctas_per_job, jobs_dev, job_host, n_jobs, n_bins,
                             bin_dirs, tau, origin_host, eps_psd, eps_mass, 1, reduced, stats);
    if (rc != GCL_OK) { p->err = std::string("gcl_scan_pipe_submit: reduce_finalize: ") + h->err; return rc; }
  }
  // map-derived statistics + pose evidence: the single-launch chain step on a read-only map (no pose update, no IW state)
  rc = gcl_scan_chain_phase(h, tail, 0, stats, (int64_t)gcl_stats_record_doubles(n_bins), 1, n_jobs, n_bins, const_cast<double*>(map_S_dir),
                            const_cast<double*>(map_S_dir_scatter), const_cast<double*>(map_N_dir), const_cast<double*>(map_N_pos),
                            const_cast<double*>(map_sum_p), const_cast<double*>(map_sum_ppT), workspace, pose_pred, nullptr, 1.0, nullptr, 0,
                            0.0, 0.0, 0.0, eps_psd, eps_mass, evidence);
  if (rc != GCL_OK) { p->err = std::string("gcl_scan_pipe_submit: chain: ") + h->err; return rc; }
  if (p->mode == GCL_PIPE_GATHER_EVIDENCE && p->world > 1) {
    rc = gcl_exchange_all_gather(h, p->x, tail, evidence, (size_t)n_jobs * GCL_EVID_DOUBLES * sizeof(double), gathered);
    if (rc != GCL_OK) { p->err = std::string("gcl_scan_pipe_submit: all_gather: ") + gcl_exchange_last_error(p->x); return rc; }
  }
  if (p->rb_stats[set])     // the caller's page-locked copies of this set's results: part of the tail, ahead of "set free"
    PIPE_CUDA(cudaMemcpyAsync(p->rb_stats[set], stats, (size_t)n_jobs * gcl_stats_record_doubles(n_bins) * sizeof(double), cudaMemcpyDeviceToHost, tail));
  if (p->rb_evidence[set])
    PIPE_CUDA(cudaMemcpyAsync(p->rb_evidence[set], evidence, (size_t)n_jobs * GCL_EVID_DOUBLES * sizeof(double), cudaMemcpyDeviceToHost, tail));
  PIPE_CUDA(cudaEventRecord(p->ev_free[set], tail));
  p->free_valid[set] = true;
  PIPE_CUDA(cudaEventRecord(p->ev_tail[lane], tail));
  p->tail_used[lane] = true;
  p->turn += 1;
  return GCL_OK;
}

int gcl_scan_pipe_set_readback(gcl_scan_pipe p, int set, double* stats_host, double* evidence_host) {
  if (!p || set < 0 || set >= p->n_sets) return GCL_ERR_ARG;
  p->rb_stats[set] = stats_host;
  p->rb_evidence[set] = evidence_host;
  return GCL_OK;
}

int gcl_scan_pipe_wait_set(gcl_scan_pipe p, int set, void* stream) {
  if (!p || set < 0 || set >= p->n_sets) return GCL_ERR_ARG;
  cudaSetDevice(p->h->device);
  if (p->free_valid[set]) {
    const cudaError_t e = cudaStreamWaitEvent((cudaStream_t)stream, p->ev_free[set], 0);
    if (e != cudaSuccess) { p->err = std::string("gcl_scan_pipe_wait_set: ") + cudaGetErrorString(e); return GCL_ERR_CUDA; }
  }
  return GCL_OK;
}

int gcl_scan_pipe_join(gcl_scan_pipe p, void* caller_stream) {
  if (!p) return GCL_ERR_ARG;
  cudaSetDevice(p->h->device);
  for (int t = 0; t < p->n_tails; ++t)
    if (p->tail_used[t]) {
      const cudaError_t e = cudaStreamWaitEvent((cudaStream_t)caller_stream, p->ev_tail[t], 0);
      if (e != cudaSuccess) { p->err = std::string("gcl_scan_pipe_join: ") + cudaGetErrorString(e); return GCL_ERR_CUDA; }
    }
  return GCL_OK;
}

}  // extern "C"
