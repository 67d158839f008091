// gcl_exchange.cu -- the one exchange step of the point-sharded path, fused with what precedes and follows it, over NVLink peer
// memory (one process per GPU).
//
// The reference has no collective at all (single process, single GPU).  When one scan's points are sharded over G GPUs every raw
// bin sum is linear in the points (archive/legacy_operators/binning.py:166-173), so the ranks exchange ONE small record per job
// ([19 x B moments | 8 scalars]: 39 KB at 256 bins) and every rank finalises the same totals.  Round 1 did that as
// reduce_finalize(finalize = 0) -> NCCL all_gather -> reduce_finalize(finalize = 1): two launches plus a collective, ~0.1 ms of
// fixed cost against 0.06 ms of scan_moments at 8 GPUs.  Here it is ONE launch:
//
//   every CTA   fixed-order sum of this rank's per-CTA partial records for its slice of the record (four threads per element,
//               each over a quarter of the partials, combined in a fixed order), stored STRAIGHT INTO EVERY PEER's gather buffer
//               (st.global on the peer's mapped memory: the bytes cross NVLink as they are produced), __threadfence_system(),
//               ticket;
//   last CTA    publishes this rank's step number in every peer's flag word, spins (with a time limit) until all peers' step
//               numbers have arrived in its own flag words, then runs the ordinary K2 body (gcl_finalize_body.cuh) over the
//               gathered records in RANK ORDER -- the same order on every rank, so the statistics are bit-identical on all ranks
//               and independent of any library's reduction order.
//
// Buffers alternate between two parities: rank r can only publish step s + 2 after it has completed step s + 1, which needed every
// peer's step s + 1 record, which each peer wrote (stream order) after finishing step s -- so nobody overwrites a buffer that is
// still being read.  Step numbers only grow; flags are never reset.
// A block holds GCL_EXCHANGE_LANES independent copies of all that (slots, flag words, ticket, step counter): a caller that keeps the
// tails of TWO consecutive scans in flight (gcl_scan_pipe, two tail streams) runs scan k on lane k mod 2, each lane on its own
// stream -- the argument above then holds per lane, and the lanes share nothing.
// The same buffers carry gcl_exchange_all_gather (hypothesis-sharded results).
#include <cstring>

#include "gcl_internal.cuh"
#include "gcl_finalize_body.cuh"

namespace gcl {

struct ExchangeDev {
  unsigned char* peer[GCL_MAX_RANKS];    // every rank's exchange block as mapped into THIS process (peer[rank] = the local block)
  int world, rank;
  unsigned long long step;               // this call's step number on its lane (1, 2, ...)
  int lane;
  size_t slot_bytes;                     // bytes per (lane, parity, rank) slot
  size_t o_flags, o_ticket, o_error;     // offsets in a block (flags, ticket: of this call's lane)
  long long timeout_clocks;
};

__device__ __forceinline__ unsigned char* slot_of(const ExchangeDev& X, int owner, int parity, int src_rank) {
  return X.peer[owner] + ((size_t)(X.lane * 2 + parity) * X.world + src_rank) * X.slot_bytes;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// threads [0, world): publish `step` in the peers' flag words, then wait for the peers' steps in the local ones.  Returns with a
// barrier; *s_ok = 0 if the wait timed out (a peer that never launched: the error word of the local block is set, nothing hangs).
__device__ __forceinline__ void publish_and_wait(const ExchangeDev& X, int* s_ok) {
  const int tid = threadIdx.x;
  if (tid == 0) *s_ok = 1;
  __syncthreads();
  if (tid < X.world) {
    st_release_sys(reinterpret_cast<unsigned long long*>(X.peer[tid] + X.o_flags) + X.rank, X.step);
    const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(X.peer[X.rank] + X.o_flags) + tid;
    const long long t0 = clock64();
    while (ld_acquire_sys(mine) < X.step) {
      if (clock64() - t0 > X.timeout_clocks) {
        *reinterpret_cast<volatile int*>(X.peer[X.rank] + X.o_error) = 1 + tid;
        *s_ok = 0;
        break;
      }
      __nanosleep(64);
    }
  }
  __syncthreads();
}

// ---- point-sharded K2 ------------------------------------------------------------------------------------------------------------
constexpr int kXEl = 128;   // record elements per CTA (x 4 partial ranges = 512 threads)

__global__ void __launch_bounds__(kFinThreads) sharded_reduce_finalize_kernel(const __grid_constant__ FinalizeParams P,
                                                                              const __grid_constant__ ExchangeDev X) {
  __shared__ FinalizeShared sh;
  __shared__ int s_last, s_ok;
  const int tid = threadIdx.x, el = tid & (kXEl - 1), sub = tid >> 7;
  const int job = blockIdx.y;
  const int rec = P.nmom * P.n_bins + GCL_NCERT;
  const int parity = (int)(X.step & 1);
  double* s_part = sh.redrec;                     // [4][kXEl]
  // ---- fixed-order sum of this rank's partials: thread (el, sub) takes the sub-th quarter of the partials; a CTA walks the
  //      chunks blockIdx.x, blockIdx.x + gridDim.x, ... of the record (few CTAs when the launch shares the GPU with the next
  //      scan's moment pass) ---------------------------------------------------------------------------------------------------------
  for (int chunk = blockIdx.x; chunk * kXEl < rec; chunk += gridDim.x) {
    const int e = chunk * kXEl + el;
    double s = 0.0;
    const bool is_max = (e == P.nmom * P.n_bins + GCL_CS_MAX_RESP);
    if (e < rec) {
      const int q = (P.n_parts + 3) >> 2;
      const int p_lo = sub * q, p_hi = min(P.n_parts, p_lo + q);
      const double* src = P.parts + (size_t)job * P.job_stride + e;
      for (int p0 = p_lo; p0 < p_hi; p0 += 8) {
        double v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (p0 + j < p_hi) ? __ldcg(src + (size_t)(p0 + j) * P.part_stride) : 0.0;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (p0 + j < p_hi) s = is_max ? fmax(s, v[j]) : s + v[j];
      }
    }
    s_part[sub * kXEl + el] = s;
    __syncthreads();
    if (e < rec) {
      const double a = s_part[el], b = s_part[kXEl + el], c = s_part[2 * kXEl + el], d = s_part[3 * kXEl + el];
      const double tot = is_max ? fmax(fmax(a, b), fmax(c, d)) : ((a + b) + (c + d));
      // this rank's record goes into slot [parity][rank] of EVERY rank's block (its own included); thread (el, sub) serves
      // peers sub, sub + 4
      for (int r = sub; r < X.world; r += 4)
        reinterpret_cast<double*>(slot_of(X, r, parity, X.rank))[(size_t)job * rec + e] = tot;
    }
    __syncthreads();
  }
  __threadfence_system();
  __syncthreads();
  if (tid == 0) {
    unsigned int* ticket = reinterpret_cast<unsigned int*>(X.peer[X.rank] + X.o_ticket);
    const unsigned int total = gridDim.x * gridDim.y;
    const unsigned int t = atomicAdd(ticket, 1u);
    s_last = (t == total - 1);
    if (s_last) *ticket = 0;                       // ready for the next launch (stream order)
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence_system();
  publish_and_wait(X, &s_ok);
  if (!s_ok) return;
  // ---- all records are here: the ordinary K2 over the gathered records, in rank order -------------------------------------------------
  FinalizeParams F = P;
  F.parts = reinterpret_cast<const double*>(slot_of(X, X.rank, parity, 0));
  F.n_parts = X.world;
  F.part_stride = (int64_t)(X.slot_bytes / sizeof(double));
  F.job_stride = rec;
  F.finalize = 1;
  for (int j = 0; j < (int)gridDim.y; ++j) {
    reduce_finalize_job(F, j, sh);
    __syncthreads();
  }
}

// ---- all-gather of a small device buffer (hypothesis-sharded evidence records) -----------------------------------------------------------
__global__ void __launch_bounds__(kFinThreads) exchange_all_gather_kernel(const __grid_constant__ ExchangeDev X, const uint4* __restrict__ src,
                                                                          size_t n16, uint4* __restrict__ dst) {
  __shared__ int s_last, s_ok;
  const int tid = threadIdx.x;
  const int parity = (int)(X.step & 1);
  // CTA c copies the source into the slots of peers c, c + gridDim.x, ...
  for (int r = blockIdx.x; r < X.world; r += gridDim.x) {
    uint4* out = reinterpret_cast<uint4*>(slot_of(X, r, parity, X.rank));
    for (size_t i = tid; i < n16; i += kFinThreads) out[i] = src[i];
  }
  __threadfence_system();
  __syncthreads();
  if (tid == 0) {
    unsigned int* ticket = reinterpret_cast<unsigned int*>(X.peer[X.rank] + X.o_ticket);
    const unsigned int t = atomicAdd(ticket, 1u);
    s_last = (t == gridDim.x - 1);
    if (s_last) *ticket = 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence_system();
  publish_and_wait(X, &s_ok);
  if (!s_ok || !dst) return;
  for (int r = 0; r < X.world; ++r) {
    const uint4* in = reinterpret_cast<const uint4*>(slot_of(X, X.rank, parity, r));
    for (size_t i = tid; i < n16; i += kFinThreads) dst[(size_t)r * n16 + i] = __ldcg(in + i);
  }
}

}  // namespace gcl

using namespace gcl;

struct gcl_exchange_s {
  int world = 1, rank = 0, device = 0;
  size_t slot_bytes = 0, block_bytes = 0, o_flags = 0, o_ticket = 0, o_error = 0;
  unsigned char* local = nullptr;
  unsigned char* peer[GCL_MAX_RANKS] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  bool connected = false;
  unsigned long long step[GCL_EXCHANGE_LANES] = {0, 0};
  double timeout_seconds = 2.0;
  int clock_khz = 1965000;
  int lane = 0;                // the lane the next calls run on (gcl_exchange_set_lane)
  int max_ctas = 0;            // 0 = one CTA per 128 record elements; > 0: at most this many CTAs (gcl_exchange_set_max_ctas)
  std::string err;
};

static ExchangeDev dev_view(gcl_exchange x, int lane) {
  ExchangeDev X;
  memset(&X, 0, sizeof(X));
  for (int r = 0; r < x->world; ++r) X.peer[r] = x->peer[r];
  X.world = x->world; X.rank = x->rank; X.step = x->step[lane]; X.lane = lane; X.slot_bytes = x->slot_bytes;
  X.o_flags = x->o_flags + (size_t)lane * 128; X.o_ticket = x->o_ticket + (size_t)lane * 32; X.o_error = x->o_error;
  X.timeout_clocks = (long long)(x->timeout_seconds * 1e3 * (double)x->clock_khz);
  return X;
}

extern "C" {

const char* gcl_exchange_last_error(gcl_exchange x) { return x ? x->err.c_str() : "gcl_exchange: NULL exchange"; }

int gcl_exchange_create(gcl_handle h, int world, int rank, size_t slot_bytes, gcl_exchange* out, unsigned char* ipc_handle_out) {
  if (!h || !out || !ipc_handle_out) return GCL_ERR_ARG;
  if (world < 1 || world > GCL_MAX_RANKS || rank < 0 || rank >= world || slot_bytes == 0) return GCL_ERR_ARG;
  static_assert(sizeof(cudaIpcMemHandle_t) == GCL_EXCHANGE_HANDLE_BYTES, "IPC handle size");
  gcl_exchange x = new gcl_exchange_s();
  x->world = world; x->rank = rank; x->device = h->device;
  x->slot_bytes = (slot_bytes + 255) & ~(size_t)255;
  x->o_flags = 2 * GCL_EXCHANGE_LANES * (size_t)world * x->slot_bytes;   // then GCL_EXCHANGE_LANES x 128 B of flag words (8 ranks x 8 B)
  x->o_ticket = x->o_flags + 256;
  x->o_error = x->o_ticket + 64;
  x->block_bytes = x->o_error + 64;
  cudaSetDevice(x->device);
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&x->local), x->block_bytes);
  if (e == cudaSuccess) e = cudaMemset(x->local, 0, x->block_bytes);
  cudaIpcMemHandle_t hd;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&hd, x->local);
  int khz = 0;
  if (e == cudaSuccess && cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, x->device) == cudaSuccess && khz > 0) x->clock_khz = khz;
  if (e != cudaSuccess) {
    if (x->local) cudaFree(x->local);
    delete x;
    return GCL_ERR_CUDA;
  }
  memcpy(ipc_handle_out, &hd, sizeof(hd));
  x->peer[rank] = x->local;
  if (const char* t = getenv("GCL_EXCHANGE_TIMEOUT_S")) x->timeout_seconds = atof(t) > 0 ? atof(t) : x->timeout_seconds;
  *out = x;
  return GCL_OK;
}

int gcl_exchange_connect(gcl_handle h, gcl_exchange x, const unsigned char* all_handles) {
  if (!h || !x || !all_handles) return GCL_ERR_ARG;
  cudaSetDevice(x->device);
  for (int r = 0; r < x->world; ++r) {
    if (r == x->rank) continue;
    cudaIpcMemHandle_t hd;
    memcpy(&hd, all_handles + (size_t)r * GCL_EXCHANGE_HANDLE_BYTES, sizeof(hd));
    void* p = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      x->err = std::string("gcl_exchange_connect: cudaIpcOpenMemHandle(rank ") + std::to_string(r) + "): " + cudaGetErrorString(e);
      cudaGetLastError();
      return GCL_ERR_CUDA;
    }
    x->peer[r] = static_cast<unsigned char*>(p);
  }
  x->connected = true;
  return GCL_OK;
}

// Tear-down is two collective phases, with a barrier of the caller's between them: every rank first unmaps the peers' blocks
// (gcl_exchange_disconnect), and only when all have done so frees its own (gcl_exchange_destroy) -- a block must not be freed while
// another process still has it mapped.
int gcl_exchange_disconnect(gcl_handle h, gcl_exchange x) {
  if (!h || !x) return GCL_ERR_ARG;
  cudaSetDevice(x->device);
  cudaDeviceSynchronize();
  for (int r = 0; r < x->world; ++r)
    if (r != x->rank && x->peer[r]) { cudaIpcCloseMemHandle(x->peer[r]); x->peer[r] = nullptr; }
  x->connected = false;
  return GCL_OK;
}

void gcl_exchange_destroy(gcl_handle h, gcl_exchange x) {
  if (!x) return;
  gcl_exchange_disconnect(h, x);
  if (x->local) cudaFree(x->local);
  delete x;
}

int gcl_exchange_set_max_ctas(gcl_exchange x, int max_ctas) {
  if (!x || max_ctas < 0) return GCL_ERR_ARG;
  x->max_ctas = max_ctas;
  return GCL_OK;
}

int gcl_exchange_set_lane(gcl_exchange x, int lane) {
  if (!x || lane < 0 || lane >= GCL_EXCHANGE_LANES) return GCL_ERR_ARG;
  x->lane = lane;
  return GCL_OK;
}

int gcl_exchange_error_flag(gcl_exchange x) {
  if (!x) return -1;
  int v = 0;
  cudaSetDevice(x->device);
  if (cudaMemcpy(&v, x->local + x->o_error, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return v;
}

int gcl_sharded_reduce_finalize(gcl_handle h, gcl_exchange x, void* stream, const double* parts, int n_parts, int64_t part_stride,
                                int64_t job_stride, const gcl_job* jobs_dev, const gcl_job* job_host, int n_jobs, int n_bins,
                                const double* bin_dirs, double tau, const double* origin_host, double eps_psd, double eps_mass,
                                double* reduced, double* stats) {
  if (!h || !x) return GCL_ERR_ARG;
  const size_t rec = (size_t)GCL_NM * n_bins + GCL_NCERT;
  if (!(x->connected || x->world == 1)) { x->err = "gcl_sharded_reduce_finalize: exchange not connected"; return GCL_ERR_ARG; }
  if (n_jobs < 1 || n_parts < 1 || n_bins < 1 || n_bins > GCL_MAX_BINS || !parts || !reduced || !stats || !bin_dirs || !origin_host ||
      !(tau > 0.0) || ((jobs_dev != nullptr) == (job_host != nullptr)) || (size_t)n_jobs * rec * sizeof(double) > x->slot_bytes) {
    x->err = "gcl_sharded_reduce_finalize: bad arguments (or n_jobs records do not fit the exchange slot)";
    return GCL_ERR_ARG;
  }
  FinalizeParams P;
  memset(&P, 0, sizeof(P));
  P.parts = parts; P.n_parts = n_parts; P.part_stride = part_stride; P.job_stride = job_stride;
  P.jobs = jobs_dev;
  if (job_host) P.job = *job_host;
  P.n_bins = n_bins; P.nmom = GCL_NM; P.bin_dirs = bin_dirs; P.tau = tau;
  P.origin[0] = origin_host[0]; P.origin[1] = origin_host[1]; P.origin[2] = origin_host[2];
  P.eps_psd = eps_psd; P.eps_mass = eps_mass; P.finalize = 1; P.from_resp = 0;
  P.reduced = reduced; P.stats = stats;
  cudaSetDevice(x->device);
  x->step[x->lane] += 1;
  const ExchangeDev X = dev_view(x, x->lane);
  unsigned chunks = (unsigned)((rec + kXEl - 1) / kXEl);
  if (x->max_ctas > 0 && chunks * (unsigned)n_jobs > (unsigned)x->max_ctas) {
    chunks = (unsigned)x->max_ctas / (unsigned)n_jobs;
    if (chunks < 1) chunks = 1;
  }
  const dim3 grid(chunks, (unsigned)n_jobs);
  sharded_reduce_finalize_kernel<<<grid, kFinThreads, 0, (cudaStream_t)stream>>>(P, X);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { x->err = std::string("gcl_sharded_reduce_finalize: ") + cudaGetErrorString(e); return GCL_ERR_CUDA; }
  h->launches += 1;
  return GCL_OK;
}

int gcl_exchange_all_gather(gcl_handle h, gcl_exchange x, void* stream, const void* src, size_t bytes, void* dst) {
  if (!h || !x || !src) return GCL_ERR_ARG;
  if (!(x->connected || x->world == 1) || bytes == 0 || (bytes & 15) || bytes > x->slot_bytes ||
      (reinterpret_cast<uintptr_t>(src) & 15) || (reinterpret_cast<uintptr_t>(dst) & 15)) {
    x->err = "gcl_exchange_all_gather: bytes must be a non-zero multiple of 16 that fits the slot, pointers 16-byte aligned";
    return GCL_ERR_ARG;
  }
  cudaSetDevice(x->device);
  x->step[x->lane] += 1;
  const ExchangeDev X = dev_view(x, x->lane);
  const int grid = x->world < 4 ? x->world : 4;
  exchange_all_gather_kernel<<<grid, kFinThreads, 0, (cudaStream_t)stream>>>(X, static_cast<const uint4*>(src), bytes / 16, static_cast<uint4*>(dst));
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { x->err = std::string("gcl_exchange_all_gather: ") + cudaGetErrorString(e); return GCL_ERR_CUDA; }
  h->launches += 1;
  return GCL_OK;
}

const void* gcl_exchange_gathered(gcl_exchange x, int src_rank) {
  if (!x || src_rank < 0 || src_rank >= x->world) return nullptr;
  return x->local + ((size_t)(x->lane * 2 + (int)(x->step[x->lane] & 1)) * x->world + src_rank) * x->slot_bytes;
}

}  // extern "C"
