// gcl_scan_tc.cu -- K1 in fp32c mode with the moment accumulation on the 5th-generation tensor cores (tcgen05 + tensor memory).
//
// Same contract as scan_moments_fused_kernel (gcl_scan_fused.cu): one pass over the raw points fuses the PointBudget gather, the
// constant-twist deskew + window weights, ray directions, the softmax responsibilities and the per-bin sums
//   M_b[19] = sum_i w_i r_ib [1, d, ddT(6), p, ppT(6)]                         (archive/legacy_operators/binning.py:157-173)
// and writes one fp64 partial record per CTA.  What changes is WHERE the 19 multiply-adds per (point, bin) pair run.  They are a
// dense contraction over the points,
//   M[moment][bin] += sum_i Phi[moment][i] * E[i][bin],      Phi_i = (w_i / Z_i) [1, d, ddT, p, ppT],  E[i][b] = exp(d_i.b/tau - 1/tau),
// (the "contraction depth 3" of the north star is the similarity dot, not this sum), so they go to the tensor
// cores while the CUDA cores compute the next tile's exponentials:
//   A = Phi^T, 128 rows x 32 points (K-major, SWIZZLE_128B): rows 32q + j (j < 5) hold the tf32-representable head of moment 5q + j,
//       rows 32q + 8 + j its remainder (value - head); all other rows are zero.  A bf16 copy of the full value sits in the head rows
//       of a second tile (64-byte rows, SWIZZLE_64B).
//   B = E tile, 256 bins x 32 points, K-major: head of e as tf32 (SWIZZLE_128B), remainder e - head as bf16 (SWIZZLE_64B).
//   D = 128 x 256 fp32 in tensor memory:  4 x tcgen05.mma kind::tf32 (K = 8) give  Phi_head E_head  (rows 32q + j) and
//       Phi_rem E_head (rows 32q + 8 + j);  2 x kind::f16 (K = 16) add  Phi E_rem  onto the head rows.  The dropped term
//       Phi_rem E_rem is 2^-22 relative.  (The tensor core truncates tf32 operands -- tools/microbench/umma_probe.cu -- so the
//       heads are cut with a mask and the remainders are exact.)
// Every 8 sub-tiles (256 points) the accumulation moves to the other half of tensor memory and the finished half is drained
// (tcgen05.ld.16x256b puts head row and remainder row of one moment in the same thread) into fp64 accumulators that live in
// registers for the whole kernel: each (moment, bin) has one owner thread, so there is no shared-memory record, no atomics, and the
// partial record is written once at the end.  Two shared-memory stages: the tensor core works on sub-tile t while the CTA computes
// sub-tile t + 1; a stage is reused when the tcgen05.commit of its last use has arrived on its mbarrier.
//
// Measured instruction cost of the six MMAs of one 32-point sub-tile at 256 bins: 769 cycles = 24 cycles per point
// (tools/microbench/umma_probe_t.cu, profiles/r2_umma_probe.md), against ~47 cycles per point of FMA-pipe time for the same sums on
// the CUDA cores -- and it runs beside phase A instead of after it.
#include "gcl_internal.cuh"
#include "gcl_point.cuh"

namespace gcl {

constexpr int kTcTeam = 256;     // threads of one producer team (8 warps = the 8 quads of a sub-tile)
constexpr int kTcProd = 512;     // two teams: team t takes the sub-tiles t, t + 2, ... and owns shared-memory stage t.  Everything a warp does
                                 // per sub-tile is a chain of dependent latencies (shuffles, reciprocal, fences, barrier round trips): with one
                                 // team (2 warps per scheduler) a sub-tile cost ~1900 cycles before its first exponential
constexpr int kTcNT = 544;       // + one warp whose lane 0 issues the MMAs (a tcgen05.mma issue blocks its thread while the queue is
                                 // full -- measured 830 cycles per sub-tile -- so it must not be a thread the others wait for)
constexpr int kTcMacroMax = 512; // points per front-end pass: 512 (packed 20-byte rows) or 256 (40-byte float64 rows: same staging bytes)
constexpr int kTcSub = 32;       // points per MMA sub-tile
constexpr int kTcGroupLog2 = 3;  // 2^3 = 8 sub-tiles (256 points, 48 instructions) accumulate in one half of tensor memory between drains:
                                 // the fp32 accumulation of the tensor core truncates, and the bias grows with the number of products an
                                 // accumulator has absorbed (16 sub-tiles: 6.8e-6 of the largest entry against the fp64 kernel; 8: half that)
constexpr int kTcGroupMask = (1 << kTcGroupLog2) - 1;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t sw128(uint32_t lin) { return lin ^ (((lin >> 7) & 7u) << 4); }
__device__ __forceinline__ uint32_t sw64(uint32_t lin) { return lin ^ (((lin >> 7) & 3u) << 4); }
// shared-memory matrix descriptor (K-major): start address, leading byte offset 16, stride between 8-row groups, version 1, layout
// (2 = SWIZZLE_128B, 4 = SWIZZLE_64B)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t sbo_bytes, int layout) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)layout << 61);
}
// instruction descriptor: D fp32, A and B of format fmt (1 = bf16, 2 = tf32), both K-major, M = 128
__device__ __forceinline__ uint32_t umma_idesc(int fmt, int N) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
               :: "r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_bf16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
               :: "r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t mbar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}\n"
               :: "r"(mbar), "r"(parity) : "memory");
}
// 16 lanes x 256 bits: thread t receives (lane t/4, columns 2(t%4), 2(t%4)+1) in r0, r1 and (lane t/4 + 8, same columns) in r2, r3
__device__ __forceinline__ void tmem_ld_16x256b(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {   // {lo -> bits 0..15, hi -> bits 16..31}, round to nearest
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// shared-memory plan of the dynamic part: two operand stages, the fp64 record, the raw-row staging (host and device agree through this)
struct TcPlan {
  int n_pad;                                   // bins rounded up to what the MMA shape needs (48, 128, 256)
  uint32_t o_a32, o_ehi, o_a16, o_elo, stage;  // inside one stage
  uint32_t o_rec, o_in, total;
};
__host__ __device__ inline TcPlan tc_plan(int nrep) {
  TcPlan p;
  p.n_pad = 16 * nrep;
  p.o_a32 = 0;
  p.o_ehi = 16384;
  p.o_a16 = p.o_ehi + (uint32_t)p.n_pad * 128;
  p.o_elo = p.o_a16 + 8192;
  p.stage = p.o_elo + (uint32_t)p.n_pad * 64;
  p.stage = (p.stage + 1023u) & ~1023u;
  p.o_rec = 2 * p.stage;                                    // [19][n_pad] doubles: this CTA's sums, carried in fp64
  p.o_in = p.o_rec + (uint32_t)GCL_NM * p.n_pad * 8;
  p.total = p.o_in + kTcMacroMax * 20 + 1024;               // 512 packed rows = 256 float64 rows; + slack for the 1024-byte alignment
  return p;
}
__device__ __forceinline__ void bar_producers() { asm volatile("bar.sync 1, %0;" :: "n"(kTcProd) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t mbar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(mbar) : "memory"); }

#ifdef GCL_K1_TRACE
__device__ long long g_tc_trace[2][12];
#define TCT(i) do { if ((tid == 0 || tid == 64) && blockIdx.x == 1 && blockIdx.y == 0) { const long long n_ = clock64(); tc_acc[i] += n_ - tc_t; tc_t = n_; } } while (0)
extern "C" int gcl_debug_tc_trace(long long* out) { return (int)cudaMemcpyFromSymbol(out, g_tc_trace, sizeof(long long) * 24); }
#else
#define TCT(i) do { } while (0)
#endif

template <int NREP>
__global__ void __launch_bounds__(kTcNT, 1) scan_moments_tc_kernel(const __grid_constant__ ScanParams P) {
  extern __shared__ unsigned char smem_raw_tc[];
  // small arrays are STATIC shared memory: distinct objects, so the compiler may move their loads across the operand-tile stores
  // (with everything carved out of one dynamic array every atlas load waited for the previous bin's stores: phase A ran at a
  // quarter of its issue rate)
  __shared__ float4 s_atlas[256];                  // [bin]: log2(e)/tau * direction
  __shared__ float4 s_dq[kTcMacroMax / 4][3];      // [quad of points][dx, dy, dz]: one float4 = that component of 4 consecutive points
  __shared__ float4 s_pw[kTcMacroMax];             // [point]: {p0, p1, p2, w}
  __shared__ double s_red[kTcNT / 32][GCL_NCERT];
  __shared__ gcl_job s_job;
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) uint64_t s_bar[6];      // operands ready (stage 0, 1), stage consumed (0, 1), tensor-memory half complete (0, 1)
  constexpr int NPAD = 16 * NREP;
  const TcPlan pl = tc_plan(NREP);
  const int B = P.n_bins;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#ifdef GCL_K1_TRACE
  long long tc_acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, tc_t = clock64();
#endif
  unsigned char* smem = smem_raw_tc + ((1024u - (smem_u32(smem_raw_tc) & 1023u)) & 1023u);   // swizzled tiles need 1024-byte alignment

  if (tid < (int)(sizeof(gcl_job) / sizeof(int)))
    reinterpret_cast<int*>(&s_job)[tid] = reinterpret_cast<const int*>(P.jobs ? &P.jobs[blockIdx.y] : &P.job)[tid];
  // operand tiles and the fp64 record start as zeros (tile rows that are never written: unused moments, bins beyond n_bins, points
  // beyond a partial tile)
  for (uint32_t i = tid; i < pl.o_in / 16; i += kTcNT) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  constexpr uint32_t kCols = (2 * NPAD <= 32) ? 32u : (2 * NPAD <= 64) ? 64u : (2 * NPAD <= 128) ? 128u : (2 * NPAD <= 256) ? 256u : 512u;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&s_tmem)), "r"(kCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    mbar_init(smem_u32(&s_bar[0]), kTcTeam); mbar_init(smem_u32(&s_bar[1]), kTcTeam);     // every thread of the stage's team arrives
    for (int i = 2; i < 6; ++i) mbar_init(smem_u32(&s_bar[i]), 1);                         // tcgen05.commit arrives
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const double cscale_d = 1.4426950408889634 / P.tau;                // logits in octaves
  const float cscale = (float)cscale_d;
  for (int b = tid; b < 256; b += kTcNT)
    s_atlas[b] = (b < B) ? make_float4((float)(P.bin_dirs[b * 3] * cscale_d), (float)(P.bin_dirs[b * 3 + 1] * cscale_d),
                                       (float)(P.bin_dirs[b * 3 + 2] * cscale_d), 0.f)
                         : make_float4(0.f, 0.f, 0.f, 0.f);
  // per-thread certificate partials (registers): raw weights, selected weights, their squares, deskewed weights, rows, entropy, max r
  double c_wraw = 0.0, c_wsel = 0.0, c_wsel2 = 0.0, c_wdesk = 0.0, c_rows = 0.0, ent_acc = 0.0;
  float mr_acc = 0.f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");        // the zero fill, before the first MMA reads the tiles
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;
  const gcl_job& J = s_job;
  const ScanConst& C = P.c;

  // --- this CTA's slice of the selected rows (as in scan_moments_fused_kernel) -----------------------------------------------------
  const int64_t n_raw = J.n_raw;
  const int64_t cap = J.n_points_cap > 0 ? J.n_points_cap : 1;
  const int64_t stride0 = (n_raw + cap - 1) / cap;                 // ceil(n/cap), point_budget.py:160
  const int64_t stride = stride0 > 1 ? stride0 : 1;
  const int64_t n_sel_total = (n_raw + stride - 1) / stride;
  const int64_t sel_begin = J.sel_begin;
  const int64_t sel_end = (J.sel_end < 0 || J.sel_end > n_sel_total) ? n_sel_total : J.sel_end;
  const int64_t rows = sel_end > sel_begin ? sel_end - sel_begin : 0;
  const int64_t rpc = (rows + gridDim.x - 1) / gridDim.x;
  const int64_t lo0 = sel_begin + (int64_t)blockIdx.x * rpc;
  const int64_t lo = lo0 < sel_end ? lo0 : sel_end;
  const int64_t hi = (lo + rpc < sel_end) ? lo + rpc : sel_end;
  double* part = P.partials + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * P.rec_doubles;
  const int macro = (P.layout == GCL_LAYOUT_REF_F64) ? 256 : 512;   // points per front-end pass (one raw-row staging buffer of 10 KB)
  const int n_sub_total = (int)(((hi - lo) / macro) * (macro / kTcSub) + (((hi - lo) % macro) + kTcSub - 1) / kTcSub);
  double* s_in = reinterpret_cast<double*>(smem + pl.o_in);
  double* s_rec = reinterpret_cast<double*>(smem + pl.o_rec);        // [moment][NPAD]

  // --- sum of ALL raw weights of this CTA's raw range (mass preservation, point_budget.py:69) ---------------------------------------
  if (stride > 1 && hi > lo) {
    const int64_t rlo = lo * stride;
    const int64_t rhi = (hi >= n_sel_total) ? n_raw : hi * stride;
    double p0 = 0, p1 = 0, p2 = 0, p3 = 0;
    int64_t r = rlo + tid;
    if (P.layout == GCL_LAYOUT_REF_F64) {
      const double* w = static_cast<const double*>(J.weights);
      for (; r + 3 * kTcNT < rhi; r += 4 * kTcNT) { p0 += w[r]; p1 += w[r + kTcNT]; p2 += w[r + 2 * kTcNT]; p3 += w[r + 3 * kTcNT]; }
      for (; r < rhi; r += kTcNT) p0 += w[r];
    } else {
      const float* w = static_cast<const float*>(J.weights);
      for (; r + 3 * kTcNT < rhi; r += 4 * kTcNT) { p0 += (double)w[r]; p1 += (double)w[r + kTcNT]; p2 += (double)w[r + 2 * kTcNT]; p3 += (double)w[r + 3 * kTcNT]; }
      for (; r < rhi; r += kTcNT) p0 += (double)w[r];
    }
    c_wraw = (p0 + p1) + (p2 + p3);
  }
  TCT(0);   // prologue

  if (warp == kTcProd / 32) {
    // ================================= the MMA warp: lane 0 feeds the tensor core, sub-tile after sub-tile ================================
    if (lane == 0) {
      const uint32_t idesc32 = umma_idesc(2, NPAD), idesc16 = umma_idesc(1, NPAD);
      uint32_t par_full[2] = {0u, 0u};
      for (int T = 0; T < n_sub_total; ++T) {
        const int stage = T & 1;
        mbar_wait(smem_u32(&s_bar[stage]), par_full[stage]);           // the stage's team has written (and fenced) both operand tiles
        par_full[stage] ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d = tmem + (uint32_t)(((T >> kTcGroupLog2) & 1) * NPAD);
        const uint32_t s0 = smem_u32(smem + (size_t)stage * pl.stage);
        uint32_t accf = (T & kTcGroupMask) ? 1u : 0u;                  // first instruction of a group overwrites its half
#if !defined(GCL_TC_DEBUG) || GCL_TC_DEBUG < 1     // debug builds: 1 = no MMAs (commits only), 2 = also no phase-A bin loop, 3 = also no scale stores
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
          mma_tf32(d, umma_desc(s0 + pl.o_a32 + kb * 32, 1024, 2), umma_desc(s0 + pl.o_ehi + kb * 32, 1024, 2), idesc32, accf);
          accf = 1u;
        }
#pragma unroll
        for (int j = 0; j < 2; ++j)
          mma_bf16(d, umma_desc(s0 + pl.o_a16 + j * 32, 512, 4), umma_desc(s0 + pl.o_elo + j * 32, 512, 4), idesc16, 1u);
#endif
        umma_commit(smem_u32(&s_bar[2 + stage]));
        if ((T & kTcGroupMask) == kTcGroupMask || T == n_sub_total - 1) umma_commit(smem_u32(&s_bar[4 + ((T >> kTcGroupLog2) & 1)]));
      }
    }
    __syncwarp();
  } else {
    // ================================= the producers: two teams of 8 warps ===================================================================
    const int team = warp >> 3, wq = warp & 7;       // team t = sub-tiles t, t + 2, ... = stage t; wq = this warp's quad of the sub-tile
    // ownership of the fp64 sums: warp = (quarter of the tensor-memory lanes, part of the bins); inside the 16x256b fragment thread t
    // holds moment row t/4 (head) and t/4 + 8 (remainder) for the two bins 2(t%4), 2(t%4) + 1 of every 8-column block
    constexpr int PARTS = (NPAD % 32 == 0) ? 4 : 2;  // a 48-bin record splits in 2 x 24 columns (warps 8..15 own nothing then)
    constexpr int REPS = NPAD / (8 * PARTS);
    const int quarter = warp & 3, bpart = warp >> 2;
    const int kq = lane >> 2, cpair = lane & 3;
    const int my_moment = 5 * quarter + kq;
    const bool owns = kq < (quarter < 3 ? 5 : 4) && bpart < PARTS;

    uint32_t par_empty = 0u, par_drain[2] = {0u, 0u};
    int drained = 0;
    auto drain = [&](int g) {
      const int buf = g & 1;
      mbar_wait(smem_u32(&s_bar[4 + buf]), par_drain[buf]);
      par_drain[buf] ^= 1u;
      if (bpart < PARTS) {                                   // warp-uniform
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t taddr = tmem + ((uint32_t)(32 * quarter) << 16) + (uint32_t)(buf * NPAD + bpart * (NPAD / PARTS));
        uint32_t v[REPS][4];
#pragma unroll
        for (int r = 0; r < REPS; ++r) tmem_ld_16x256b(taddr + 8 * r, v[r]);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (owns) {
          double* rec = s_rec + (size_t)my_moment * NPAD + bpart * (NPAD / PARTS) + 2 * cpair;
#pragma unroll
          for (int r = 0; r < REPS; ++r) {
            double2 a = *reinterpret_cast<double2*>(rec + 8 * r);
            a.x += (double)(__uint_as_float(v[r][0]) + __uint_as_float(v[r][2]));     // head row + remainder row (2^-11 of it)
            a.y += (double)(__uint_as_float(v[r][1]) + __uint_as_float(v[r][3]));
            *reinterpret_cast<double2*>(rec + 8 * r) = a;
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      }
    };

    // raw rows of the macro tile [t0, t0 + macro) -> s_in, asynchronously (one row per thread)
    auto prefetch_tile = [&](int64_t t0) {
      const int64_t srow = t0 + tid;
      if (tid >= macro || srow >= hi) return;
      const int64_t r = srow * stride;                                        // raw row (PointBudget gather)
      const int64_t rp = (J.flags & GCL_JOB_COMPACT_ROWS) ? srow : r;
      if (P.layout == GCL_LAYOUT_REF_F64) {
        const double* pts = static_cast<const double*>(J.points) + rp * 3;
        double* dst = s_in + (size_t)tid * 5;
        cp_async8(dst + 0, pts + 0); cp_async8(dst + 1, pts + 1); cp_async8(dst + 2, pts + 2);
        cp_async8(dst + 3, static_cast<const double*>(J.timestamps) + rp);
        cp_async8(dst + 4, static_cast<const double*>(J.weights) + r);
      } else {
        float* base = reinterpret_cast<float*>(s_in);
        cp_async16(base + (size_t)tid * 4, static_cast<const float4*>(J.points) + rp);
        cp_async4(base + (size_t)kTcMacroMax * 4 + tid, static_cast<const float*>(J.weights) + r);
      }
    };
    prefetch_tile(lo);

    float xi32[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) xi32[k] = (float)J.xi_body[k];
    // phase A: WARP = one quad of points (4 consecutive points of the sub-tile), LANE = the bins lane, lane + 32, ...: a warp carries its
    // quad from the exponentials through the normaliser to the scaled rows without meeting the other warps (no CTA barrier inside a
    // sub-tile; the hand-off to the MMA warp is the mbarrier).  The lane's atlas entries never change: they live in registers.
    // The softmax shift is the same for every point (fixed shift -1/tau): it is folded into the bin's constants, and a bin beyond
    // n_bins gets -1e30 instead, which makes its exponential exactly 0 -- so the bin loop has no branches (only its two stores are
    // predicated) and the compiler interleaves the 8 x 4 independent exponential chains of a lane.
    constexpr int NJ = (NPAD + 31) / 32;
    float abx[NJ], aby[NJ], abz[NJ], ash[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const float4 bb = s_atlas[lane + 32 * j];
      abx[j] = bb.x; aby[j] = bb.y; abz[j] = bb.z; ash[j] = (lane + 32 * j < B) ? -cscale : -1.0e30f;
    }
    const uint32_t off_hi0 = sw128((uint32_t)(lane >> 3) * 1024u + (uint32_t)(lane & 7) * 128u + (uint32_t)wq * 16u);
    const uint32_t off_lo0 = sw64((uint32_t)(lane >> 3) * 512u + (uint32_t)(lane & 7) * 64u + (uint32_t)wq * 8u);
    const float* s_dq_f = reinterpret_cast<const float*>(&s_dq[0][0]);
    // scale: lane = (point of the quad, moment group): point (lane >> 2) & 3, moments g, g + 8, g + 16 with g = (lane & 3) + 4 (lane >> 4)
    const int sc_pt = (lane >> 2) & 3, sc_g = (lane & 3) + 4 * (lane >> 4);
    const double ent_eps = (double)B * C.eps_mass;
    unsigned char* const sb = smem + (size_t)team * pl.stage;          // this team's stage
    const uint32_t bar_full = smem_u32(&s_bar[team]), bar_empty = smem_u32(&s_bar[2 + team]);

    int T0 = 0;                                     // sub-tile index of the macro tile's first sub-tile (always even)
    int my_subs = 0;                                // sub-tiles this team has handed over so far
    for (int64_t tile0 = lo; tile0 < hi; tile0 += macro) {
      const int count = (int)((hi - tile0 < macro) ? hi - tile0 : macro);
      cp_async_wait_all();
      bar_producers();          // raw rows landed; every warp is done with the previous macro tile's s_dq / s_pw
      // ======================================= front: one point per thread =======================================================
      if (tid < macro) {
        float d0 = 0.f, d1 = 0.f, d2 = 0.f;
        float4 pw = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tid < count) {
          const int64_t srow = tile0 + tid;
          double w_sel;
          PointOut32 o;
          if (P.layout == GCL_LAYOUT_REF_F64) {
            const double* in = s_in + (size_t)tid * 5;
            w_sel = in[4];
            point_front32(J, C, xi32, (float)in[0], (float)in[1], (float)in[2], (float)(in[3] - J.scan_start_time), o);
          } else {
            const float* base = reinterpret_cast<const float*>(s_in);
            const float4 xyzt = *reinterpret_cast<const float4*>(base + (size_t)tid * 4);
            w_sel = (double)base[(size_t)kTcMacroMax * 4 + tid];
            point_front32(J, C, xi32, xyzt.x, xyzt.y, xyzt.z, xyzt.w, o);
          }
          const double lam = J.point_lambda ? J.point_lambda[srow] : 1.0;
          const double w_desk = w_sel * (double)o.w_time;
          c_wsel += w_sel; c_wsel2 += w_sel * w_sel; c_wdesk += w_desk; c_rows += 1.0;
          d0 = o.d[0]; d1 = o.d[1]; d2 = o.d[2];
          pw = make_float4(o.p0[0], o.p0[1], o.p0[2], (float)(w_desk * lam));
        }
        float* dq = reinterpret_cast<float*>(&s_dq[0][0]) + (size_t)(tid >> 2) * 12 + (tid & 3);
        dq[0] = d0; dq[4] = d1; dq[8] = d2;
        s_pw[tid] = pw;
      }
      bar_producers();
      prefetch_tile(tile0 + macro);
      TCT(1);   // row wait + front end
      const int n_sub = (count + kTcSub - 1) / kTcSub;
      for (int st = team; st < n_sub; st += 2, ++my_subs) {
        const int T = T0 + st;
        if (my_subs >= 1) { mbar_wait(bar_empty, par_empty); par_empty ^= 1u; }   // this team's previous sub-tile is through the tensor core
        TCT(2);   // stage wait
        if (drained < (T >> kTcGroupLog2) && (T & kTcGroupMask) >= 2) drain(drained++);
        TCT(3);   // drain
        // ===================================== phase A: exponentials of this warp's 4 points x this lane's bins ===========================
        const int quad = st * (kTcSub / 4) + wq;
        const float4 DX = s_dq[quad][0], DY = s_dq[quad][1], DZ = s_dq[quad][2];
        const float2 dx01 = make_float2(DX.x, DX.y), dx23 = make_float2(DX.z, DX.w), dy01 = make_float2(DY.x, DY.y), dy23 = make_float2(DY.z, DY.w);
        const float2 dz01 = make_float2(DZ.x, DZ.y), dz23 = make_float2(DZ.z, DZ.w);
        float2 Z01 = make_float2(0.f, 0.f), Z23 = Z01, S01 = Z01, S23 = Z01;
        float m0 = -3.0e38f, m1 = m0, m2 = m0, m3 = m0;
        {
          unsigned char* ehi = sb + pl.o_ehi + off_hi0;
          unsigned char* elo = sb + pl.o_elo + off_lo0;
#if defined(GCL_TC_DEBUG) && GCL_TC_DEBUG >= 2
#pragma unroll
          for (int j = 0; j < 0; ++j) {
#else
#pragma unroll
          for (int j = 0; j < NJ; ++j) {
#endif
            const float2 bx = make_float2(abx[j], abx[j]), by = make_float2(aby[j], aby[j]), bz = make_float2(abz[j], abz[j]);
            const float2 sh = make_float2(ash[j], ash[j]);
            const float2 x01 = ffma2(dx01, bx, ffma2(dy01, by, ffma2(dz01, bz, sh)));
            const float2 x23 = ffma2(dx23, bx, ffma2(dy23, by, ffma2(dz23, bz, sh)));
            const float2 e01 = make_float2(fast_exp2(x01.x), fast_exp2(x01.y));
            const float2 e23 = make_float2(fast_exp2(x23.x), fast_exp2(x23.y));
            const uint32_t h0 = __float_as_uint(e01.x) & 0xFFFFE000u, h1 = __float_as_uint(e01.y) & 0xFFFFE000u;
            const uint32_t h2 = __float_as_uint(e23.x) & 0xFFFFE000u, h3 = __float_as_uint(e23.y) & 0xFFFFE000u;
            const float2 l01 = fadd2(e01, make_float2(-__uint_as_float(h0), -__uint_as_float(h1)));
            const float2 l23 = fadd2(e23, make_float2(-__uint_as_float(h2), -__uint_as_float(h3)));
            Z01 = fadd2(Z01, e01); Z23 = fadd2(Z23, e23);
            S01 = ffma2(e01, x01, S01); S23 = ffma2(e23, x23, S23);        // a masked bin adds 0 * -1e30 = -0
            m0 = fmaxf(m0, x01.x); m1 = fmaxf(m1, x01.y); m2 = fmaxf(m2, x23.x); m3 = fmaxf(m3, x23.y);
            if (NPAD % 32 == 0 || j < NJ - 1 || lane + 32 * j < NPAD) {     // only the last step of a 48-row tile has rows to skip
              *reinterpret_cast<uint4*>(ehi + j * 4096) = make_uint4(h0, h1, h2, h3);
              *reinterpret_cast<uint2*>(elo + j * 2048) = make_uint2(pack_bf16x2(l01.x, l01.y), pack_bf16x2(l23.x, l23.y));
            }
          }
        }
        // ---- sums over the 32 lanes (= all bins) of {Z0..3, S0..3} and max of {m0..3}: reduce-scatter (each step halves what a lane
        //      carries), 15 shuffles instead of 60: afterwards lane l holds v[(l >> 2) & 7] (Z of point l >> 2 for l < 16, S of point
        //      (l >> 2) & 3 for l >= 16) and the maximum of point (l >> 3) & 3
        float sumv, maxv;
        {
          float v[8] = {Z01.x, Z01.y, Z23.x, Z23.y, S01.x, S01.y, S23.x, S23.y};
          const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
          float w4[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {                         // keep v[i] (lanes < 16) or v[4 + i] (lanes >= 16)
            const float keep = b4 ? v[4 + i] : v[i], give = b4 ? v[i] : v[4 + i];
            w4[i] = keep + __shfl_xor_sync(0xffffffffu, give, 16);
          }
          float w2[2];
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const float keep = b3 ? w4[2 + i] : w4[i], give = b3 ? w4[i] : w4[2 + i];
            w2[i] = keep + __shfl_xor_sync(0xffffffffu, give, 8);
          }
          {
            const float keep = b2 ? w2[1] : w2[0], give = b2 ? w2[0] : w2[1];
            sumv = keep + __shfl_xor_sync(0xffffffffu, give, 4);
          }
          sumv += __shfl_xor_sync(0xffffffffu, sumv, 2);
          sumv += __shfl_xor_sync(0xffffffffu, sumv, 1);
          float u2[2];
          {
            const float k0 = b4 ? m2 : m0, g0 = b4 ? m0 : m2, k1 = b4 ? m3 : m1, g1 = b4 ? m1 : m3;
            u2[0] = fmaxf(k0, __shfl_xor_sync(0xffffffffu, g0, 16));
            u2[1] = fmaxf(k1, __shfl_xor_sync(0xffffffffu, g1, 16));
          }
          {
            const float keep = b3 ? u2[1] : u2[0], give = b3 ? u2[0] : u2[1];
            maxv = fmaxf(keep, __shfl_xor_sync(0xffffffffu, give, 8));
          }
          maxv = fmaxf(maxv, __shfl_xor_sync(0xffffffffu, maxv, 4));
          maxv = fmaxf(maxv, __shfl_xor_sync(0xffffffffu, maxv, 2));
          maxv = fmaxf(maxv, __shfl_xor_sync(0xffffffffu, maxv, 1));
        }
        TCT(4);   // phase A
        // ===================================== scale: Z_i, certificate partials, the Phi tile ==========================================
        {
          // v index held by lane l: bit4 -> +4 (S instead of Z), bit3 -> +2, bit2 -> +1: lanes 0..15 hold Z of point (l >> 2),
          // lanes 16..31 hold S of point (l >> 2) & 3; the maximum of point p sits in the lanes with (bit4, bit3) = (p >> 1, p & 1)
          const float Zp = __shfl_sync(0xffffffffu, sumv, sc_pt << 2);
          const int p = 4 * wq + sc_pt;                         // point of the sub-tile
          const float4 bq = s_pw[st * kTcSub + p];
          const float* dqp = s_dq_f + (size_t)quad * 12 + sc_pt;
          const float a_x = dqp[0], a_y = dqp[4], a_z = dqp[8];
          const float invZ = __frcp_rn(Zp);
          const float wn = bq.w * invZ;
          const float xmax = __shfl_sync(0xffffffffu, maxv, ((sc_pt >> 1) << 4) | ((sc_pt & 1) << 3));
          {
            // one lane per point holds its S (lanes 16, 20, 24, 28); no branch: the other lanes add 0 / keep their maximum
            // H_i = -sum_b r log(r + eps) = log Z - S1/Z - B*eps + O(B eps^2 / r)   (binning.py:70-72); S1 in octaves
            const bool mine = lane >= 16 && (lane & 3) == 0 && st * kTcSub + p < count;
            const float h = __logf(Zp) - 0.6931471805599453f * sumv * invZ;
            ent_acc += mine ? (double)h - ent_eps : 0.0;
            mr_acc = fmaxf(mr_acc, mine ? fast_exp2(xmax) * invZ : 0.f);
          }
          unsigned char* a32 = sb + pl.o_a32;
          unsigned char* a16 = sb + pl.o_a16;
          const float wd0 = wn * a_x, wd1 = wn * a_y, wd2 = wn * a_z, wp0 = wn * bq.x, wp1 = wn * bq.y, wp2 = wn * bq.z;
          // moments sc_g, sc_g + 8, sc_g + 16 by selection (no branches: 8 lane groups, 3 rounds)
          const float r0[8] = {wn, wd0, wd1, wd2, wd0 * a_x, wd0 * a_y, wd0 * a_z, wd1 * a_y};
          const float r1[8] = {wd1 * a_z, wd2 * a_z, wp0, wp1, wp2, wp0 * bq.x, wp0 * bq.y, wp0 * bq.z};
          const float r2[8] = {wp1 * bq.y, wp1 * bq.z, wp2 * bq.z, 0.f, 0.f, 0.f, 0.f, 0.f};
          float vv[3] = {r0[0], r1[0], r2[0]};
#pragma unroll
          for (int i = 1; i < 8; ++i) { if (sc_g == i) { vv[0] = r0[i]; vv[1] = r1[i]; vv[2] = r2[i]; } }
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const int k = sc_g + 8 * j;
#if defined(GCL_TC_DEBUG) && GCL_TC_DEBUG >= 3
            if (k < 0) {
#else
            if (k < GCL_NM) {
#endif
              const float v = vv[j];
              const uint32_t hbits = __float_as_uint(v) & 0xFFFFE000u;
              const float rem = v - __uint_as_float(hbits);
              const uint32_t row = (uint32_t)(32 * (k / 5) + (k % 5));
              *reinterpret_cast<uint32_t*>(a32 + sw128((row >> 3) * 1024u + (row & 7u) * 128u + (uint32_t)p * 4u)) = hbits;
              *reinterpret_cast<float*>(a32 + sw128(((row + 8) >> 3) * 1024u + ((row + 8) & 7u) * 128u + (uint32_t)p * 4u)) = rem;
              const uint32_t pk = pack_bf16x2(v, 0.f);
              *reinterpret_cast<unsigned short*>(a16 + sw64((row >> 3) * 512u + (row & 7u) * 64u + (uint32_t)p * 2u)) = (unsigned short)(pk & 0xFFFFu);
            }
          }
        }
        TCT(6);   // scale
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // this thread's st.shared -> visible to the tensor core's async proxy
        mbar_arrive(bar_full);                                           // ... and handed to the MMA warp (the team's 256 arrivals complete the phase)
        TCT(7);   // proxy fence + arrive
      }
      T0 += n_sub;
    }
    // --- whatever has not been drained yet (the MMA warp commits the last, possibly partial, group itself) ---------------------------------
    const int groups = (T0 + kTcGroupMask) >> kTcGroupLog2;
    while (drained < groups) drain(drained++);
  }

  // --- the fp64 record of this CTA -> its partial record; certificate scalars (as in scan_moments_fused_kernel) ----------------------------
  double cert[GCL_NCERT] = {c_wraw, c_wsel, c_wsel2, c_wdesk, ent_acc, (double)mr_acc, c_rows, 0.0};
#pragma unroll
  for (int k = 0; k < GCL_NCERT; ++k) {
    double v = cert[k];
    for (int o = 16; o > 0; o >>= 1) {
      const double other = __shfl_xor_sync(0xffffffffu, v, o);
      v = (k == GCL_CS_MAX_RESP) ? fmax(v, other) : v + other;
    }
    if (lane == 0) s_red[warp][k] = v;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  for (int i = tid; i < GCL_NM * B; i += kTcNT) {
    const int k = i / B, b = i - k * B;
    part[i] = s_rec[(size_t)k * NPAD + b];
  }
  if (tid < GCL_NCERT) {
    const int src = (tid == GCL_CS_SUM_W_RAW && stride == 1) ? GCL_CS_SUM_W_SEL : tid;
    double v = s_red[0][src];
    for (int w = 1; w < kTcNT / 32; ++w) v = (src == GCL_CS_MAX_RESP) ? fmax(v, s_red[w][src]) : v + s_red[w][src];
    part[(size_t)GCL_NM * B + tid] = v;
  }
#ifdef GCL_K1_TRACE
  TCT(11);  // tail: last drains, write-out, certificates
  if ((tid == 0 || tid == 64) && blockIdx.x == 1 && blockIdx.y == 0)
    for (int i = 0; i < 12; ++i) g_tc_trace[tid == 0 ? 0 : 1][i] = tc_acc[i];
#endif
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(kCols) : "memory");
}

// ------------------------------------------------------------------------------------------------------------------------------------
bool tc_path_supported(int n_bins, int precision, int fixed_shift) {
  // GCL_K1_TC: 0 = never, 2 = whenever the kernel can (any bin count up to 256), default = where it wins.  The instruction stream of the
  // tensor core costs the same ~21-24 cycles per point at 48, 128 and 256 bins (profiles/r2_umma_probe.md), the CUDA-core sums scale with
  // the bins: measured cross-over between 128 bins (CUDA cores 191 us, tensor cores 213 us per 1 M points) and 256 (370 / 300 us)
  static const int mode = getenv("GCL_K1_TC") ? atoi(getenv("GCL_K1_TC")) : 1;
  if (mode == 0 || precision != GCL_F32C || n_bins < 1 || n_bins > 256 || !(fixed_shift & 1)) return false;
  return mode == 2 || n_bins > 128;
}

template <int NREP>
static cudaError_t launch_tc_one(const ScanParams& P, dim3 grid, cudaStream_t stream) {
  auto kern = scan_moments_tc_kernel<NREP>;
  static std::atomic<uint64_t> optin{0};
  const size_t smem = tc_plan(NREP).total;
  if (cudaError_t e = optin_dynamic_smem(kern, (int)smem, optin)) return e;
  kern<<<grid, kTcNT, smem, stream>>>(P);
  return cudaGetLastError();
}

cudaError_t launch_scan_moments_tc(ScanParams& P, int n_jobs, int ctas_per_job, cudaStream_t stream) {
  P.lanes = 0; P.tile_points = kTcSub; P.n_chunks = 1;
  const dim3 grid(ctas_per_job, n_jobs);
  if (P.n_bins <= 48) return launch_tc_one<3>(P, grid, stream);
  if (P.n_bins <= 128) return launch_tc_one<8>(P, grid, stream);
  return launch_tc_one<16>(P, grid, stream);
}

}  // namespace gcl
