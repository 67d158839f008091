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
// Every 16 sub-tiles (512 points) the accumulation moves to the other half of tensor memory and the finished half is drained
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

constexpr int kTcNT = 256;       // threads per CTA
constexpr int kTcMacro = 256;    // points per front-end pass
constexpr int kTcSub = 32;       // points per MMA sub-tile
constexpr int kTcGroup = 16;     // sub-tiles accumulated in one half of tensor memory (512 points: 96 fp32 accumulations)
constexpr int kTcIssuer = 224;   // the thread that issues the MMAs (warp 7, lane 0; warp 0 carries the certificates)

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

// shared-memory plan (dynamic part; host and device agree through this)
struct TcPlan {
  int n_pad;                                   // bins rounded up to what the MMA shape needs (48, 128, 256)
  uint32_t o_a32, o_ehi, o_a16, o_elo, stage;  // inside one stage
  uint32_t o_pt, o_atlas, o_z2, o_in, total;
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
  p.o_pt = 2 * p.stage;
  p.o_atlas = p.o_pt + kTcMacro * 32;
  p.o_z2 = p.o_atlas + 256 * 16;
  p.o_in = p.o_z2 + 3 * 8 * kTcSub * 4;
  p.total = p.o_in + kTcMacro * 40 + 1024;     // + slack for the manual 1024-byte alignment of the base
  return p;
}

template <int NREP>
__global__ void __launch_bounds__(kTcNT, 1) scan_moments_tc_kernel(const __grid_constant__ ScanParams P) {
  extern __shared__ unsigned char smem_raw_tc[];
  __shared__ double s_red[kTcNT / 32][GCL_NCERT];
  __shared__ double s_cert[kTcNT * GCL_NCERT];
  __shared__ gcl_job s_job;
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) uint64_t s_bar[4];      // stage 0, stage 1, drain of tensor-memory half 0, half 1
  constexpr int NPAD = 16 * NREP;
  const TcPlan pl = tc_plan(NREP);
  const int B = P.n_bins;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned char* smem = smem_raw_tc + ((1024u - (smem_u32(smem_raw_tc) & 1023u)) & 1023u);   // swizzled tiles need 1024-byte alignment

  if (tid < (int)(sizeof(gcl_job) / sizeof(int)))
    reinterpret_cast<int*>(&s_job)[tid] = reinterpret_cast<const int*>(P.jobs ? &P.jobs[blockIdx.y] : &P.job)[tid];
  // operand tiles start as zeros: rows that are never written (unused moments, bins beyond n_bins, points beyond a partial tile)
  for (uint32_t i = tid; i < 2 * pl.stage / 16; i += kTcNT) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (warp == 0) {
    constexpr uint32_t cols = (2 * NPAD <= 32) ? 32u : (2 * NPAD <= 64) ? 64u : (2 * NPAD <= 128) ? 128u : (2 * NPAD <= 256) ? 256u : 512u;
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&s_tmem)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&s_bar[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
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

  float4* s_pt = reinterpret_cast<float4*>(smem + pl.o_pt);          // [point][2]: {d0, d1, d2, shift} {p0, p1, p2, w}
  float4* s_atlas = reinterpret_cast<float4*>(smem + pl.o_atlas);    // [bin]: log2(e)/tau * direction
  float* s_z2 = reinterpret_cast<float*>(smem + pl.o_z2);            // [3][8 warps][32 points]: partial Z, sum e x, max x
  double* s_in = reinterpret_cast<double*>(smem + pl.o_in);

  const double cscale_d = 1.4426950408889634 / P.tau;                // logits in octaves
  const float cscale = (float)cscale_d;
  for (int b = tid; b < 256; b += kTcNT)
    s_atlas[b] = (b < B) ? make_float4((float)(P.bin_dirs[b * 3] * cscale_d), (float)(P.bin_dirs[b * 3 + 1] * cscale_d),
                                       (float)(P.bin_dirs[b * 3 + 2] * cscale_d), 0.f)
                         : make_float4(0.f, 0.f, 0.f, 0.f);

  double* my_cert = s_cert + tid;
#define CERT(k) my_cert[(k) * kTcNT]
#pragma unroll
  for (int k = 0; k < GCL_NCERT; ++k) CERT(k) = 0.0;

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
    CERT(GCL_CS_SUM_W_RAW) = (p0 + p1) + (p2 + p3);
  }

  // --- ownership of the fp64 accumulators: warp = (quarter of the tensor-memory lanes, half of the bins); inside the 16x256b fragment
  //     thread t holds moment row t/4 (head) and t/4 + 8 (remainder) for the two bins 2(t%4), 2(t%4) + 1 of every 8-column block
  const int quarter = warp & 3, half = warp >> 2;
  const int kq = lane >> 2, cpair = lane & 3;
  const int my_moment = 5 * quarter + kq;
  const bool owns = kq < (quarter < 3 ? 5 : 4);
  double acc[NREP][2];
#pragma unroll
  for (int r = 0; r < NREP; ++r) { acc[r][0] = 0.0; acc[r][1] = 0.0; }

  uint32_t par_stage[2] = {0u, 0u}, par_drain[2] = {0u, 0u};
  int drained = 0;
  auto drain = [&](int g) {
    const int buf = g & 1;
    mbar_wait(smem_u32(&s_bar[2 + buf]), par_drain[buf]);
    par_drain[buf] ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t taddr = tmem + ((uint32_t)(32 * quarter) << 16) + (uint32_t)(buf * NPAD + half * (NPAD / 2));
    uint32_t v[NREP][4];
#pragma unroll
    for (int r = 0; r < NREP; ++r) tmem_ld_16x256b(taddr + 8 * r, v[r]);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int r = 0; r < NREP; ++r) {
      acc[r][0] += (double)__uint_as_float(v[r][0]) + (double)__uint_as_float(v[r][2]);
      acc[r][1] += (double)__uint_as_float(v[r][1]) + (double)__uint_as_float(v[r][3]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  };

  // raw rows of the macro tile [t0, t0 + 256) -> s_in, asynchronously (one row per thread)
  auto prefetch_tile = [&](int64_t t0) {
    const int64_t srow = t0 + tid;
    if (srow >= hi) return;
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
      cp_async4(base + (size_t)kTcMacro * 4 + tid, static_cast<const float*>(J.weights) + r);
    }
  };
  prefetch_tile(lo);

  float xi32[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) xi32[k] = (float)J.xi_body[k];
  const uint32_t idesc32 = umma_idesc(2, NPAD), idesc16 = umma_idesc(1, NPAD);
  const int q4 = tid & 7, slice = tid >> 3;       // phase A: 4 consecutive points x the bins slice, slice + 32, ...

  int T = 0;                                      // sub-tiles issued so far
  for (int64_t tile0 = lo; tile0 < hi; tile0 += kTcMacro) {
    const int count = (int)((hi - tile0 < kTcMacro) ? hi - tile0 : kTcMacro);
    cp_async_wait_all();
    __syncthreads();          // raw rows landed; the previous macro tile's last scale step has read s_pt
    // ======================================= front: one point per thread =======================================================
    {
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = make_float4(0.f, 0.f, 0.f, 0.f);
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
          w_sel = (double)base[(size_t)kTcMacro * 4 + tid];
          point_front32(J, C, xi32, xyzt.x, xyzt.y, xyzt.z, xyzt.w, o);
        }
        const double lam = J.point_lambda ? J.point_lambda[srow] : 1.0;
        const double w_desk = w_sel * (double)o.w_time;
        CERT(GCL_CS_SUM_W_SEL) += w_sel; CERT(GCL_CS_SUM_W_SEL_SQ) += w_sel * w_sel; CERT(GCL_CS_SUM_W_DESKEW) += w_desk;
        CERT(GCL_CS_ROWS) += 1.0;
        a = make_float4(o.d[0], o.d[1], o.d[2], -cscale);          // fixed shift: the largest possible logit (|d.b| <= 1)
        b = make_float4(o.p0[0], o.p0[1], o.p0[2], (float)(w_desk * lam));
      }
      s_pt[2 * tid] = a; s_pt[2 * tid + 1] = b;
    }
    __syncthreads();
    prefetch_tile(tile0 + kTcMacro);
    const int n_sub = (count + kTcSub - 1) / kTcSub;
    for (int st = 0; st < n_sub; ++st, ++T) {
      const int stage = T & 1;
      unsigned char* sb = smem + (size_t)stage * pl.stage;
      if (T >= 2) { mbar_wait(smem_u32(&s_bar[stage]), par_stage[stage]); par_stage[stage] ^= 1u; }   // sub-tile T - 2 is through the tensor core
      if (drained < (T >> 4) && (T & 15) >= 2) drain(drained++);
      // ===================================== phase A: exponentials of 4 points x this thread's bins ===============================
      {
        const float4* pp = s_pt + 2 * (st * kTcSub + 4 * q4);
        const float4 a0 = pp[0], a1 = pp[2], a2 = pp[4], a3 = pp[6];
        const float2 dx01 = make_float2(a0.x, a1.x), dy01 = make_float2(a0.y, a1.y), dz01 = make_float2(a0.z, a1.z), sh01 = make_float2(a0.w, a1.w);
        const float2 dx23 = make_float2(a2.x, a3.x), dy23 = make_float2(a2.y, a3.y), dz23 = make_float2(a2.z, a3.z), sh23 = make_float2(a2.w, a3.w);
        float2 Z01 = make_float2(0.f, 0.f), Z23 = Z01, S01 = Z01, S23 = Z01;
        float m0 = -3.0e38f, m1 = m0, m2 = m0, m3 = m0;
        unsigned char* ehi = sb + pl.o_ehi;
        unsigned char* elo = sb + pl.o_elo;
#pragma unroll 2
        for (int b = slice; b < B; b += 32) {
          const float4 bb = s_atlas[b];
          const float2 bx = make_float2(bb.x, bb.x), by = make_float2(bb.y, bb.y), bz = make_float2(bb.z, bb.z);
          const float2 x01 = ffma2(dx01, bx, ffma2(dy01, by, ffma2(dz01, bz, sh01)));
          const float2 x23 = ffma2(dx23, bx, ffma2(dy23, by, ffma2(dz23, bz, sh23)));
          const float2 e01 = make_float2(fast_exp2(x01.x), fast_exp2(x01.y));
          const float2 e23 = make_float2(fast_exp2(x23.x), fast_exp2(x23.y));
          const uint32_t h0 = __float_as_uint(e01.x) & 0xFFFFE000u, h1 = __float_as_uint(e01.y) & 0xFFFFE000u;
          const uint32_t h2 = __float_as_uint(e23.x) & 0xFFFFE000u, h3 = __float_as_uint(e23.y) & 0xFFFFE000u;
          const float2 l01 = fadd2(e01, make_float2(-__uint_as_float(h0), -__uint_as_float(h1)));
          const float2 l23 = fadd2(e23, make_float2(-__uint_as_float(h2), -__uint_as_float(h3)));
          Z01 = fadd2(Z01, e01); Z23 = fadd2(Z23, e23);
          S01 = ffma2(e01, x01, S01); S23 = ffma2(e23, x23, S23);
          m0 = fmaxf(m0, x01.x); m1 = fmaxf(m1, x01.y); m2 = fmaxf(m2, x23.x); m3 = fmaxf(m3, x23.y);
          const uint32_t row = (uint32_t)b;
          *reinterpret_cast<uint4*>(ehi + sw128((row >> 3) * 1024u + (row & 7u) * 128u + (uint32_t)q4 * 16u)) = make_uint4(h0, h1, h2, h3);
          *reinterpret_cast<uint2*>(elo + sw64((row >> 3) * 512u + (row & 7u) * 64u + (uint32_t)q4 * 8u)) =
              make_uint2(pack_bf16x2(l01.x, l01.y), pack_bf16x2(l23.x, l23.y));
        }
        // the four slices of a warp: lanes q4 + 8 j
        float z[4] = {Z01.x, Z01.y, Z23.x, Z23.y}, s1[4] = {S01.x, S01.y, S23.x, S23.y}, mx[4] = {m0, m1, m2, m3};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          z[i] += __shfl_xor_sync(0xffffffffu, z[i], 8); z[i] += __shfl_xor_sync(0xffffffffu, z[i], 16);
          s1[i] += __shfl_xor_sync(0xffffffffu, s1[i], 8); s1[i] += __shfl_xor_sync(0xffffffffu, s1[i], 16);
          mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], 8)); mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], 16));
        }
        if (lane < 8) {
          float4* zz = reinterpret_cast<float4*>(s_z2);
          zz[(0 * 8 + warp) * 8 + q4] = make_float4(z[0], z[1], z[2], z[3]);
          zz[(1 * 8 + warp) * 8 + q4] = make_float4(s1[0], s1[1], s1[2], s1[3]);
          zz[(2 * 8 + warp) * 8 + q4] = make_float4(mx[0], mx[1], mx[2], mx[3]);
        }
      }
      __syncthreads();
      // ===================================== scale: Z_i, certificate partials, the Phi tile ==========================================
      {
        const int p = lane;                                   // point of the sub-tile; warp u writes moments u, u + 8, u + 16
        float Z = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) Z += s_z2[(0 * 8 + w) * kTcSub + p];
        const float4 a = s_pt[2 * (st * kTcSub + p)], bq = s_pt[2 * (st * kTcSub + p) + 1];
        const float invZ = 1.f / Z;
        const float wn = bq.w * invZ;
        if (warp == 0 && st * kTcSub + p < count) {
          float S1 = 0.f, xmax = -3.0e38f;
#pragma unroll
          for (int w = 0; w < 8; ++w) { S1 += s_z2[(1 * 8 + w) * kTcSub + p]; xmax = fmaxf(xmax, s_z2[(2 * 8 + w) * kTcSub + p]); }
          // H_i = -sum_b r log(r + eps) = log Z - S1/Z - B*eps + O(B eps^2 / r)   (binning.py:70-72); S1 in octaves
          CERT(GCL_CS_SUM_ENTROPY) += (double)(__logf(Z) - 0.6931471805599453f * S1 * invZ) - (double)B * C.eps_mass;
          const double mr = (double)(fast_exp2(xmax) * invZ);
          if (mr > CERT(GCL_CS_MAX_RESP)) CERT(GCL_CS_MAX_RESP) = mr;
        }
        unsigned char* a32 = sb + pl.o_a32;
        unsigned char* a16 = sb + pl.o_a16;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const int k = warp + 8 * j;                       // warp-uniform
          if (k < GCL_NM) {
            float v;
            const float wd0 = wn * a.x, wd1 = wn * a.y, wd2 = wn * a.z, wp0 = wn * bq.x, wp1 = wn * bq.y, wp2 = wn * bq.z;
            switch (k) {
              case 0: v = wn; break;
              case 1: v = wd0; break;        case 2: v = wd1; break;        case 3: v = wd2; break;
              case 4: v = wd0 * a.x; break;  case 5: v = wd0 * a.y; break;  case 6: v = wd0 * a.z; break;
              case 7: v = wd1 * a.y; break;  case 8: v = wd1 * a.z; break;  case 9: v = wd2 * a.z; break;
              case 10: v = wp0; break;       case 11: v = wp1; break;       case 12: v = wp2; break;
              case 13: v = wp0 * bq.x; break; case 14: v = wp0 * bq.y; break; case 15: v = wp0 * bq.z; break;
              case 16: v = wp1 * bq.y; break; case 17: v = wp1 * bq.z; break; default: v = wp2 * bq.z; break;
            }
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
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // st.shared -> visible to the tensor core's async proxy
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();
      if (tid == kTcIssuer) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d = tmem + (uint32_t)(((T >> 4) & 1) * NPAD);
        const uint32_t s0 = smem_u32(sb);
        uint32_t accf = (T & 15) ? 1u : 0u;                            // first instruction of a group overwrites its half
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
          mma_tf32(d, umma_desc(s0 + pl.o_a32 + kb * 32, 1024, 2), umma_desc(s0 + pl.o_ehi + kb * 32, 1024, 2), idesc32, accf);
          accf = 1u;
        }
#pragma unroll
        for (int j = 0; j < 2; ++j)
          mma_bf16(d, umma_desc(s0 + pl.o_a16 + j * 32, 512, 4), umma_desc(s0 + pl.o_elo + j * 32, 512, 4), idesc16, 1u);
        umma_commit(smem_u32(&s_bar[stage]));
        if ((T & 15) == 15) umma_commit(smem_u32(&s_bar[2 + ((T >> 4) & 1)]));
      }
    }
  }
  // --- tail: the last (partial) group, then whatever has not been drained yet ---------------------------------------------------------
  if (tid == kTcIssuer && (T & 15) != 0) umma_commit(smem_u32(&s_bar[2 + ((T >> 4) & 1)]));
  const int groups = (T + 15) >> 4;
  while (drained < groups) drain(drained++);

  if (owns) {
#pragma unroll
    for (int r = 0; r < NREP; ++r) {
      const int b0 = half * (NPAD / 2) + 8 * r + 2 * cpair;
      if (b0 < B) part[(size_t)my_moment * B + b0] = acc[r][0];
      if (b0 + 1 < B) part[(size_t)my_moment * B + b0 + 1] = acc[r][1];
    }
  }

  // --- certificate scalars of this CTA (as in scan_moments_fused_kernel) -----------------------------------------------------------------
#pragma unroll
  for (int k = 0; k < GCL_NCERT; ++k) {
    double v = CERT(k);
    for (int o = 16; o > 0; o >>= 1) {
      const double other = __shfl_xor_sync(0xffffffffu, v, o);
      v = (k == GCL_CS_MAX_RESP) ? fmax(v, other) : v + other;
    }
    if (lane == 0) s_red[warp][k] = v;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < GCL_NCERT) {
    const int src = (tid == GCL_CS_SUM_W_RAW && stride == 1) ? GCL_CS_SUM_W_SEL : tid;
    double v = s_red[0][src];
    for (int w = 1; w < kTcNT / 32; ++w) v = (src == GCL_CS_MAX_RESP) ? fmax(v, s_red[w][src]) : v + s_red[w][src];
    part[(size_t)GCL_NM * B + tid] = v;
  }
#undef CERT
  if (warp == 0) {
    constexpr uint32_t cols = (2 * NPAD <= 32) ? 32u : (2 * NPAD <= 64) ? 64u : (2 * NPAD <= 128) ? 128u : (2 * NPAD <= 256) ? 256u : 512u;
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(cols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------------------------------------------
bool tc_path_supported(int n_bins, int precision, int fixed_shift) {
  static const bool enabled = !(getenv("GCL_K1_TC") && atoi(getenv("GCL_K1_TC")) == 0);
  return enabled && precision == GCL_F32C && n_bins >= 1 && n_bins <= 256 && (fixed_shift & 1);
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
