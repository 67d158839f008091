// gcl_scan_tc.cu -- K1 in fp32c mode with the moment accumulation on the 5th-generation tensor cores (tcgen05 + tensor memory).
//
// Same contract as scan_moments_fused_kernel (gcl_scan_fused.cu): one pass over the raw points fuses the PointBudget gather, the
// constant-twist deskew + window weights, ray directions, the softmax responsibilities and the per-bin sums
//   M_b[19] = sum_i w_i r_ib [1, d, ddT(6), p, ppT(6)]                         (archive/legacy_operators/binning.py:157-173)
// and writes one fp64 partial record per CTA.  What changes is WHERE the 19 multiply-adds per (point, bin) pair run.  They are a
// dense contraction over the points,
//   M[moment][bin] += sum_i Phi[moment][i] * E[i][bin],      Phi_i = (w_i / Z_i) [1, d, ddT, p, ppT],  E[i][b] = exp(d_i.b/tau - 1/tau),
// (the "contraction depth 3" of the north star is the similarity dot, not this sum), so they go to the tensor
// cores while the CUDA cores compute the next tile's exponentials.
//
// Operands (round 2, second layout).  Everything is bf16, K-major, 128-byte rows (SWIZZLE_128B); a sub-tile is 32 points = K 64:
//   K index of a row = (quad of points q4 = 0..7, part, point of the quad): bytes [16 q4, 16 q4 + 8) hold the HEAD part of the quad's four
//       points, bytes [16 q4 + 8, 16 q4 + 16) the REMAINDER part.
//   B = E tile, n_pad bins x 64: head = bf16(e) (round to nearest), remainder = bf16(e - head): 16-17 significant bits of e.  One lane
//       owns a bin row and writes the 16 bytes of its warp's quad with ONE conflict-free 128-bit store.
//   A = Phi^T, 128 rows x 64: row 32 q + 8 c + j holds part c (0, 1, 2 = three bf16 pieces, 24 bits together) of moment 5 q + j, written
//       twice per quad (against the head and against the remainder of e); all other rows stay zero.  (Two parts only -- 16-17 bits, like
//       e -- were measured: 244 -> 238 us per 1 M x 256 launch, but kappa of the small "exact_fit" golden scan moves to 1.09e-5 of the
//       reference, over the 1e-5 bar: bins that a handful of points dominate see the element-wise error of Phi.  Three parts stay.)
//   D = 128 x n_pad fp32 in tensor memory: 4 x tcgen05.mma kind::f16 (K = 16 each) per sub-tile; rows 32 q + {j, 8 + j, 16 + j} are the
//       three parts of moment 5 q + j and are added when the accumulators are drained.
// The first layout (head of e as tf32 + remainder as bf16 in two tiles, 6 instructions per sub-tile) moved 6 bytes per (point, bin) pair
// into shared memory and 6 out again for the tensor core: ncu showed shared memory 72 % busy (LSU 49 % + tensor-core operand reads
// 22 %) at 300 us per 1 M x 256 launch -- the kernel was bound by shared-memory bandwidth, not by the tensor pipe (31 %) or MUFU (29 %).
// This layout moves 4 bytes per pair each way and needs 4 instructions.
//
// Pipeline.  Two teams of 8 producer warps and one MMA warp.  A team owns every other 256-point macro tile of the CTA's rows: its own
// raw-row staging, front end (thread per point) and named barrier -- the teams never wait for each other.  Inside a sub-tile a WARP owns a
// quad of points (lane = bins lane, lane + 32, ...): exponentials, normaliser (15-shuffle reduce-scatter), then lane = moment for the Phi
// rows; no CTA barrier, the hand-off to the MMA warp is an mbarrier.  Sub-tiles go round-robin through THREE shared-memory stages in the
// order team 0, team 1, team 0, ...; a stage is reused when the tcgen05.commit of its last use has arrived.  Every 16 sub-tiles the
// accumulation moves to the other half of tensor memory and the finished half is drained (tcgen05.ld.16x256b, twice: rows 0-15 and
// 16-31 of each lane quarter) into the CTA's fp64 record in shared memory; the MMA warp waits for those drains (mbarrier) before it
// overwrites that half again.
//
// Measured building blocks: tools/microbench/umma_probe*.cu, profiles/r2_umma_probe.md.
#include "gcl_internal.cuh"
#include "gcl_point.cuh"

namespace gcl {

constexpr int kTcTeam = 256;     // threads of one producer team (8 warps = the 8 quads of a sub-tile)
constexpr int kTcProd = 512;     // two teams
constexpr int kTcNT = 544;       // + one warp whose lane 0 issues the MMAs (a tcgen05.mma issue blocks its thread while the queue is
                                 // full -- measured 830 cycles per sub-tile -- so it must not be a thread the others wait for)
constexpr int kTcMacro = 256;    // points per front-end pass of one team (packed 20-byte rows; 128 with 40-byte float64 rows: same bytes)
constexpr int kTcSub = 32;       // points per MMA sub-tile (K = 64 with the head / remainder parts)
constexpr int kTcStages = 3;
constexpr int kTcGroupLog2 = 4;  // 2^4 = 16 sub-tiles (512 points, 64 instructions) accumulate in one half of tensor memory between drains:
                                 // the fp32 accumulation of the tensor core truncates (-1.7e-8 of the accumulator per instruction,
                                 // profiles/r2_umma_probe.md), so the bias grows with the instructions an accumulator has absorbed
constexpr int kTcGroupMask = (1 << kTcGroupLog2) - 1;
constexpr int kTcRecPad = 8;     // the fp64 record's row pitch is n_pad + 8 doubles: moment rows k, k + 1 fall in different halves of the
                                 // 128-byte bank window, so the 128-bit read-modify-writes of a drain are conflict free

__device__ __forceinline__ uint32_t sw128(uint32_t lin) { return lin ^ (((lin >> 7) & 7u) << 4); }
// shared-memory matrix descriptor (K-major): start address, leading byte offset 16, stride between 8-row groups, version 1, layout
// (2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t sbo_bytes, int layout) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)layout << 61);
}
// instruction descriptor: D fp32, A and B of format fmt (1 = bf16), both K-major, M = 128
__device__ __forceinline__ uint32_t umma_idesc(int fmt, int N) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void mma_bf16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
               :: "r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(mbar) : "memory");
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
__device__ __forceinline__ float bf16lo_f32(uint32_t pk) { return __uint_as_float(pk << 16); }
__device__ __forceinline__ float bf16hi_f32(uint32_t pk) { return __uint_as_float(pk & 0xFFFF0000u); }
template <int ID>
__device__ __forceinline__ void bar_team() { asm volatile("bar.sync %0, %1;" :: "n"(ID), "n"(kTcTeam) : "memory"); }

// the two factors of moment k out of the per-point vector {1, d0, d1, d2, p0, p1, p2, w}: Phi_k = (w / Z) v[a_k] v[b_k]; three bits per
// moment, packed in immediates (a __constant__ table indexed by the lane was re-read inside the sub-tile loop under register pressure:
// a divergent LDC, 19 replays, and the scale step took 1000 cycles)
constexpr unsigned long long kTcFacA = 0x1ad926b1a449688ull;   // [0, 1, 2, 3, 1, 1, 1, 2, 2, 3, 4, 5, 6, 4, 4, 4, 5, 5, 6]
constexpr unsigned long long kTcFacB = 0x1b5d6001b4d1000ull;   // [0, 0, 0, 0, 1, 2, 3, 2, 3, 3, 0, 0, 0, 4, 5, 6, 5, 6, 6]

// shared-memory plan of the dynamic part (host and device agree through this)
struct TcPlan {
  int n_pad;                                   // bins rounded up to what the MMA shape needs (48, 128, 256)
  uint32_t o_a, o_e, stage;                    // inside one stage
  uint32_t o_rec, rec_pitch, o_in, in_team, total;
};
__host__ __device__ inline TcPlan tc_plan(int nrep) {
  TcPlan p;
  p.n_pad = 16 * nrep;
  p.o_a = 0;                                                 // 128 rows x 128 bytes
  p.o_e = 16384;                                             // n_pad rows x 128 bytes
  p.stage = (p.o_e + (uint32_t)p.n_pad * 128u + 1023u) & ~1023u;
  p.o_rec = kTcStages * p.stage;                             // [19][n_pad + 8] doubles: this CTA's sums, carried in fp64
  p.rec_pitch = (uint32_t)p.n_pad + kTcRecPad;
  p.o_in = p.o_rec + (uint32_t)GCL_NM * p.rec_pitch * 8u;
  p.in_team = kTcMacro * 20;                                 // raw rows of one team's macro tile: 256 packed rows = 128 float64 rows
  p.total = p.o_in + 2 * p.in_team + 1024;                   // + slack for the manual 1024-byte alignment of the base
  return p;
}

#ifdef GCL_K1_TRACE
__device__ long long g_tc_trace[2][12];
#define TCT(i) do { if ((tid == 0 || tid == 256) && blockIdx.x == 1 && blockIdx.y == 0) { const long long n_ = clock64(); tc_acc[i] += n_ - tc_t; tc_t = n_; } } while (0)
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
  __shared__ float4 s_dq[2][kTcMacro / 4][3];      // [team][quad of points][dx, dy, dz]: one float4 = that component of 4 consecutive points
  __shared__ float4 s_v[2][kTcMacro][2];           // [team][point]: {1, d0, d1, d2}, {p0, p1, p2, w}
  __shared__ double s_red[kTcNT / 32][GCL_NCERT];
  __shared__ gcl_job s_job;
  __shared__ int2 s_lane[32];                      // [lane = moment k]: {4 a_k | 4 b_k << 8 | (k % 5) << 16, byte offset of row 32 (k / 5) + k % 5 in the A tile}
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) uint64_t s_bar[3 * kTcStages + 1];   // operands ready [stage], stage consumed [stage], tensor-memory half complete [2],
                                                               // tensor-memory half drained [2]
  constexpr int NPAD = 16 * NREP;
  constexpr int BAR_FULL = 0, BAR_EMPTY = kTcStages, BAR_DONE = 2 * kTcStages, BAR_DRAINED = 2 * kTcStages + 2;
  static_assert(BAR_DRAINED + 2 <= 3 * kTcStages + 1, "barrier array");
  const TcPlan pl = tc_plan(NREP);
  const int B = P.n_bins;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#ifdef GCL_K1_TRACE
  long long tc_acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, tc_t = clock64();
#endif
  unsigned char* smem = smem_raw_tc + ((1024u - (smem_u32(smem_raw_tc) & 1023u)) & 1023u);   // swizzled tiles need 1024-byte alignment

  if (tid < (int)(sizeof(gcl_job) / sizeof(int)))
    reinterpret_cast<int*>(&s_job)[tid] = reinterpret_cast<const int*>(P.jobs ? &P.jobs[blockIdx.y] : &P.job)[tid];
  // operand tiles and the fp64 record start as zeros (tile rows that are never written: unused moment rows, bins beyond n_bins, points
  // beyond a partial sub-tile)
  for (uint32_t i = tid; i < pl.o_in / 16; i += kTcNT) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  constexpr uint32_t kCols = (2 * NPAD <= 32) ? 32u : (2 * NPAD <= 64) ? 64u : (2 * NPAD <= 128) ? 128u : (2 * NPAD <= 256) ? 256u : 512u;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&s_tmem)), "r"(kCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    for (int s = 0; s < kTcStages; ++s) {
      mbar_init(smem_u32(&s_bar[BAR_FULL + s]), kTcTeam);      // every thread of the producing team arrives
      mbar_init(smem_u32(&s_bar[BAR_EMPTY + s]), 1);           // tcgen05.commit arrives
    }
    for (int h = 0; h < 2; ++h) {
      mbar_init(smem_u32(&s_bar[BAR_DONE + h]), 1);            // tcgen05.commit arrives
      mbar_init(smem_u32(&s_bar[BAR_DRAINED + h]), kTcProd);   // every producer thread arrives after its share of the drain
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 32) {
    const int k = tid < GCL_NM ? tid : 0, q = k / 5, j = k - 5 * q;
    const int fa = (int)((kTcFacA >> (3 * k)) & 7ull), fb = (int)((kTcFacB >> (3 * k)) & 7ull);
    s_lane[tid] = make_int2((4 * fa) | ((4 * fb) << 8) | (j << 16), (4 * q) * 1024 + j * 128);
  }
  const double cscale_d = 1.4426950408889634 / P.tau;                // logits in octaves
  const float cscale = (float)cscale_d;
  for (int b = tid; b < 256; b += kTcNT)
    s_atlas[b] = (b < B) ? make_float4((float)(P.bin_dirs[b * 3] * cscale_d), (float)(P.bin_dirs[b * 3 + 1] * cscale_d),
                                       (float)(P.bin_dirs[b * 3 + 2] * cscale_d), 0.f)
                         : make_float4(0.f, 0.f, 0.f, 0.f);
  // per-thread certificate partials (registers): raw weights, selected weights, their squares, deskewed weights, rows, entropy, max r
  // (entropy in fp32: a lane sums a few hundred values of order 1; rows as an integer)
  double c_wraw = 0.0, c_wsel = 0.0, c_wsel2 = 0.0, c_wdesk = 0.0;
  float ent_acc = 0.f, mr_acc = 0.f;
  int c_rows = 0, ent_rows = 0;
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
  const int macro = (P.layout == GCL_LAYOUT_REF_F64) ? kTcMacro / 2 : kTcMacro;   // points per front-end pass of a team (5 KB of raw rows)
  double* s_rec = reinterpret_cast<double*>(smem + pl.o_rec);        // [moment][rec_pitch]

  // --- the order of the sub-tiles: macro tile m belongs to team m & 1; the MMA warp takes team 0's r-th sub-tile, then team 1's r-th --------
  const int spm = macro / kTcSub;                                    // sub-tiles of a full macro tile
  const int n_full = (int)((hi - lo) / macro);
  const int rem_sub = (int)((((hi - lo) % macro) + kTcSub - 1) / kTcSub);
  const int n_sub0 = ((n_full + 1) >> 1) * spm + ((n_full & 1) ? 0 : rem_sub);
  const int n_sub1 = (n_full >> 1) * spm + ((n_full & 1) ? rem_sub : 0);
  const int n_min = n_sub0 < n_sub1 ? n_sub0 : n_sub1;
  const int n_sub_total = n_sub0 + n_sub1;
  const int n_groups = (n_sub_total + kTcGroupMask) >> kTcGroupLog2;

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
      const uint32_t idesc = umma_idesc(1, NPAD);
      for (int T = 0; T < n_sub_total; ++T) {
        const int stage = T % kTcStages, use = T / kTcStages;
        const int g = T >> kTcGroupLog2, half = g & 1;
        if ((T & kTcGroupMask) == 0 && g >= 2)                           // this half still holds group g - 2 until every warp has drained it
          mbar_wait(smem_u32(&s_bar[BAR_DRAINED + half]), (uint32_t)(((g >> 1) - 1) & 1));
        mbar_wait(smem_u32(&s_bar[BAR_FULL + stage]), (uint32_t)(use & 1));   // the producing team has written (and fenced) both operand tiles
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d = tmem + (uint32_t)(half * NPAD);
        const uint32_t s0 = smem_u32(smem + (size_t)stage * pl.stage);
        uint32_t accf = (T & kTcGroupMask) ? 1u : 0u;                    // first instruction of a group overwrites its half
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
          mma_bf16(d, umma_desc(s0 + pl.o_a + kb * 32, 1024, 2), umma_desc(s0 + pl.o_e + kb * 32, 1024, 2), idesc, accf);
          accf = 1u;
        }
        umma_commit(smem_u32(&s_bar[BAR_EMPTY + stage]));
        if ((T & kTcGroupMask) == kTcGroupMask || T == n_sub_total - 1) umma_commit(smem_u32(&s_bar[BAR_DONE + half]));
      }
    }
    __syncwarp();
  } else {
    // ================================= the producers: two teams of 8 warps ===================================================================
    const int team = warp >> 3, wq = warp & 7;       // wq = this warp's quad of the sub-tile
    const int ttid = tid - team * kTcTeam;           // thread of the team = point of the team's macro tile in the front end
    // ownership in a drain: warp = (quarter of the tensor-memory lanes, part of the bins); inside the 16x256b fragment thread t holds
    // rows t/4 and t/4 + 8 for the two bins 2(t%4), 2(t%4) + 1 of every 8-column block
    constexpr int PARTS = (NPAD % 32 == 0) ? 4 : 2;  // a 48-bin record splits in 2 x 24 columns (warps 8..15 only arrive then)
    constexpr int REPS = NPAD / (8 * PARTS);
    const int quarter = warp & 3, bpart = warp >> 2;
    const int kq = lane >> 2, cpair = lane & 3;
    const int my_moment = 5 * quarter + kq;
    const bool owns = kq < (quarter < 3 ? 5 : 4) && bpart < PARTS;

    int drained = 0;
    auto drain = [&](int g) {
      const int half = g & 1;
      mbar_wait(smem_u32(&s_bar[BAR_DONE + half]), (uint32_t)((g >> 1) & 1));
      if (bpart < PARTS) {                                   // warp-uniform
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t taddr = tmem + ((uint32_t)(32 * quarter) << 16) + (uint32_t)(half * NPAD + bpart * (NPAD / PARTS));
        double* rec = s_rec + (size_t)my_moment * pl.rec_pitch + bpart * (NPAD / PARTS) + 2 * cpair;
#pragma unroll
        for (int r0 = 0; r0 < REPS; r0 += 4) {
          uint32_t v[4][4], u[4][4];
#pragma unroll
          for (int r = 0; r < 4; ++r)
            if (r0 + r < REPS) { tmem_ld_16x256b(taddr + 8 * (r0 + r), v[r]); tmem_ld_16x256b(taddr + (16u << 16) + 8 * (r0 + r), u[r]); }
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (owns) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
              if (r0 + r < REPS) {
                double2 a = *reinterpret_cast<double2*>(rec + 8 * (r0 + r));
                // part 0 (row j) + part 1 (row 8 + j) + part 2 (row 16 + j)
                a.x += (double)((__uint_as_float(v[r][0]) + __uint_as_float(v[r][2])) + __uint_as_float(u[r][0]));
                a.y += (double)((__uint_as_float(v[r][1]) + __uint_as_float(v[r][3])) + __uint_as_float(u[r][1]));
                *reinterpret_cast<double2*>(rec + 8 * (r0 + r)) = a;
              }
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      }
      mbar_arrive(smem_u32(&s_bar[BAR_DRAINED + half]));      // the MMA warp may overwrite this half (group g + 2) once all 512 have arrived
    };

    // raw rows of this team's macro tile [t0, t0 + macro) -> its staging buffer, asynchronously (one row per thread)
    double* s_in = reinterpret_cast<double*>(smem + pl.o_in + (size_t)team * pl.in_team);
    auto prefetch_tile = [&](int64_t t0) {
      const int64_t srow = t0 + ttid;
      if (ttid >= macro || srow >= hi) return;
      const int64_t r = srow * stride;                                        // raw row (PointBudget gather)
      const int64_t rp = (J.flags & GCL_JOB_COMPACT_ROWS) ? srow : r;
      if (P.layout == GCL_LAYOUT_REF_F64) {
        const double* pts = static_cast<const double*>(J.points) + rp * 3;
        double* dst = s_in + (size_t)ttid * 5;
        cp_async8(dst + 0, pts + 0); cp_async8(dst + 1, pts + 1); cp_async8(dst + 2, pts + 2);
        cp_async8(dst + 3, static_cast<const double*>(J.timestamps) + rp);
        cp_async8(dst + 4, static_cast<const double*>(J.weights) + r);
      } else {
        float* base = reinterpret_cast<float*>(s_in);
        cp_async16(base + (size_t)ttid * 4, static_cast<const float4*>(J.points) + rp);
        cp_async4(base + (size_t)kTcMacro * 4 + ttid, static_cast<const float*>(J.weights) + r);
      }
    };
    prefetch_tile(lo + (int64_t)team * macro);

    float xi32[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) xi32[k] = (float)J.xi_body[k];
    // phase A: WARP = one quad of points (4 consecutive points of the sub-tile), LANE = the bins lane, lane + 32, ...: a warp carries its
    // quad from the exponentials through the normaliser to the scaled rows without meeting the other warps.  The lane's atlas entries
    // never change: they live in registers.  The softmax shift is the same for every point (fixed shift -1/tau): it is folded into the
    // bin's constants, and a bin beyond n_bins gets -1e30 instead, which makes its exponential exactly 0 -- so the bin loop has no
    // branches (only the stores of a 48-row tile's last step are predicated).
    constexpr int NJ = (NPAD + 31) / 32;
    float abx[NJ], aby[NJ], abz[NJ], ash[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const float4 bb = s_atlas[lane + 32 * j];
      abx[j] = bb.x; aby[j] = bb.y; abz[j] = bb.z; ash[j] = (lane + 32 * j < B) ? -cscale : -1.0e30f;
    }
    // E tile: bin row lane + 32 j, 16-byte chunk wq of the 128-byte row
    const uint32_t off_e0 = pl.o_e + sw128((uint32_t)(lane >> 3) * 1024u + (uint32_t)(lane & 7) * 128u + (uint32_t)wq * 16u);
    const float* s_v_f = reinterpret_cast<const float*>(&s_v[team][0][0]);

    int my_subs = 0;                                // sub-tiles this team has handed over so far
    for (int64_t tile0 = lo + (int64_t)team * macro; tile0 < hi; tile0 += 2 * macro) {
      const int count = (int)((hi - tile0 < macro) ? hi - tile0 : macro);
      cp_async_wait_all();
      if (team == 0) bar_team<1>(); else bar_team<2>();   // raw rows landed; every warp of the team is done with the previous macro tile's s_dq / s_v
      // ======================================= front: one point per thread =======================================================
      if (ttid < macro) {
        float d0 = 0.f, d1 = 0.f, d2 = 0.f;
        float4 pw = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ttid < count) {
          const int64_t srow = tile0 + ttid;
          double w_sel;
          PointOut32 o;
          if (P.layout == GCL_LAYOUT_REF_F64) {
            const double* in = s_in + (size_t)ttid * 5;
            w_sel = in[4];
            point_front32(J, C, xi32, (float)in[0], (float)in[1], (float)in[2], (float)(in[3] - J.scan_start_time), o);
          } else {
            const float* base = reinterpret_cast<const float*>(s_in);
            const float4 xyzt = *reinterpret_cast<const float4*>(base + (size_t)ttid * 4);
            w_sel = (double)base[(size_t)kTcMacro * 4 + ttid];
            point_front32(J, C, xi32, xyzt.x, xyzt.y, xyzt.z, xyzt.w, o);
          }
          const double lam = J.point_lambda ? J.point_lambda[srow] : 1.0;
          const double w_desk = w_sel * (double)o.w_time;
          c_wsel += w_sel; c_wsel2 += w_sel * w_sel; c_wdesk += w_desk; c_rows += 1;
          d0 = o.d[0]; d1 = o.d[1]; d2 = o.d[2];
          pw = make_float4(o.p0[0], o.p0[1], o.p0[2], (float)(w_desk * lam));
        }
        float* dq = reinterpret_cast<float*>(&s_dq[team][0][0]) + (size_t)(ttid >> 2) * 12 + (ttid & 3);
        dq[0] = d0; dq[4] = d1; dq[8] = d2;
        s_v[team][ttid][0] = make_float4(1.f, d0, d1, d2);
        s_v[team][ttid][1] = pw;
      }
      if (team == 0) bar_team<1>(); else bar_team<2>();
      prefetch_tile(tile0 + 2 * macro);
      TCT(1);   // row wait + front end
      const int n_sub = (count + kTcSub - 1) / kTcSub;
      for (int st = 0; st < n_sub; ++st, ++my_subs) {
        const int T = my_subs < n_min ? 2 * my_subs + team : 2 * n_min + (my_subs - n_min);   // place in the MMA warp's order
        const int stage = T % kTcStages, use = T / kTcStages;
        // groups up to g - 2 must be drained by every warp before the MMA warp may start group g (it waits for that, and this warp's
        // stage wait below may in turn wait for the MMA warp): those drains come first
        while (drained < (T >> kTcGroupLog2) - 1) drain(drained++);
        if (use >= 1) mbar_wait(smem_u32(&s_bar[BAR_EMPTY + stage]), (uint32_t)((use - 1) & 1));   // the stage's previous sub-tile is through the tensor core
        TCT(2);   // stage wait
        if (drained < (T >> kTcGroupLog2) && (T & kTcGroupMask) >= 2) drain(drained++);             // group g - 1: complete by now, normally
        TCT(3);   // drain
        unsigned char* const sb = smem + (size_t)stage * pl.stage;
        // ===================================== phase A: exponentials of this warp's 4 points x this lane's bins ===========================
        const int quad = st * (kTcSub / 4) + wq;
        const float4 DX = s_dq[team][quad][0], DY = s_dq[team][quad][1], DZ = s_dq[team][quad][2];
        const float2 dx01 = make_float2(DX.x, DX.y), dx23 = make_float2(DX.z, DX.w), dy01 = make_float2(DY.x, DY.y), dy23 = make_float2(DY.z, DY.w);
        const float2 dz01 = make_float2(DZ.x, DZ.y), dz23 = make_float2(DZ.z, DZ.w);
        float2 Z01 = make_float2(0.f, 0.f), Z23 = Z01, S01 = Z01, S23 = Z01;
        float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f;                        // largest exponential of each point over this lane's bins
        {
          unsigned char* et = sb + off_e0;
          float2 p01 = make_float2(0.f, 0.f), p23 = p01;                     // the previous step's exponentials (3-input maxima, two steps at a time)
#pragma unroll
          for (int j = 0; j < NJ; ++j) {
            const float2 bx = make_float2(abx[j], abx[j]), by = make_float2(aby[j], aby[j]), bz = make_float2(abz[j], abz[j]);
            const float2 sh = make_float2(ash[j], ash[j]);
            const float2 x01 = ffma2(dx01, bx, ffma2(dy01, by, ffma2(dz01, bz, sh)));
            const float2 x23 = ffma2(dx23, bx, ffma2(dy23, by, ffma2(dz23, bz, sh)));
            const float2 e01 = make_float2(fast_exp2(x01.x), fast_exp2(x01.y));
            const float2 e23 = make_float2(fast_exp2(x23.x), fast_exp2(x23.y));
            const uint32_t h01 = pack_bf16x2(e01.x, e01.y), h23 = pack_bf16x2(e23.x, e23.y);          // heads, round to nearest
            const float2 l01 = fadd2(e01, make_float2(-bf16lo_f32(h01), -bf16hi_f32(h01)));         // exact remainders
            const float2 l23 = fadd2(e23, make_float2(-bf16lo_f32(h23), -bf16hi_f32(h23)));
            Z01 = fadd2(Z01, e01); Z23 = fadd2(Z23, e23);
            S01 = ffma2(e01, x01, S01); S23 = ffma2(e23, x23, S23);        // a masked bin adds 0 * -1e30 = -0
            if ((j & 1) || j == NJ - 1) {
              m0 = fmaxf(fmaxf(m0, p01.x), e01.x); m1 = fmaxf(fmaxf(m1, p01.y), e01.y);
              m2 = fmaxf(fmaxf(m2, p23.x), e23.x); m3 = fmaxf(fmaxf(m3, p23.y), e23.y);
            } else {
              p01 = e01; p23 = e23;
            }
            if (NPAD % 32 == 0 || j < NJ - 1 || lane + 32 * j < NPAD)      // only the last step of a 48-row tile has rows to skip
              *reinterpret_cast<uint4*>(et + j * 4096) = make_uint4(h01, h23, pack_bf16x2(l01.x, l01.y), pack_bf16x2(l23.x, l23.y));
          }
        }
        TCT(4);   // phase A
        // ---- sums over the 32 lanes (= all bins): one REDUX per value on 2^23-scaled integers (a lane's partial is below 2^3 NJ <= 64
        //      in magnitude... Z partial <= NJ, |S| partial <= 0.6 NJ: the 32-lane sums stay below 2^31), maxima with the float REDUX.
        //      Every lane ends up with all values (uniform registers): no shuffles in the scale step either.
        float Zs[4], Ss[4], Ms[4];
        {
          const float zin[4] = {Z01.x, Z01.y, Z23.x, Z23.y}, sin_[4] = {S01.x, S01.y, S23.x, S23.y}, min_[4] = {m0, m1, m2, m3};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint32_t zi, zo; int32_t si, so; float mo;
            asm("cvt.rni.u32.f32 %0, %1;" : "=r"(zi) : "f"(zin[i] * 8388608.f));
            asm("cvt.rni.s32.f32 %0, %1;" : "=r"(si) : "f"(sin_[i] * 8388608.f));
            asm volatile("redux.sync.add.u32 %0, %1, 0xffffffff;" : "=r"(zo) : "r"(zi));
            asm volatile("redux.sync.add.s32 %0, %1, 0xffffffff;" : "=r"(so) : "r"(si));
            asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(mo) : "f"(min_[i]));
            Zs[i] = (float)zo * 1.1920928955078125e-07f; Ss[i] = (float)so * 1.1920928955078125e-07f; Ms[i] = mo;
          }
        }
        TCT(5);   // reduction
        // ===================================== scale: Z_i, certificate partials, the Phi rows ==========================================
        {
          float invZ[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(invZ[i]) : "f"(Zs[i]));
          {
            // H_i = -sum_b r log(r + eps) = log Z - S1/Z - B*eps + O(B eps^2 / r)   (binning.py:70-72); S1 in octaves.  Every lane computes
            // the four points of the quad (the inputs are warp-uniform), lane 0 keeps the sums
            const int left = count - (st * kTcSub + 4 * wq);            // valid points of this quad: >= 4, or fewer at the end of the tile
            float hs = 0.f, mr = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float h = 0.6931471805599453f * (__log2f(Zs[i]) - Ss[i] * invZ[i]);
              hs += (i < left) ? h : 0.f;
              mr = fmaxf(mr, (i < left) ? Ms[i] * invZ[i] : 0.f);
            }
            if (lane == 0) { ent_acc += hs; ent_rows += left < 4 ? (left > 0 ? left : 0) : 4; mr_acc = fmaxf(mr_acc, mr); }
          }
          // lane = moment k (< 19): Phi_k of the quad's four points = (w / Z) v[a_k] v[b_k], cut in three bf16 parts
          if (lane < GCL_NM) {
            const int2 lc = s_lane[lane];
            const unsigned char* vq = reinterpret_cast<const unsigned char*>(s_v_f + (size_t)(st * kTcSub + 4 * wq) * 8);
            const unsigned char* va = vq + (lc.x & 0xff);
            const unsigned char* vb = vq + ((lc.x >> 8) & 0xff);
            float phi[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float w = *reinterpret_cast<const float*>(vq + 32 * i + 28);
              phi[i] = (w > 0.f ? w * invZ[i] : 0.f) * (*reinterpret_cast<const float*>(va + 32 * i) * *reinterpret_cast<const float*>(vb + 32 * i));
            }                                                          // (rows beyond the tile's count carry w = 0)
            unsigned char* at = sb + pl.o_a + (uint32_t)lc.y + (uint32_t)((wq ^ (lc.x >> 16)) & 7) * 16u;   // rows 32 q + 8 c + j, chunk wq
            const uint32_t a01 = pack_bf16x2(phi[0], phi[1]), a23 = pack_bf16x2(phi[2], phi[3]);
            const float r0 = phi[0] - bf16lo_f32(a01), r1 = phi[1] - bf16hi_f32(a01), r2 = phi[2] - bf16lo_f32(a23), r3 = phi[3] - bf16hi_f32(a23);
            const uint32_t b01 = pack_bf16x2(r0, r1), b23 = pack_bf16x2(r2, r3);
            const uint32_t c01 = pack_bf16x2(r0 - bf16lo_f32(b01), r1 - bf16hi_f32(b01)), c23 = pack_bf16x2(r2 - bf16lo_f32(b23), r3 - bf16hi_f32(b23));
            *reinterpret_cast<uint4*>(at) = make_uint4(a01, a23, a01, a23);             // against the head and against the remainder of e
            *reinterpret_cast<uint4*>(at + 1024) = make_uint4(b01, b23, b01, b23);
            *reinterpret_cast<uint4*>(at + 2048) = make_uint4(c01, c23, c01, c23);
          }
        }
        TCT(6);   // scale
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // this thread's st.shared -> visible to the tensor core's async proxy
        mbar_arrive(smem_u32(&s_bar[BAR_FULL + stage]));                 // ... and handed to the MMA warp (the team's 256 arrivals complete the phase)
        TCT(7);   // proxy fence + arrive
      }
    }
    // --- whatever has not been drained yet (the MMA warp commits the last, possibly partial, group itself) ---------------------------------
    while (drained < n_groups) drain(drained++);
  }

  // --- the fp64 record of this CTA -> its partial record; certificate scalars (as in scan_moments_fused_kernel) ----------------------------
  double cert[GCL_NCERT] = {c_wraw, c_wsel, c_wsel2, c_wdesk, (double)ent_acc - (double)ent_rows * ((double)B * C.eps_mass), (double)mr_acc,
                            (double)c_rows, 0.0};
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
    part[i] = s_rec[(size_t)k * pl.rec_pitch + b];
  }
  if (tid < GCL_NCERT) {
    const int src = (tid == GCL_CS_SUM_W_RAW && stride == 1) ? GCL_CS_SUM_W_SEL : tid;
    double v = s_red[0][src];
    for (int w = 1; w < kTcNT / 32; ++w) v = (src == GCL_CS_MAX_RESP) ? fmax(v, s_red[w][src]) : v + s_red[w][src];
    part[(size_t)GCL_NM * B + tid] = v;
  }
#ifdef GCL_K1_TRACE
  TCT(11);  // tail: last drains, write-out, certificates
  if ((tid == 0 || tid == 256) && blockIdx.x == 1 && blockIdx.y == 0)
    for (int i = 0; i < 12; ++i) g_tc_trace[tid == 0 ? 0 : 1][i] = tc_acc[i];
#endif
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(kCols) : "memory");
}

// ------------------------------------------------------------------------------------------------------------------------------------
bool tc_path_supported(int n_bins, int precision, int fixed_shift) {
  // GCL_K1_TC: 0 = never, 2 = whenever the kernel can (any bin count up to 256), default = where it wins.  The tensor-core kernel's cost
  // per point hardly depends on the bins (the instruction stream of the tensor core, the normaliser reduction and the Phi rows are per
  // point), the CUDA-core sums scale with them.  Measured per 1 M points (tools/tc_report.py): 256 bins 244 us against 369 us on the CUDA
  // cores, 128 bins 179 / 191 us, 48 bins (200 scans x 7500 rows) 267 / 218 us -- the cross-over is just below 128 bins.
  static const int mode = getenv("GCL_K1_TC") ? atoi(getenv("GCL_K1_TC")) : 1;
  if (mode == 0 || precision != GCL_F32C || n_bins < 1 || n_bins > 256 || !(fixed_shift & 1)) return false;
  return mode == 2 || n_bins >= 128;
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
