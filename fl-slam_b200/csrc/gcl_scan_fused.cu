// gcl_scan_fused.cu -- K1: the one kernel that touches O(N) data.
//
// One pass over the raw points fuses: PointBudget gather (stride), constant-twist deskew + time-window weights,
// ray directions, softmax responsibilities over the bin atlas and the per-bin weighted moment sums
//   M_b[19] = sum_i w_i r_ib [1, d, ddT(6), p, ppT(6)]
// without materialising the (N,B) responsibilities or the (N,3,3) outer products the reference builds
// (archive/legacy_operators/binning.py:69,:171).
//
// Per CTA (256 threads, persistent over its slice of the selected rows), tile = TP points (TP = 256/TA, a power of two):
//   front     threads [0,TP): cp.async-staged raw row -> deskew, window weight, direction d_i (kept in registers),
//             {d_i, shift} to shared memory.
//   phase A   TA threads per point, each over B/TA bins: e_ib = exp(d_i . b'_b + shift) with the atlas pre-scaled
//             (fp32: log2(e)/tau and ex2.approx; fp64: 128 log2(e)/tau and a table-driven exp, below); e_ib goes to a
//             shared-memory tile [point][bin] and is computed ONCE; partial Z, sum e x (entropy) and max per thread.
//   scale     threads [0,TP): Z_i, certificate partials, row (w_i/Z_i) * phi_i to shared memory.
//   phase B   "GEMM by hand" on CUDA cores: thread = (bin lane, point slice) owns G consecutive bins x 19 accumulators in
//             registers; per point one vector LDS of its G exponentials and warp-broadcast LDS.128 of the row, then
//             19 G multiply-adds (fp32: packed fma.rn.f32x2 over bin pairs; fp64: DFMA).
//   flush     per-thread partials are combined across point slices in a FIXED order and carried in fp64 in this CTA's
//             record (no atomics => bit-reproducible).  fp32 mode flushes at least every ~512 accumulated points per
//             thread so that fp32 chains stay short ("fp32c").
// The similarity dot has contraction depth 3 (no tensor-core shape), but phase B is a dense contraction over the POINTS: in fp32c
// mode with 128..256 bins it runs on the tensor cores (gcl_scan_tc.cu); this file is the CUDA-core kernel: the fp64 parity mode, and
// fp32c outside that range.
#include "gcl_internal.cuh"
#include "gcl_point.cuh"

#ifndef GCL_A_NG4
#define GCL_A_NG4 2     // fp64, G = 4: groups of two bins in flight per thread in phase A (registers: the 76-double tile stays live)
#endif
#ifndef GCL_A_NGMMA
#define GCL_A_NGMMA 4   // fp64 DMMA variant: groups of two bins in flight per thread in phase A (the accumulators are 24 doubles there)
#endif
#ifndef GCL_BULK_ON
#define GCL_BULK_ON 1   // 0: every tile takes the per-thread cp.async path (A/B builds)
#endif

namespace gcl {

constexpr int kNT = 256;       // threads per CTA
constexpr int kRowPitch = 20;  // values per scaled row: 19 moments + 1 pad (16-byte aligned rows)
constexpr int kRowPitchMma = 24;  // ... of the DMMA variant: three 8-row fragments; 24 = 8 (mod 16) keeps the fragment loads conflict-free
constexpr int kKC = 5;         // moments per flush pass (10 was measured: no gain, and the G = 4 fp64 variant starts to spill)

template <typename T> struct Num;
template <> struct Num<float> {
  static constexpr double kScale = 1.4426950408889634;          // logits in octaves
  static constexpr double kToNat = 0.6931471805599453;          // octaves -> nats
  static constexpr int kEpMod = 8, kEpRem = 4;                  // e-tile pitch = 4 (mod 8) floats
  static constexpr int kVec = 4;                                // bins per phase-A step
};
template <> struct Num<double> {
  static constexpr double kScale = 1477.3197218702985;          // logits in 1/1024 octaves (kExpTab entries, gcl_point.cuh)
  static constexpr double kToNat = 6.769015435155716e-04;
  static constexpr int kEpMod = 4, kEpRem = 2;                  // e-tile pitch = 2 (mod 4) doubles
  static constexpr int kVec = 2;
};

__host__ __device__ inline int e_pitch(int cols, int mod, int rem) {
  int p = cols + ((rem - cols % mod) + mod) % mod;
  return p;
}

// ---- register tile of phase B ------------------------------------------------------------------------------------------
template <typename T, int G> struct Tile;
template <int G> struct Tile<float, G> {
  static_assert(G % 2 == 0, "fp32 pairs bins");
  float2 a[G / 2][19];
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int gp = 0; gp < G / 2; ++gp)
#pragma unroll
      for (int k = 0; k < 19; ++k) a[gp][k] = make_float2(0.f, 0.f);
  }
  static constexpr int VB = (G % 4 == 0) ? 4 : 2;  // bins per vector load; thread owns bins q*VB*L + VB*binlane + [0,VB)
  __device__ __forceinline__ void step(const float* erow /* + VB*binlane */, int L, const float* row) {
    float2 e[G / 2];
    if (VB == 4) {
#pragma unroll
      for (int q = 0; q < G / 4; ++q) {
        const float4 v = *reinterpret_cast<const float4*>(erow + q * 4 * L);
        e[2 * q] = make_float2(v.x, v.y); e[2 * q + 1] = make_float2(v.z, v.w);
      }
    } else {
#pragma unroll
      for (int q = 0; q < G / 2; ++q) e[q] = *reinterpret_cast<const float2*>(erow + q * 2 * L);
    }
    const float4* r4 = reinterpret_cast<const float4*>(row);
    const float4 f0 = r4[0], f1 = r4[1], f2 = r4[2], f3 = r4[3], f4 = r4[4];
    const float f[19] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w, f2.x, f2.y, f2.z, f2.w, f3.x, f3.y, f3.z, f3.w, f4.x, f4.y, f4.z};
#pragma unroll
    for (int k = 0; k < 19; ++k) {
      const float2 fk = make_float2(f[k], f[k]);
#pragma unroll
      for (int gp = 0; gp < G / 2; ++gp) a[gp][k] = ffma2(e[gp], fk, a[gp][k]);
    }
  }
  __device__ __forceinline__ float get(int g, int k) const { return (g & 1) ? a[g >> 1][k].y : a[g >> 1][k].x; }
};
template <int G> struct Tile<double, G> {
  double a[G][19];
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
      for (int k = 0; k < 19; ++k) a[g][k] = 0.0;
  }
  static constexpr int VB = (G % 2 == 0) ? 2 : 1;
  __device__ __forceinline__ void step(const double* erow /* + VB*binlane */, int L, const double* row) {
    double e[G];
    if (VB == 2) {
#pragma unroll
      for (int q = 0; q < G / 2; ++q) {
        const double2 v = *reinterpret_cast<const double2*>(erow + q * 2 * L);
        e[2 * q] = v.x; e[2 * q + 1] = v.y;
      }
    } else {
#pragma unroll
      for (int g = 0; g < G; ++g) e[g] = erow[g * L];
    }
    const double2* r2 = reinterpret_cast<const double2*>(row);
#pragma unroll
    for (int q = 0; q < 10; ++q) {
      const double2 f = r2[q];
#pragma unroll
      for (int g = 0; g < G; ++g) a[g][2 * q] = fma(e[g], f.x, a[g][2 * q]);
      if (2 * q + 1 < 19) {
#pragma unroll
        for (int g = 0; g < G; ++g) a[g][(2 * q + 1 < 19) ? 2 * q + 1 : 0] = fma(e[g], f.y, a[g][(2 * q + 1 < 19) ? 2 * q + 1 : 0]);
      }
    }
  }
  __device__ __forceinline__ double get(int g, int k) const { return a[g][k]; }
};

// ---- phase A: exponentials of one point over bins [b0, b1), stored to its e-tile row; partial Z, sum e x, max x -----------
// fp32 atlas: per group of 4 bins three float4 {bx0..3} {by0..3} {bz0..3}; bins >= n_bins hold zeros and are masked to e = 0
// NG groups of 4 bins at a time, written stage by stage (loads, similarity, exp, store/accumulate) so that the in-order
// issue sees NG independent chains and the MUFU latency of one group hides behind the others
template <int NG, bool TAIL>
__device__ __forceinline__ void phase_a_groups(const float4* __restrict__ p, int b, int B, float2 dx, float2 dy, float2 dz, float2 o2,
                                               float* __restrict__ erow, float2& Z2, float2& S2, float& m) {
  float4 ax[NG], ay[NG], az[NG];
#pragma unroll
  for (int g = 0; g < NG; ++g) { ax[g] = p[3 * g]; ay[g] = p[3 * g + 1]; az[g] = p[3 * g + 2]; }
  float2 s0[NG], s1[NG], e0[NG], e1[NG];
#pragma unroll
  for (int g = 0; g < NG; ++g) { s0[g] = ffma2(dz, make_float2(az[g].x, az[g].y), o2); s1[g] = ffma2(dz, make_float2(az[g].z, az[g].w), o2); }
#pragma unroll
  for (int g = 0; g < NG; ++g) { s0[g] = ffma2(dy, make_float2(ay[g].x, ay[g].y), s0[g]); s1[g] = ffma2(dy, make_float2(ay[g].z, ay[g].w), s1[g]); }
#pragma unroll
  for (int g = 0; g < NG; ++g) { s0[g] = ffma2(dx, make_float2(ax[g].x, ax[g].y), s0[g]); s1[g] = ffma2(dx, make_float2(ax[g].z, ax[g].w), s1[g]); }
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    e0[g] = make_float2(fast_exp2(s0[g].x), fast_exp2(s0[g].y));
    e1[g] = make_float2(fast_exp2(s1[g].x), fast_exp2(s1[g].y));
  }
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    if (TAIL) {
      const int bb = b + 4 * g;
      if (bb + 0 >= B) { e0[g].x = 0.f; s0[g].x = -3.0e38f; }
      if (bb + 1 >= B) { e0[g].y = 0.f; s0[g].y = -3.0e38f; }
      if (bb + 2 >= B) { e1[g].x = 0.f; s1[g].x = -3.0e38f; }
      if (bb + 3 >= B) { e1[g].y = 0.f; s1[g].y = -3.0e38f; }
    }
    *reinterpret_cast<float4*>(erow + b + 4 * g) = make_float4(e0[g].x, e0[g].y, e1[g].x, e1[g].y);
    Z2 = fadd2(Z2, fadd2(e0[g], e1[g]));
    S2 = ffma2(e0[g], s0[g], S2); S2 = ffma2(e1[g], s1[g], S2);   // masked bins: 0 * -3e38 = -0
    m = fmaxf(fmaxf(m, fmaxf(s0[g].x, s0[g].y)), fmaxf(s1[g].x, s1[g].y));
  }
}
template <int NGB, bool LOWGUARD /* fp64 only: ex2.approx underflows to zero by itself */>
__device__ __forceinline__ void phase_a_bins(const float* __restrict__ atlas, const double*, int b0, int b1, int B, float d0,
                                             float d1, float d2, float off, float* __restrict__ erow, float& Z, float& S1, float& xmax) {
  const float2 dx = make_float2(d0, d0), dy = make_float2(d1, d1), dz = make_float2(d2, d2), o2 = make_float2(off, off);
  float2 Z2 = make_float2(0.f, 0.f), S2 = make_float2(0.f, 0.f);
  float m = -3.0e38f;
  const float4* p = reinterpret_cast<const float4*>(atlas);
  const int bmain = (b1 <= B) ? b1 : ((B & ~3) > b0 ? (B & ~3) : b0);  // groups below bmain are entirely real bins
  int b = b0;
  for (; b + 4 * NGB <= bmain; b += 4 * NGB) phase_a_groups<NGB, false>(p + 3 * (b >> 2), b, B, dx, dy, dz, o2, erow, Z2, S2, m);
  for (; b < bmain; b += 4) phase_a_groups<1, false>(p + 3 * (b >> 2), b, B, dx, dy, dz, o2, erow, Z2, S2, m);
  for (; b < b1; b += 4) phase_a_groups<1, true>(p + 3 * (b >> 2), b, B, dx, dy, dz, o2, erow, Z2, S2, m);
  Z = Z2.x + Z2.y; S1 = S2.x + S2.y; xmax = m;
}
// fp64 atlas: per group of 2 bins three double2 {bx0,bx1} {by0,by1} {bz0,bz1}.  NG groups (2 NG bins) at a time, written
// stage by stage: the exp is a chain of ~12 dependent FP64 operations, and with two warps per scheduler the FP64 pipe only
// fills when each warp carries 2 NG independent chains (measured: one group in flight ran phase A at 2.5x its FP64 bound).
// Logits come in units of 1/1024 octave: 2^(x/1024) = 2^k T[j] exp(r ln2/1024), n = rint(x) = 1024 k + j, |r| <= 1/2, so a
// degree-3 polynomial is enough (truncation 5.5e-16 relative) -- 12 FP64-pipe operations per pair with the normaliser and the entropy
// sum (14 with the 128-entry table and degree 4).  The row maximum is taken over the EXPONENTIALS as 64-bit integers (they are
// positive doubles: same order), on the integer pipe; a double-precision max costs the FP64 pipe a DSETP per pair and the
// scheduler four more issue slots of selects and moves.
template <int NG, bool TAIL, bool LOWGUARD>
__device__ __forceinline__ void phase_a_groups(const double2* __restrict__ p, const double* __restrict__ tab, int b, int B, double d0,
                                               double d1, double d2, double off, double* __restrict__ erow, double& z, double& s1,
                                               unsigned long long& m) {
  constexpr int N = 2 * NG;
  const double magic = 6755399441055744.0;  // 2^52 + 2^51
  double x[N], tm[N], r[N], q[N], e[N];
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    const double2 ax = p[3 * g], ay = p[3 * g + 1], az = p[3 * g + 2];
    x[2 * g] = fma(d0, ax.x, fma(d1, ay.x, fma(d2, az.x, off)));
    x[2 * g + 1] = fma(d0, ax.y, fma(d1, ay.y, fma(d2, az.y, off)));
  }
  if (LOWGUARD) {   // row-maximum shift (tiny tau): logits reach far below 2^-1022, where the exponent arithmetic below would wrap
#pragma unroll
    for (int k = 0; k < N; ++k) x[k] = fmax(x[k], -1.0e6);   // 2^-976 against a row maximum of 1
  }
#pragma unroll
  for (int k = 0; k < N; ++k) tm[k] = x[k] + magic;
#pragma unroll
  for (int k = 0; k < N; ++k) r[k] = x[k] - (tm[k] - magic);
#pragma unroll
  for (int k = 0; k < N; ++k) q[k] = fma(5.169222938345892e-11, r[k], 2.2909784980688163e-07);
#pragma unroll
  for (int k = 0; k < N; ++k) q[k] = fma(q[k], r[k], 6.769015435155716e-04);
#pragma unroll
  for (int k = 0; k < N; ++k) q[k] = fma(q[k], r[k], 1.0);
#pragma unroll
  for (int k = 0; k < N; ++k) {
    const int n = __double2loint(tm[k]);
    const double y = tab[n & (kExpTab - 1)] * q[k];
    e[k] = __hiloint2double(__double2hiint(y) + ((n >> kExpTabLog2) << 20), __double2loint(y));
  }
  if (TAIL) {
#pragma unroll
    for (int k = 0; k < N; ++k)
      if (b + k >= B) { e[k] = 0.0; x[k] = -1.0e300; }
  }
#pragma unroll
  for (int g = 0; g < NG; ++g) *reinterpret_cast<double2*>(erow + b + 2 * g) = make_double2(e[2 * g], e[2 * g + 1]);
#pragma unroll
  for (int k = 0; k < N; ++k) {
    z += e[k]; s1 = fma(e[k], x[k], s1);
    const unsigned long long eb = (unsigned long long)__double_as_longlong(e[k]);
    m = eb > m ? eb : m;
  }
}
template <int NGB, bool LOWGUARD>
__device__ __forceinline__ void phase_a_bins(const double* __restrict__ atlas, const double* __restrict__ tab, int b0, int b1, int B,
                                             double d0, double d1, double d2, double off, double* __restrict__ erow, double& Z,
                                             double& S1, double& xmax) {
  double z = 0.0, s1 = 0.0;
  unsigned long long m = 0ull;   // bits of the largest exponential (xmax carries max_b e, not max_b x, in this precision)
  const double2* p = reinterpret_cast<const double2*>(atlas);
  const int bmain = (b1 <= B) ? b1 : ((B & ~1) > b0 ? (B & ~1) : b0);
  int b = b0;
  for (; b + 2 * NGB <= bmain; b += 2 * NGB) phase_a_groups<NGB, false, LOWGUARD>(p + 3 * (b >> 1), tab, b, B, d0, d1, d2, off, erow, z, s1, m);
  for (; b < bmain; b += 2) phase_a_groups<1, false, LOWGUARD>(p + 3 * (b >> 1), tab, b, B, d0, d1, d2, off, erow, z, s1, m);
  for (; b < b1; b += 2) phase_a_groups<1, true, LOWGUARD>(p + 3 * (b >> 1), tab, b, B, d0, d1, d2, off, erow, z, s1, m);
  Z = z; S1 = s1; xmax = __longlong_as_double((long long)m);
}

// row maximum of the scaled similarity over ALL bins (only when the fixed shift could underflow: tiny tau)
__device__ __forceinline__ float sim_rowmax(const float* atlas, int B, float d0, float d1, float d2) {
  float m = -3.0e38f;
  for (int b = 0; b < B; ++b) {
    const float* g = atlas + 12 * (b >> 2) + (b & 3);
    m = fmaxf(m, fmaf(d0, g[0], fmaf(d1, g[4], d2 * g[8])));
  }
  return m;
}
__device__ __forceinline__ double sim_rowmax(const double* atlas, int B, double d0, double d1, double d2) {
  double m = -1.0e300;
  for (int b = 0; b < B; ++b) {
    const double* g = atlas + 6 * (b >> 1) + (b & 1);
    m = fmax(m, fma(d0, g[0], fma(d1, g[2], d2 * g[4])));
  }
  return m;
}

// ---- shared-memory plan (host and device agree through these) ----------------------------------------------------------------
struct FusedPlan {
  int G, L, TP, TA, BA, EP;   // bins/thread, bin lanes, points/tile, phase-A threads/point, bins per phase-A thread, e pitch
  size_t o_tab, o_sd, o_zp, o_rows, o_pt, o_etile, o_in, total;
};
// mma = 1: the fp64 variant whose phase B runs on mma.sync.m8n8k4.f64 (rows of 24 values, e-tile pitch = 4 (mod 16) doubles: both
// operand fragments load without bank conflicts)
__host__ __device__ inline FusedPlan fused_plan(int B, int tsz, int G, int L, int TP, int mma = 0, int nt = 256) {
  FusedPlan p;
  p.G = G; p.L = L; p.TP = TP; p.TA = nt / TP;
  const int vec = tsz == 4 ? 4 : 2;
  p.BA = (((B + p.TA - 1) / p.TA) + vec - 1) / vec * vec;
  const int cols = (G * L > p.TA * p.BA) ? G * L : p.TA * p.BA;
  p.EP = tsz == 4 ? e_pitch(cols, 8, 4) : (mma == 2 ? e_pitch(cols, 16, 8) : (mma ? e_pitch(cols, 16, 4) : e_pitch(cols, 4, 2)));
  size_t off = mma == 2 ? (size_t)(G * L) * 4 * tsz                  // DMMA phase A: [bin][bx, by, bz, 1]
                        : (size_t)(p.TA * p.BA) * 3 * tsz;           // phase-A atlas, 3 components per bin
  off = (off + 15) & ~(size_t)15;
  p.o_tab = off;  off += (tsz == 8 ? kExpTab : 2) * 8;
  p.o_sd = off;   off += (size_t)TP * 4 * tsz;
  p.o_zp = off;   off += (size_t)p.TA * TP * 3 * tsz;
  off = (off + 15) & ~(size_t)15;
  p.o_rows = off; off += (size_t)TP * (mma ? kRowPitchMma : kRowPitch) * tsz;
  p.o_pt = off;   off += (size_t)(2 * TP <= nt ? 2 : 1) * 7 * TP * tsz;   // front end -> scale step, one buffer per thread group
  off = (off + 15) & ~(size_t)15;
  p.o_in = off;   off += (size_t)2 * TP * 5 * 8;                            // two staging buffers of raw rows
  p.o_etile = off;
  size_t et = (size_t)TP * p.EP * tsz;
  const size_t stage = mma ? 0 : (size_t)(nt / L) * kKC * (size_t)(G * L) * tsz;
  if (stage > et) et = stage;
  off += et;
  p.total = off + 16;
  return p;
}

#ifdef GCL_K1_TRACE
__device__ long long g_k1_trace[16];
#define K1T(i) do { if (threadIdx.x == 0 && blockIdx.x == 1 && (blockIdx.y == 7 || gridDim.y == 1)) { const long long n_ = clock64(); k1_acc[i] += n_ - k1_t; k1_t = n_; } } while (0)
extern "C" int gcl_debug_k1_trace(long long* out) { return (int)cudaMemcpyFromSymbol(out, g_k1_trace, sizeof(long long) * 16); }
#else
#define K1T(i) do { } while (0)
#endif

__device__ __forceinline__ void dmma884z(double& c0, double& c1, double a, double b) {   // C = 0
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};" : "=d"(c0), "=d"(c1) : "d"(a), "d"(b), "d"(0.0), "d"(0.0));
}
// Phase A of the MMA = 2 variant: the logits of 8 points x 8 bins per DMMA, x = [d0, d1, d2, shift] . [bx', by', bz', 1] (K = 4: one
// instruction where the FMA chain takes six per 64 logits, and phase A is bound by dispatch slots); the lane holds the point lane / 4 and
// bins 2 (lane % 4) + {0, 1} of each 8-bin fragment.  NG fragments in flight, then the same exponential / normaliser / entropy / maximum
// steps as phase_a_groups.  bfrag: this lane's atlas word of the first fragment (32 doubles per fragment); erow: its e-tile words.
template <int NG, bool TAIL>
__device__ __forceinline__ void phase_a_mma_groups(const double* __restrict__ bfrag, const double* __restrict__ tab, double afrag, int b,
                                                   int B, bool live, double* __restrict__ erow, double& z, double& s1, unsigned long long& m) {
  constexpr int N = 2 * NG;
  const double magic = 6755399441055744.0;  // 2^52 + 2^51
  double x[N], tm[N], r[N], q[N], e[N];
#pragma unroll
  for (int g = 0; g < NG; ++g) dmma884z(x[2 * g], x[2 * g + 1], afrag, bfrag[32 * g]);
#pragma unroll
  for (int k = 0; k < N; ++k) tm[k] = x[k] + magic;
#pragma unroll
  for (int k = 0; k < N; ++k) r[k] = x[k] - (tm[k] - magic);
#pragma unroll
  for (int k = 0; k < N; ++k) q[k] = fma(5.169222938345892e-11, r[k], 2.2909784980688163e-07);
#pragma unroll
  for (int k = 0; k < N; ++k) q[k] = fma(q[k], r[k], 6.769015435155716e-04);
#pragma unroll
  for (int k = 0; k < N; ++k) q[k] = fma(q[k], r[k], 1.0);
#pragma unroll
  for (int k = 0; k < N; ++k) {
    const int n = __double2loint(tm[k]);
    const double y = tab[n & (kExpTab - 1)] * q[k];
    // 2^(n >> 10) into the exponent field as shift + multiply-add (written as (n >> 10) << 20 it becomes shift + mask + add: one
    // dispatch slot more per pair, and this variant has the registers for it)
    int kq;
    asm("shr.s32 %0, %1, %2;" : "=r"(kq) : "r"(n), "n"(kExpTabLog2));
    e[k] = __hiloint2double(__double2hiint(y) + kq * 1048576, __double2loint(y));
  }
  if (TAIL) {
#pragma unroll
    for (int k = 0; k < N; ++k)
      if (b + 8 * (k >> 1) + (k & 1) >= B) { e[k] = 0.0; x[k] = -1.0e300; }
  }
  if (!live) {   // a row that pads the CTA's last tile: zeros for phase B (its own sums are never read)
#pragma unroll
    for (int k = 0; k < N; ++k) e[k] = 0.0;
  }
#pragma unroll
  for (int g = 0; g < NG; ++g) *reinterpret_cast<double2*>(erow + 8 * g) = make_double2(e[2 * g], e[2 * g + 1]);
#pragma unroll
  for (int k = 0; k < N; ++k) {
    z += e[k]; s1 = fma(e[k], x[k], s1);
    const unsigned long long eb = (unsigned long long)__double_as_longlong(e[k]);
    m = eb > m ? eb : m;
  }
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// MMA = 1 (fp64, G = 4 on 64 bin lanes, i.e. 129..256 bins): phase B as M[moment][bin] += Phi^T[moment][point] E[point][bin] on
// mma.sync.m8n8k4.f64 (SASS DMMA).  Same FP64 rate as DFMA (tools/microbench/dmma_rate.cu: 63 FMA/clk/SM), and 19 moments pad to 24
// rows -- but an eighth of the issue slots, and the accumulators are 24 doubles per thread instead of 76 (a warp owns 32 bins x all
// moments over ALL points of the tile: no point slices, so the flush is a plain store), which leaves phase A the registers for eight
// chains in flight instead of four.
__device__ unsigned int g_sm_arrivals[256];   // CTAs that have started on each SM (two-CTAs-per-SM variant: parity = first or second)

template <typename T, int G, int MINB, int MMA = 0, int NT_ = 256>
__global__ void __launch_bounds__(NT_, MINB) scan_moments_fused_kernel(const __grid_constant__ ScanParams P) {
  static_assert(!MMA || (sizeof(T) == 8 && G == 4), "the DMMA variant is fp64, G = 4");
  static_assert(NT_ == 256 || (MMA == 1 && NT_ == 512) || (MMA == 2 && NT_ == 128),
                "512 threads: the DMMA phase-B variant only (its accumulators leave room for 16 warps); 128 threads: the all-DMMA variant, two CTAs per SM");
  static_assert(sizeof(gcl_job) / sizeof(int) <= 128, "the job record is copied by the first threads of the CTA");
  constexpr int kNT = NT_;                   // threads per CTA (shadows the file-scope default)
  constexpr int NTW = 256 / (NT_ / 32) / 8;  // DMMA variant: 8-bin fragments per warp (a warp owns 256 / warps bins)
  constexpr int RP = MMA ? kRowPitchMma : kRowPitch;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int B = P.n_bins, L = P.lanes;
  __shared__ double s_red[kNT / 32][GCL_NCERT];
  __shared__ double s_cert[kNT * GCL_NCERT];

  const int tid = threadIdx.x;
#ifdef GCL_K1_TRACE
  long long k1_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, k1_t = clock64();
#endif
  // the job record is read all through the kernel (array pointers at every prefetch, times and twist at every front end): keep
  // it in shared memory instead of going back to global memory for every field
  __shared__ gcl_job s_job;
  __shared__ __align__(8) uint64_t s_mbar[2];       // completion of the bulk copies of a tile's raw rows, one per staging buffer
  __shared__ int s_bulk_on;                         // whether this CTA's full tiles go by bulk copy (kept in shared memory: the <double, 4>
                                                    // variant has no register to spare across phase B)
  if (tid < (int)(sizeof(gcl_job) / sizeof(int)))
    reinterpret_cast<int*>(&s_job)[tid] = reinterpret_cast<const int*>(P.jobs ? &P.jobs[blockIdx.y] : &P.job)[tid];
  if (tid == 0) {
    mbar_init(smem_u32(&s_mbar[0]), 1);
    mbar_init(smem_u32(&s_mbar[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const gcl_job& J = s_job;
  const ScanConst& C = P.c;

  // --- this CTA's slice of the selected rows -----------------------------------------------------------------
  const int64_t n_raw = J.n_raw;
  const int64_t cap = J.n_points_cap > 0 ? J.n_points_cap : 1;
  const int64_t stride0 = (n_raw + cap - 1) / cap;                 // ceil(n/cap), point_budget.py:160
  const int64_t stride = stride0 > 1 ? stride0 : 1;
  const int64_t n_sel_total = (n_raw + stride - 1) / stride;       // len(arange(0, n, stride))
  const int64_t sel_begin = J.sel_begin;
  const int64_t sel_end = (J.sel_end < 0 || J.sel_end > n_sel_total) ? n_sel_total : J.sel_end;
  const int64_t rows = sel_end > sel_begin ? sel_end - sel_begin : 0;
  // rows per CTA; even for contiguous float64 rows, so that every CTA of a job that starts at an even row starts at one too (16-byte
  // aligned bulk copies of 24-byte and 8-byte rows)
  const int64_t rpc0 = (rows + gridDim.x - 1) / gridDim.x;
  const int64_t rpc = (P.layout == GCL_LAYOUT_REF_F64 && stride == 1) ? ((rpc0 + 1) & ~(int64_t)1) : rpc0;
  const int64_t lo0 = sel_begin + (int64_t)blockIdx.x * rpc;
  const int64_t lo = lo0 < sel_end ? lo0 : sel_end;
  const int64_t hi = (lo + rpc < sel_end) ? lo + rpc : sel_end;
  double* part = P.partials + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * P.rec_doubles;
  // points per tile: the launch's maximum (shared-memory bound), fewer when this CTA has few rows (then more threads
  // share a point in phase A)
  int TP = P.tile_points;
  while (MMA != 2 && TP > 16 && (TP >> 1) >= rpc) TP >>= 1;   // (MMA = 2: eight warps x eight points, always)
  const FusedPlan pl = fused_plan(B, (int)sizeof(T), G, L, TP, MMA, kNT);
  const int TA = pl.TA, BA = pl.BA, EP = pl.EP, PS = kNT / L, Bp = G * L;
  T* s_atlas = reinterpret_cast<T*>(smem_raw);
  double* s_tab = reinterpret_cast<double*>(smem_raw + pl.o_tab);
  T* s_d = reinterpret_cast<T*>(smem_raw + pl.o_sd);
  T* s_zp = reinterpret_cast<T*>(smem_raw + pl.o_zp);
  T* s_rows = reinterpret_cast<T*>(smem_raw + pl.o_rows);
  T* s_pt = reinterpret_cast<T*>(smem_raw + pl.o_pt);
  T* s_etile = reinterpret_cast<T*>(smem_raw + pl.o_etile);
  T* s_stage = s_etile;  // flush staging aliases the e-tile (a flush happens between tiles)
  double* s_in = reinterpret_cast<double*>(smem_raw + pl.o_in);

  const double cscale_d = Num<T>::kScale / P.tau;
  const T cscale = (T)cscale_d;

  // --- atlas (pre-scaled) and exp table to shared (the CTA's record is written, not accumulated, by its first flush) ----------
  bool flushed = false;
  {
    const int nb = TA * BA;  // padded bin count of phase A
    constexpr int V = Num<T>::kVec;
    if (MMA == 2) {
      for (int i = tid; i < Bp * 4; i += kNT) {   // [bin][bx', by', bz', 1]: a DMMA B fragment is 32 consecutive words
        const int b = i >> 2, c = i & 3;
        s_atlas[i] = (b < B) ? (c < 3 ? (T)(P.bin_dirs[b * 3 + c] * cscale_d) : (T)1) : (T)0;
      }
    } else
    for (int i = tid; i < nb * 3; i += kNT) {   // group-major: [group][component][lane in group]
      const int g = i / (3 * V), rem = i - g * 3 * V, c = rem / V, b = g * V + (rem - c * V);
      s_atlas[i] = (b < B) ? (T)(P.bin_dirs[b * 3 + c] * cscale_d) : (T)0;
    }
    if (sizeof(T) == 8) {
      for (int i = tid; i < kExpTab; i += kNT) s_tab[i] = exp2((double)i * (1.0 / kExpTab));
    }
  }
  const int binlane = tid & (L - 1);
  const int slice = tid / L;
  const int pi = tid & (TP - 1);   // phase-A point of this thread
  const int pa = tid / TP;         // phase-A sub-thread

  constexpr int VB = Tile<T, G>::VB;
  Tile<T, G> acc;
  acc.zero();
  double cm[3][NTW][2];   // DMMA variant: [moment fragment][bin fragment of this warp's 8 NTW bins][2]
#pragma unroll
  for (int i = 0; i < 6 * NTW; ++i) (&cm[0][0][0])[i] = 0.0;

  const TimeConst tconst = time_const(J);
  // per-thread certificate partials live in shared memory (they are touched once per tile; registers are for the tile)
  // scalar k of thread t lives at s_cert[k * kNT + t]: consecutive lanes, consecutive words (thread-major rows of 8 doubles put
  // the 32 lanes of a warp on two banks: every one of the ~10 accesses per point was a 16-way conflict)
  double* my_cert = s_cert + tid;
#define CERT(k) my_cert[(k) * kNT]
#pragma unroll
  for (int k = 0; k < GCL_NCERT; ++k) CERT(k) = 0.0;
  double c_wraw = 0;

  // --- sum of ALL raw weights of this CTA's raw range (mass preservation, point_budget.py:69) ------------------
  if (stride > 1 && hi > lo) {
    const int64_t rlo = lo * stride;
    const int64_t rhi = (hi >= n_sel_total) ? n_raw : hi * stride;
    // eight loads in flight per thread (this is a latency-bound prologue: ~24 loads per thread on the stream workload, and a
    // load -> add pair per iteration would pay the DRAM latency every time); the partial sums are combined in a fixed order
    double p0 = 0, p1 = 0, p2 = 0, p3 = 0;
    int64_t r = rlo + tid;
    if (P.layout == GCL_LAYOUT_REF_F64) {
      const double* w = static_cast<const double*>(J.weights);
      for (; r + 7 * kNT < rhi; r += 8 * kNT) {
        const double a0 = w[r], a1 = w[r + kNT], a2 = w[r + 2 * kNT], a3 = w[r + 3 * kNT];
        const double a4 = w[r + 4 * kNT], a5 = w[r + 5 * kNT], a6 = w[r + 6 * kNT], a7 = w[r + 7 * kNT];
        p0 += a0; p1 += a1; p2 += a2; p3 += a3; p0 += a4; p1 += a5; p2 += a6; p3 += a7;
      }
      for (; r + 3 * kNT < rhi; r += 4 * kNT) { p0 += w[r]; p1 += w[r + kNT]; p2 += w[r + 2 * kNT]; p3 += w[r + 3 * kNT]; }
      for (; r < rhi; r += kNT) p0 += w[r];
    } else {
      const float* w = static_cast<const float*>(J.weights);
      for (; r + 7 * kNT < rhi; r += 8 * kNT) {
        const float a0 = w[r], a1 = w[r + kNT], a2 = w[r + 2 * kNT], a3 = w[r + 3 * kNT];
        const float a4 = w[r + 4 * kNT], a5 = w[r + 5 * kNT], a6 = w[r + 6 * kNT], a7 = w[r + 7 * kNT];
        p0 += (double)a0; p1 += (double)a1; p2 += (double)a2; p3 += (double)a3; p0 += (double)a4; p1 += (double)a5; p2 += (double)a6; p3 += (double)a7;
      }
      for (; r + 3 * kNT < rhi; r += 4 * kNT) { p0 += (double)w[r]; p1 += (double)w[r + kNT]; p2 += (double)w[r + 2 * kNT]; p3 += (double)w[r + 3 * kNT]; }
      for (; r < rhi; r += kNT) p0 += (double)w[r];
    }
    c_wraw = (p0 + p1) + (p2 + p3);
    CERT(GCL_CS_SUM_W_RAW) = c_wraw;
  }

  // fixed-order cross-slice reduction, kKC moments at a time, carried in fp64 in this CTA's record
  auto flush = [&]() {
    if (MMA) {   // thread (warp, lane) holds rows 8 mt + lane / 4, bins 32 warp + 8 nt + 2 (lane % 4) + {0, 1}: whole sums, stored as they are
      const int w = tid >> 5, lane = tid & 31;
#pragma unroll
      for (int mt = 0; mt < 3; ++mt)
#pragma unroll
        for (int nt = 0; nt < NTW; ++nt)
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int m = 8 * mt + (lane >> 2), b = 8 * NTW * w + 8 * nt + 2 * (lane & 3) + i;
            if (m < GCL_NM && b < B) {
              if (flushed) part[(size_t)m * B + b] += cm[mt][nt][i]; else part[(size_t)m * B + b] = cm[mt][nt][i];
            }
            cm[mt][nt][i] = 0.0;
          }
      flushed = true;
      return;
    }
#pragma unroll
    for (int k0 = 0; k0 < GCL_NM; k0 += kKC) {
#pragma unroll
      for (int kc = 0; kc < kKC; ++kc) {
        if (k0 + kc < GCL_NM) {
#pragma unroll
          for (int g = 0; g < G; ++g)
            s_stage[((size_t)slice * kKC + kc) * Bp + (g / VB) * VB * L + VB * binlane + (g % VB)] = acc.get(g, (k0 + kc < GCL_NM) ? k0 + kc : 0);
        }
      }
      __syncthreads();
      for (int e = tid; e < kKC * Bp; e += kNT) {
        const int kc = e / Bp, b = e - kc * Bp, k = k0 + kc;
        if (k < GCL_NM && b < B) {
          double s = 0.0;
          for (int sl = 0; sl < PS; ++sl) s += (double)s_stage[((size_t)sl * kKC + kc) * Bp + b];
          if (flushed) part[(size_t)k * B + b] += s; else part[(size_t)k * B + b] = s;   // no read-modify-write the first time
        }
      }
      __syncthreads();
    }
    flushed = true;
    acc.zero();
  };

  // ---- tile schedule ---------------------------------------------------------------------------------------------------------
  // Per tile only phase A and phase B use all 256 threads; the per-point steps (front end before phase A, scale between the phases)
  // are one thread per point, i.e. TP/32 warps, and each is a long dependent fp64 chain.  Tiles therefore alternate between two
  // thread groups (group of tile i = i & 1, threads [g TP, g TP + TP), when TP <= 128): while group g runs the scale step of tile i,
  // the other group runs the front end of tile i + 1 in the same barrier interval.  What the scale step needs from the front end
  // (deskewed point, direction, weight) waits in shared memory (s_pt, one buffer per group), not in registers.  With TP = 256 there
  // is one group: scale(i) then front(i + 1) back to back, one barrier fewer per tile than a front end of its own.
  //
  // Raw rows are staged two tiles ahead into two buffers.  Contiguous float64 rows (no PointBudget stride: the dense clouds) go as three
  // bulk copies (TMA engine) issued by ONE thread -- points, timestamps, weights of a full tile from the CTA's even first row (16-byte
  // alignment), arrays of rows in shared memory, completion counted by the buffer's mbarrier.  Issuing them costs that thread's warp
  // ~400 cycles, so the thread is lane 0 of the LAST warp and the moment is the start of the scale interval, when that warp idles for
  // TP <= 64.  Everything else (strided gather, packed rows, unaligned arrays or an odd first row, a CTA's last partial tile) takes one
  // row per thread with cp.async, issued by the group that will consume it.
  const int two_groups = (2 * TP <= kNT) ? 1 : 0;
  if (tid == 0) {
    s_bulk_on = (GCL_BULK_ON && P.layout == GCL_LAYOUT_REF_F64 && stride == 1 && !(lo & 1) &&
                 ((reinterpret_cast<uintptr_t>(J.points) | reinterpret_cast<uintptr_t>(J.timestamps) | reinterpret_cast<uintptr_t>(J.weights)) & 15) == 0) ? 1 : 0;
  }
  __syncthreads();
  const size_t in_stride = (size_t)TP * 5;            // doubles per staging buffer
  // full tiles only: a CTA's last, partial tile takes the per-thread path
  auto tile_is_bulk = [&](int64_t t0) -> bool { return t0 + TP <= hi && s_bulk_on; };
  auto prefetch_bulk = [&](int64_t t0, int ti) {      // one thread
    const uint32_t mb = smem_u32(&s_mbar[ti & 1]);
    double* in = s_in + (size_t)(ti & 1) * in_stride;
    mbar_arrive_expect_tx(mb, (uint32_t)TP * 40u);
    bulk_copy_g2s(in, static_cast<const double*>(J.points) + t0 * 3, (uint32_t)TP * 24u, mb);
    bulk_copy_g2s(in + (size_t)TP * 3, static_cast<const double*>(J.timestamps) + t0, (uint32_t)TP * 8u, mb);
    bulk_copy_g2s(in + (size_t)TP * 4, static_cast<const double*>(J.weights) + t0, (uint32_t)TP * 8u, mb);
  };
  // the per-thread path: called by the threads of tile ti's group with their index li in the group; ALWAYS commits a (possibly empty)
  // cp.async group so that the wait of the front end can count groups
  auto prefetch_rows = [&](int64_t t0, int ti, int li) {
    const int64_t srow = t0 + li;
    if (t0 < hi && !tile_is_bulk(t0) && srow < hi) {
      const int64_t r = srow * stride;                                        // raw row (PointBudget gather)
      const int64_t rp = (J.flags & GCL_JOB_COMPACT_ROWS) ? srow : r;         // row in the points / timestamps arrays
      double* in = s_in + (size_t)(ti & 1) * in_stride;
      if (P.layout == GCL_LAYOUT_REF_F64) {
        const double* pts = static_cast<const double*>(J.points) + rp * 3;
        double* dst = in + (size_t)li * 5;
        cp_async8(dst + 0, pts + 0); cp_async8(dst + 1, pts + 1); cp_async8(dst + 2, pts + 2);
        cp_async8(dst + 3, static_cast<const double*>(J.timestamps) + rp);
        cp_async8(dst + 4, static_cast<const double*>(J.weights) + r);
      } else {
        float* base = reinterpret_cast<float*>(in);
        cp_async16(base + (size_t)li * 4, static_cast<const float4*>(J.points) + rp);
        cp_async4(base + (size_t)TP * 4 + li, static_cast<const float*>(J.weights) + r);
      }
    }
    cp_async_commit();
  };
  // ======================================= front =============================================================
  // thread li of tile ti's group: raw row -> deskew, window weight, direction; {d, shift} for phase A, point state for the scale step
  auto front = [&](int64_t tile0, int ti, int li, bool newer_group_in_flight) {
    const int count = (int)((hi - tile0 < TP) ? hi - tile0 : TP);
    const bool from_bulk = tile_is_bulk(tile0);
    if (from_bulk) mbar_wait(smem_u32(&s_mbar[ti & 1]), (uint32_t)((ti >> 1) & 1));
    else if (newer_group_in_flight) cp_async_wait_group<1>();
    else cp_async_wait_group<0>();
    if (li >= count) return;
    const double* in_buf = s_in + (size_t)(ti & 1) * in_stride;
    T q0, q1, q2, d0, d1, d2;
    float xi32[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) xi32[k] = (float)J.xi_body[k];
    const int64_t srow = tile0 + li;   // selected-row index; raw index = srow * stride (PointBudget gather)
    double w_sel, w_time;
    if (P.layout == GCL_LAYOUT_REF_F64) {
      double px, py, pz, tt;
      if (from_bulk) {                            // arrays of rows
        px = in_buf[(size_t)li * 3]; py = in_buf[(size_t)li * 3 + 1]; pz = in_buf[(size_t)li * 3 + 2];
        tt = in_buf[(size_t)TP * 3 + li]; w_sel = in_buf[(size_t)TP * 4 + li];
      } else {
        const double* in = in_buf + (size_t)li * 5;  // {x, y, z, t, w}
        px = in[0]; py = in[1]; pz = in[2]; tt = in[3]; w_sel = in[4];
      }
      if (sizeof(T) == 8) {
        PointOut64 o;
        point_front64_fast(J, C, tconst, s_tab, px, py, pz, tt, o);
        d0 = (T)o.d[0]; d1 = (T)o.d[1]; d2 = (T)o.d[2];
        q0 = (T)o.p0[0]; q1 = (T)o.p0[1]; q2 = (T)o.p0[2];
        w_time = o.w_time;
      } else {
        PointOut32 o;
        point_front32(J, C, xi32, (float)px, (float)py, (float)pz, (float)(tt - J.scan_start_time), o);
        d0 = (T)o.d[0]; d1 = (T)o.d[1]; d2 = (T)o.d[2];
        q0 = (T)o.p0[0]; q1 = (T)o.p0[1]; q2 = (T)o.p0[2];
        w_time = (double)o.w_time;
      }
    } else {
      const float* base = reinterpret_cast<const float*>(in_buf);
      const float4 xyzt = *reinterpret_cast<const float4*>(base + (size_t)li * 4);
      w_sel = (double)base[(size_t)TP * 4 + li];
      PointOut32 o;
      point_front32(J, C, xi32, xyzt.x, xyzt.y, xyzt.z, xyzt.w, o);
      d0 = (T)o.d[0]; d1 = (T)o.d[1]; d2 = (T)o.d[2];
      q0 = (T)o.p0[0]; q1 = (T)o.p0[1]; q2 = (T)o.p0[2];
      w_time = (double)o.w_time;
    }
    const double lam = J.point_lambda ? J.point_lambda[srow] : 1.0;
    const double w_desk = w_sel * w_time;
    CERT(GCL_CS_SUM_W_SEL) += w_sel; CERT(GCL_CS_SUM_W_SEL_SQ) += w_sel * w_sel; CERT(GCL_CS_SUM_W_DESKEW) += w_desk;
    CERT(GCL_CS_ROWS) += 1.0;
    // Shift = the largest possible logit (|d.b| <= 1) when that cannot underflow, else the true row maximum
    // (jax.nn.softmax, binning.py:69).
    const T offv = (P.fixed_shift & 1) ? -cscale : -sim_rowmax(s_atlas, B, d0, d1, d2);
    T* sd = s_d + (size_t)li * 4;
    sd[0] = d0; sd[1] = d1; sd[2] = d2; sd[3] = offv;
    T* pt = s_pt + (size_t)(two_groups ? (ti & 1) : 0) * 7 * TP + li;   // scalar-major: conflict-free
    pt[0] = q0; pt[TP] = q1; pt[2 * TP] = q2; pt[3 * TP] = d0; pt[4 * TP] = d1; pt[5 * TP] = d2; pt[6 * TP] = (T)(w_desk * lam);
  };
  // ======================================= scale =============================================================
  // thread li of tile ti's group: Z_i, certificate partials, row (w_i/Z_i) phi_i for phase B
  auto scale = [&](int ti, int li) {
    T Z = 0, S1 = 0, xmax = (T)-1.0e30;   // fp32: largest logit; fp64: largest exponential (>= 0)
    for (int a = 0; a < TA; ++a) {
      const T* zp = s_zp + ((size_t)a * TP + li) * 3;
      Z += zp[0]; S1 += zp[1]; xmax = zp[2] > xmax ? zp[2] : xmax;
    }
    // H_i = -sum_b r log(r + eps) = log Z - S1/Z - B*eps + O(B eps^2 / r)   (binning.py:70-72); S1 in logit units
    const double Zd = (double)Z;
    if (sizeof(T) == 4) {
      const float invZ = 1.f / (float)Z;
      CERT(GCL_CS_SUM_ENTROPY) += (double)(__logf((float)Z) - (float)Num<T>::kToNat * (float)S1 * invZ) - (double)B * C.eps_mass;
      const double mr = (double)(fast_exp2((float)xmax) * invZ);
      if (mr > CERT(GCL_CS_MAX_RESP)) CERT(GCL_CS_MAX_RESP) = mr;
    } else {
      const double rZ = 1.0 / Zd;
      CERT(GCL_CS_SUM_ENTROPY) += log(Zd) - Num<T>::kToNat * (double)S1 * rZ - (double)B * C.eps_mass;
      const double mr = (double)xmax * rZ;   // fp64 phase A hands over max_b e itself
      if (mr > CERT(GCL_CS_MAX_RESP)) CERT(GCL_CS_MAX_RESP) = mr;
    }
    const T* pt = s_pt + (size_t)(two_groups ? (ti & 1) : 0) * 7 * TP + li;
    const T q0 = pt[0], q1 = pt[TP], q2 = pt[2 * TP], d0 = pt[3 * TP], d1 = pt[4 * TP], d2 = pt[5 * TP], wq = pt[6 * TP];
    const T w = wq / Z;
    T f[kRowPitch];
    f[0] = w;
    f[1] = w * d0; f[2] = w * d1; f[3] = w * d2;
    f[4] = f[1] * d0; f[5] = f[1] * d1; f[6] = f[1] * d2; f[7] = f[2] * d1; f[8] = f[2] * d2; f[9] = f[3] * d2;
    f[10] = w * q0; f[11] = w * q1; f[12] = w * q2;
    f[13] = f[10] * q0; f[14] = f[10] * q1; f[15] = f[10] * q2; f[16] = f[11] * q1; f[17] = f[11] * q2; f[18] = f[12] * q2;
    f[19] = 0;
    T* row = s_rows + (size_t)li * RP;
    if (MMA) {   // the padding rows of the third moment fragment must be zeros (they are multiplied like the others)
      double2* r2 = reinterpret_cast<double2*>(row);
      r2[10] = make_double2(0.0, 0.0); r2[11] = make_double2(0.0, 0.0);
    }
    if (sizeof(T) == 4) {
      float4* r4 = reinterpret_cast<float4*>(row);
#pragma unroll
      for (int k = 0; k < 5; ++k) r4[k] = make_float4((float)f[4 * k], (float)f[4 * k + 1], (float)f[4 * k + 2], (float)f[4 * k + 3]);
    } else {
      double2* r2 = reinterpret_cast<double2*>(row);
#pragma unroll
      for (int k = 0; k < 10; ++k) r2[k] = make_double2((double)f[2 * k], (double)f[2 * k + 1]);
    }
  };
  // group-local index of this thread in the group of tile ti, or -1 when it is not in it
  auto group_index = [&](int ti) -> int {
    const int base = two_groups ? (ti & 1) * TP : 0;
    return (tid >= base && tid < base + TP) ? tid - base : -1;
  };

  // prologue: rows of tiles 0 and 1 on their way
  {
    const int64_t t1 = lo + TP;
    if (tid == kNT - 32) {
      if (lo < hi && tile_is_bulk(lo)) prefetch_bulk(lo, 0);
      if (t1 < hi && tile_is_bulk(t1)) prefetch_bulk(t1, 1);
    }
    const int l0 = group_index(0), l1 = group_index(1);
    if (l0 >= 0) prefetch_rows(lo, 0, l0);
    if (l1 >= 0) prefetch_rows(t1, 1, l1);
  }
  if (MMA == 2 && NT_ == 128 && P.stagger_ns > 0) {
    // the second CTA to arrive on an SM (odd arrival count: the counter only ever grows, two co-resident CTAs draw consecutive numbers)
    // starts part of a tile late, so that one CTA's phase A meets the other's phase B from then on
    __shared__ unsigned int s_arrival;
    if (tid == 0) {
      unsigned int smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      s_arrival = atomicAdd(&g_sm_arrivals[smid & 255], 1u);
    }
    __syncthreads();
    if (s_arrival & 1u) {
      const long long t_end = clock64() + (long long)P.stagger_ns * 2;   // ~2 cycles per ns
      while (clock64() < t_end) __nanosleep(100);
    }
    __syncthreads();
  }
  K1T(0);   // prologue: job record, atlas, exp table, raw-weight sum

  int pts_since_flush = 0;
  const int flush_points = (sizeof(T) == 4) ? 512 * PS : (1 << 30);  // per-thread chain length ~512 in fp32

  // iteration ti = -1 is the front end of the first tile alone (one copy of the per-point code in the kernel: it is large)
  const int n_tiles = (int)((hi - lo + TP - 1) / TP);
  for (int ti = -1; ti < n_tiles; ++ti) {
    const int64_t tile0 = lo + (int64_t)ti * TP;
    const int count = ti < 0 ? 0 : (int)((hi - tile0 < TP) ? hi - tile0 : TP);
    // ======================================= phase A ===========================================================
    if (ti >= 0 && MMA == 2) {
      if constexpr (sizeof(T) == 8) {
        const int w = tid >> 5, lane = tid & 31, p = 8 * w + (lane >> 2), a = lane & 3;
        const double afrag = s_d[(size_t)p * 4 + a];               // {d0, d1, d2, shift}[a] of point p
        const double* bf = s_atlas + (size_t)(lane >> 2) * 4 + a;  // + 32 per 8-bin fragment
        double* er = s_etile + (size_t)p * EP + 2 * a;             // + 8 per fragment
        const bool live = p < count;
        double z = 0.0, s1 = 0.0;
        unsigned long long m = 0ull;
        const int nt_full = B >> 3, nt_all = (B + 7) >> 3;
        constexpr int NGM = GCL_A_NGMMA;
        int nt = 0;
        for (; nt + NGM <= nt_full; nt += NGM) phase_a_mma_groups<NGM, false>(bf + 32 * nt, s_tab, afrag, 8 * nt + 2 * a, B, live, er + 8 * nt, z, s1, m);
        for (; nt < nt_full; ++nt) phase_a_mma_groups<1, false>(bf + 32 * nt, s_tab, afrag, 8 * nt + 2 * a, B, live, er + 8 * nt, z, s1, m);
        for (; nt < nt_all; ++nt) phase_a_mma_groups<1, true>(bf + 32 * nt, s_tab, afrag, 8 * nt + 2 * a, B, live, er + 8 * nt, z, s1, m);
        T* zp = s_zp + ((size_t)a * TP + p) * 3;
        zp[0] = z; zp[1] = s1; zp[2] = __longlong_as_double((long long)m);
      }
      __syncthreads();
    } else if (ti >= 0) {
      if (pi < count) {
        const T* sd = s_d + (size_t)pi * 4;
        T Z, S1, xmax;
        constexpr int NGB = sizeof(T) == 4 ? 3 : (MMA ? (NT_ == 512 ? 2 : GCL_A_NGMMA) : (G >= 4 ? GCL_A_NG4 : 4));
        if (sizeof(T) == 4 || (P.fixed_shift & 1))
          phase_a_bins<NGB, false>(s_atlas, s_tab, pa * BA, (pa + 1) * BA, B, sd[0], sd[1], sd[2], sd[3], s_etile + (size_t)pi * EP, Z, S1, xmax);
        else
          phase_a_bins<1, true>(s_atlas, s_tab, pa * BA, (pa + 1) * BA, B, sd[0], sd[1], sd[2], sd[3], s_etile + (size_t)pi * EP, Z, S1, xmax);
        T* zp = s_zp + ((size_t)pa * TP + pi) * 3;
        zp[0] = Z; zp[1] = S1; zp[2] = xmax;
      } else if (MMA && pi < ((count + 3) & ~3)) {
        // DMMA steps take four points at a time: the rows that pad a CTA's last tile multiply as zeros (stale or uninitialised
        // shared memory could hold anything, NaN bit patterns included)
        T* er = s_etile + (size_t)pi * EP;
        for (int b = pa * BA; b < (pa + 1) * BA; ++b) er[b] = (T)0;
      }
      __syncthreads();
    }
    K1T(2);   // phase A
    // =========================== scale of this tile | front end of the next | rows of the tile after that ======================
    {
      const int64_t t1 = tile0 + TP, t2 = tile0 + 2 * (int64_t)TP;
      if (ti >= 0) {
        if (tid == kNT - 32 && t2 < hi && tile_is_bulk(t2)) prefetch_bulk(t2, ti + 2);   // its buffer is free since the front end of this tile
        const int ls = group_index(ti);
        if (ls >= 0) {
          prefetch_rows(t2, ti + 2, ls);               // same group, same buffer as this tile
          if (ls < count) scale(ti, ls);
          else if (MMA && ls < ((count + 3) & ~3)) {
            double2* r2 = reinterpret_cast<double2*>(s_rows + (size_t)ls * RP);
#pragma unroll
            for (int k = 0; k < RP / 2; ++k) r2[k] = make_double2(0.0, 0.0);
          }
        }
      }
      const int lf = group_index(ti + 1);
      if (lf >= 0 && t1 < hi) front(t1, ti + 1, lf, !two_groups);
    }
    __syncthreads();
    K1T(3);   // scale + next front end
    if (ti < 0) continue;
    // ======================================= phase B ===========================================================
    if (MMA) {
      const int w = tid >> 5, lane = tid & 31;
      const double* ap = reinterpret_cast<const double*>(s_rows) + (size_t)(lane & 3) * RP + (lane >> 2);
      const double* bp = reinterpret_cast<const double*>(s_etile) + (size_t)(lane & 3) * EP + 8 * NTW * w + (lane >> 2);
      const int nks = (count + 3) >> 2;
#pragma unroll 2
      for (int ks = 0; ks < nks; ++ks) {
        double a[3], b[NTW];
#pragma unroll
        for (int mt = 0; mt < 3; ++mt) a[mt] = ap[(size_t)ks * 4 * RP + 8 * mt];
#pragma unroll
        for (int nt = 0; nt < NTW; ++nt) b[nt] = bp[(size_t)ks * 4 * EP + 8 * nt];
#pragma unroll
        for (int mt = 0; mt < 3; ++mt)
#pragma unroll
          for (int nt = 0; nt < NTW; ++nt) dmma884(cm[mt][nt][0], cm[mt][nt][1], a[mt], b[nt]);
      }
    } else {
      const T* ebase = s_etile + (size_t)binlane * VB;
#pragma unroll 2
      for (int pt = slice; pt < count; pt += PS) acc.step(ebase + (size_t)pt * EP, L, s_rows + (size_t)pt * kRowPitch);
    }
    K1T(4);   // phase B (this thread's share; the barrier below is charged to phase A of the next tile)
    pts_since_flush += count;
    __syncthreads();   // e-tile and rows are free for the next tile (and for the flush staging, which aliases the e-tile)
    if (pts_since_flush >= flush_points) {
      flush();
      pts_since_flush = 0;
    }
  }
  K1T(5);   // (the loop ends on a barrier)
  if (pts_since_flush > 0) flush();
  K1T(6);   // flush
  if (!flushed)   // a CTA without rows: its record is all zeros
    for (int i = tid; i < GCL_NM * B; i += kNT) part[i] = 0.0;

  // --- certificate scalars of this CTA -------------------------------------------------------------------------
#pragma unroll
  for (int k = 0; k < GCL_NCERT; ++k) {
    double v = CERT(k);
    for (int o = 16; o > 0; o >>= 1) {
      const double other = __shfl_xor_sync(0xffffffffu, v, o);
      v = (k == GCL_CS_MAX_RESP) ? fmax(v, other) : v + other;
    }
    if ((tid & 31) == 0) s_red[tid >> 5][k] = v;
  }
  __syncthreads();
  if (tid < GCL_NCERT) {
    // at stride 1 the raw-weight sum IS the selected-weight sum (same numbers, same order)
    const int src = (tid == GCL_CS_SUM_W_RAW && stride == 1) ? GCL_CS_SUM_W_SEL : tid;
    double v = s_red[0][src];
    for (int w = 1; w < kNT / 32; ++w) v = (src == GCL_CS_MAX_RESP) ? fmax(v, s_red[w][src]) : v + s_red[w][src];
    part[(size_t)GCL_NM * B + tid] = v;
  }
#undef CERT
#ifdef GCL_K1_TRACE
  K1T(7);   // certificate reduction
  if (threadIdx.x == 0 && blockIdx.x == 1 && (blockIdx.y == 7 || gridDim.y == 1))
    for (int i = 0; i < 8; ++i) g_k1_trace[i] = k1_acc[i];
#endif
}

// ------------------------------------------------------------------------------------------------------------
// K1-WS: the fp64 kernel, warp-specialised (129..256 bins, float64 rows, fixed softmax shift)
// ------------------------------------------------------------------------------------------------------------
// The kernel above runs its phases one after another on the same eight warps.  Phase B is FP64-pipe bound, but phase A runs at 2.5x its
// FP64 bound (issue slots and latency: profiles/r2_scan_moments_fp64_1Mx256.md), and through the scale interval six of the
// eight warps idle.  Here the phases run BESIDE each other on different warps:
//   warps 0-7   consumers: phase B of tile t (G = 4 bins x 19 accumulators per thread, as above) and the final flush; 208 registers
//   warps 8-10  phase-A warps: lanes = the 32 points of tile t + 1, warp = a third of the bins; exponentials into the OTHER e-tile buffer
//   warp 11     point warp: front end of tile t + 2, scale step of tile t + 1, raw rows staged two tiles ahead (cp.async)
// Registers move between the roles with setmaxnreg (the accumulating warps take 208, the producers 88: 256 x 208 + 128 x 88 = 384 x 168, the allocation of the launch -- the pool setmaxnreg draws on).
// Hand-offs are named barriers: FULL[buf] (point warp arrives, consumers wait), EMPTY[buf] (consumers arrive, phase-A warps wait),
// PROD (the four producer warps, once per tile: phase A of tile t done, front end of tile t + 1 published).  Same partial-record
// format and the same fixed summation order inside a CTA from run to run.
constexpr int kWsTP = 32;     // points per tile
enum { kBarFull0 = 1, kBarEmpty0 = 3, kBarProd = 5, kBarCons = 6 };

__device__ __forceinline__ void nb_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void nb_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

struct WsPlan { int BA, EP; size_t o_tab, o_sd, o_zp, o_rows, o_pt, o_in, o_etile, total; };
__host__ __device__ inline WsPlan ws_plan(int B, int G, int L, int kWsCons, int kWsTA, int nga) {
  WsPlan p;
  p.BA = (((B + kWsTA - 1) / kWsTA) + 2 * nga - 1) / (2 * nga) * (2 * nga);   // whole phase-A steps of nga groups of two bins
  const int cols = (G * L > kWsTA * p.BA) ? G * L : kWsTA * p.BA;
  p.EP = e_pitch(cols, 4, 2);
  size_t off = (size_t)(kWsTA * p.BA) * 3 * 8;
  off = (off + 15) & ~(size_t)15;
  p.o_tab = off;  off += (size_t)kExpTab * 8;
  p.o_sd = off;   off += (size_t)2 * kWsTP * 4 * 8;
  p.o_zp = off;   off += (size_t)2 * kWsTA * kWsTP * 3 * 8;
  off = (off + 15) & ~(size_t)15;
  p.o_rows = off; off += (size_t)2 * kWsTP * kRowPitch * 8;
  p.o_pt = off;   off += (size_t)3 * 7 * kWsTP * 8;   // three buffers: the front end of tile t + 2 may run beside the scale step of tile t
  p.o_in = off;   off += (size_t)2 * kWsTP * 5 * 8;
  off = (off + 15) & ~(size_t)15;
  p.o_etile = off;
  size_t et = (size_t)2 * kWsTP * p.EP * 8;
  const size_t stage = (size_t)(kWsCons / L) * kKC * (size_t)(G * L) * 8;
  if (stage > et) et = stage;
  p.total = off + et + 16;
  return p;
}

// NCW consumer warps, TA phase-A warps (TA threads per point), one point warp; RC / RA registers of the consumers / the phase-A warps
// (setmaxnreg; the point warp follows the phase-A warps when it shares their warpgroup, else it keeps the launch allocation);
// NGA groups of two bins in flight per phase-A thread
#ifdef GCL_K1_TRACE
__device__ long long g_ws_trace[16];
extern "C" int gcl_debug_ws_trace(long long* out) { return (int)cudaMemcpyFromSymbol(out, g_ws_trace, sizeof(long long) * 16); }
#define WST_DECL long long wst_t = clock64(), wst_acc[4] = {0, 0, 0, 0}; const bool wst_on = (blockIdx.x == 1 && (threadIdx.x & 31) == 0)
#define WST(i) do { if (wst_on) { const long long n_ = clock64(); wst_acc[i] += n_ - wst_t; wst_t = n_; } } while (0)
#define WST_OUT(base) do { if (wst_on) for (int i_ = 0; i_ < 4; ++i_) g_ws_trace[(base) + i_] = wst_acc[i_]; } while (0)
#else
#define WST_DECL do { } while (0)
#define WST(i) do { } while (0)
#define WST_OUT(base) do { } while (0)
#endif
template <int G, int NCW, int TA_, int NPW, int RC, int RA, int NGA>
__global__ void __launch_bounds__(32 * (NCW + TA_ + NPW), 1) scan_moments_ws64_kernel(const __grid_constant__ ScanParams P) {
  using T = double;
  constexpr int kWsCons = 32 * NCW, kWsTA = TA_, kWsProd = 32 * (TA_ + NPW), kWsNT = kWsCons + kWsProd;
  static_assert(NCW % 4 == 0 && (TA_ + NPW) % 4 == 0 && (NPW == 1 || NPW == 2), "setmaxnreg works on warpgroups of four warps");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int B = P.n_bins, L = P.lanes;
  __shared__ gcl_job s_job;
  __shared__ double s_wraw[kWsNT];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid < (int)(sizeof(gcl_job) / sizeof(int)))
    reinterpret_cast<int*>(&s_job)[tid] = reinterpret_cast<const int*>(P.jobs ? &P.jobs[blockIdx.y] : &P.job)[tid];
  __syncthreads();
  const gcl_job& J = s_job;
  const ScanConst& C = P.c;
  const int64_t n_raw = J.n_raw;
  const int64_t cap = J.n_points_cap > 0 ? J.n_points_cap : 1;
  const int64_t stride0 = (n_raw + cap - 1) / cap;
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
  constexpr int TP = kWsTP, TA = kWsTA;
  const WsPlan pl = ws_plan(B, G, L, kWsCons, kWsTA, NGA);
  const int BA = pl.BA, EP = pl.EP, PS = kWsCons / L, Bp = G * L;
  T* s_atlas = reinterpret_cast<T*>(smem_raw);
  double* s_tab = reinterpret_cast<double*>(smem_raw + pl.o_tab);
  T* s_d = reinterpret_cast<T*>(smem_raw + pl.o_sd);         // [buf][TP][4]
  T* s_zp = reinterpret_cast<T*>(smem_raw + pl.o_zp);        // [buf][TA][TP][3]
  T* s_rows = reinterpret_cast<T*>(smem_raw + pl.o_rows);    // [buf][TP][kRowPitch]
  T* s_pt = reinterpret_cast<T*>(smem_raw + pl.o_pt);        // [tile % 3][7][TP]
  double* s_in = reinterpret_cast<double*>(smem_raw + pl.o_in);   // [buf][TP][5]
  T* s_etile = reinterpret_cast<T*>(smem_raw + pl.o_etile);  // [buf][TP][EP]
  T* s_stage = s_etile;                                      // flush staging: after the last tile
  const double cscale_d = Num<T>::kScale / P.tau;

  {  // atlas (pre-scaled, group-major as above) and exp table
    const int nb = TA * BA;
    constexpr int V = Num<T>::kVec;
    for (int i = tid; i < nb * 3; i += kWsNT) {
      const int g = i / (3 * V), rem = i - g * 3 * V, c = rem / V, b = g * V + (rem - c * V);
      s_atlas[i] = (b < B) ? (T)(P.bin_dirs[b * 3 + c] * cscale_d) : (T)0;
    }
    for (int i = tid; i < kExpTab; i += kWsNT) s_tab[i] = exp2((double)i * (1.0 / kExpTab));
  }
  {  // sum of ALL raw weights of this CTA's raw range (mass preservation, point_budget.py:69); partials combined in a fixed order
    double p0 = 0, p1 = 0, p2 = 0, p3 = 0;
    if (stride > 1 && hi > lo) {
      const int64_t rlo = lo * stride;
      const int64_t rhi = (hi >= n_sel_total) ? n_raw : hi * stride;
      const double* w = static_cast<const double*>(J.weights);
      int64_t r = rlo + tid;
      for (; r + 3 * kWsNT < rhi; r += 4 * kWsNT) { p0 += w[r]; p1 += w[r + kWsNT]; p2 += w[r + 2 * kWsNT]; p3 += w[r + 3 * kWsNT]; }
      for (; r < rhi; r += kWsNT) p0 += w[r];
    }
    s_wraw[tid] = (p0 + p1) + (p2 + p3);
  }
  __syncthreads();
  const int n_tiles = (int)((hi - lo + TP - 1) / TP);

  if (warp < kWsCons / 32) {
    // ================================================ consumers: phase B ========================================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(RC));
    constexpr int VB = Tile<T, G>::VB;
    const int binlane = tid & (L - 1), slice = tid / L;
    Tile<T, G> acc;
    acc.zero();
    WST_DECL;
    for (int t = 0; t < n_tiles; ++t) {
      const int buf = t & 1;
      const int64_t tile0 = lo + (int64_t)t * TP;
      const int count = (int)((hi - tile0 < TP) ? hi - tile0 : TP);
      WST(0);
      nb_sync(kBarFull0 + buf, kWsCons + 32);
      WST(1);
      const T* ebase = s_etile + (size_t)buf * TP * EP + (size_t)binlane * VB;
      const T* rws = s_rows + (size_t)buf * TP * kRowPitch;
#pragma unroll 2
      for (int pt = slice; pt < count; pt += PS) acc.step(ebase + (size_t)pt * EP, L, rws + (size_t)pt * kRowPitch);
      if (t + 2 < n_tiles) nb_arrive(kBarEmpty0 + buf, kWsCons + 32 * TA);
    }
    WST(0);
    if (warp == 0) WST_OUT(0);   // consumer warp 0: [0] phase B, [1] wait FULL
    // flush: fixed-order cross-slice sum, kKC moments at a time (the producers are done with the e-tile buffers)
    if (n_tiles > 0) {
#pragma unroll
      for (int k0 = 0; k0 < GCL_NM; k0 += kKC) {
#pragma unroll
        for (int kc = 0; kc < kKC; ++kc) {
          if (k0 + kc < GCL_NM) {
#pragma unroll
            for (int g = 0; g < G; ++g)
              s_stage[((size_t)slice * kKC + kc) * Bp + (g / VB) * VB * L + VB * binlane + (g % VB)] = acc.get(g, (k0 + kc < GCL_NM) ? k0 + kc : 0);
          }
        }
        nb_sync(kBarCons, kWsCons);
        for (int e = tid; e < kKC * Bp; e += kWsCons) {
          const int kc = e / Bp, b = e - kc * Bp, k = k0 + kc;
          if (k < GCL_NM && b < B) {
            double sm = 0.0;
            for (int sl = 0; sl < PS; ++sl) sm += (double)s_stage[((size_t)sl * kKC + kc) * Bp + b];
            part[(size_t)k * B + b] = sm;
          }
        }
        nb_sync(kBarCons, kWsCons);
      }
    } else {
      for (int i = tid; i < GCL_NM * B; i += kWsCons) part[i] = 0.0;
    }
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(RA));
    if (warp < kWsCons / 32 + TA) {
      // ============================================ phase-A warps =================================================================
      const int pa = warp - kWsCons / 32, pi = lane;
      nb_sync(kBarProd, kWsProd);   // front end of tile 0 published
      WST_DECL;
      for (int t = 0; t < n_tiles; ++t) {
        const int buf = t & 1;
        const int64_t tile0 = lo + (int64_t)t * TP;
        const int count = (int)((hi - tile0 < TP) ? hi - tile0 : TP);
        WST(2);
        if (t >= 2) nb_sync(kBarEmpty0 + buf, kWsCons + 32 * TA);   // the consumers have finished tile t - 2 in this buffer
        WST(0);
        if (pi < count) {
          const T* sd = s_d + ((size_t)buf * TP + pi) * 4;
          T Z, S1, xmax;
          phase_a_bins<NGA, false>(s_atlas, s_tab, pa * BA, (pa + 1) * BA, B, sd[0], sd[1], sd[2], sd[3],
                                 s_etile + ((size_t)buf * TP + pi) * EP, Z, S1, xmax);
          T* zp = s_zp + (((size_t)buf * TA + pa) * TP + pi) * 3;
          zp[0] = Z; zp[1] = S1; zp[2] = xmax;
        }
        WST(1);
        __threadfence_block();
        nb_sync(kBarProd, kWsProd);
      }
      WST(2);
      if (pa == 0) WST_OUT(4);   // phase-A warp 0: [0] wait EMPTY, [1] phase A, [2] wait PROD
    } else {
      // ============================ point warp(s): front end + row staging | scale (one warp, or one warp each) ====================
      const bool do_front = (NPW == 1) || warp == kWsCons / 32 + TA, do_scale = (NPW == 1) || !do_front;
      const int li = lane;
      const TimeConst tconst = time_const(J);
      double c_sel = 0, c_sel2 = 0, c_desk = 0, c_rows = 0, c_ent = 0, c_max = 0, c_wraw = 0;
      for (int j = 0; j < kWsNT / 32; ++j) c_wraw += s_wraw[li + 32 * j];
      auto prefetch_rows = [&](int u) {   // tile u -> staging buffer u & 1; always commits a group
        const int64_t srow = lo + (int64_t)u * TP + li;
        if (u < n_tiles && srow < hi) {
          const int64_t r = srow * stride;
          const int64_t rp = (J.flags & GCL_JOB_COMPACT_ROWS) ? srow : r;
          const double* pts = static_cast<const double*>(J.points) + rp * 3;
          double* dst = s_in + ((size_t)(u & 1) * TP + li) * 5;
          cp_async8(dst + 0, pts + 0); cp_async8(dst + 1, pts + 1); cp_async8(dst + 2, pts + 2);
          cp_async8(dst + 3, static_cast<const double*>(J.timestamps) + rp);
          cp_async8(dst + 4, static_cast<const double*>(J.weights) + r);
        }
        cp_async_commit();
      };
      auto front = [&](int u) {           // tile u: raw rows -> {d, shift} for phase A, point state for the scale step
        cp_async_wait_group<1>();
        __syncwarp();
        const int64_t tile0 = lo + (int64_t)u * TP;
        const int count = (int)((hi - tile0 < TP) ? hi - tile0 : TP);
        const int buf = u & 1;
        if (li < count) {
          const double* in = s_in + ((size_t)buf * TP + li) * 5;
          const double w_sel = in[4];
          PointOut64 o;
          point_front64_fast(J, C, tconst, s_tab, in[0], in[1], in[2], in[3], o);
          const int64_t srow = tile0 + li;
          const double lam = J.point_lambda ? J.point_lambda[srow] : 1.0;
          const double w_desk = w_sel * o.w_time;
          c_sel += w_sel; c_sel2 += w_sel * w_sel; c_desk += w_desk; c_rows += 1.0;
          T* sd = s_d + ((size_t)buf * TP + li) * 4;
          sd[0] = o.d[0]; sd[1] = o.d[1]; sd[2] = o.d[2]; sd[3] = -cscale_d;   // fixed shift: the largest possible logit
          T* pt = s_pt + (size_t)(u % 3) * 7 * TP + li;
          pt[0] = o.p0[0]; pt[TP] = o.p0[1]; pt[2 * TP] = o.p0[2]; pt[3 * TP] = o.d[0]; pt[4 * TP] = o.d[1]; pt[5 * TP] = o.d[2];
          pt[6 * TP] = w_desk * lam;
        }
        __syncwarp();
        prefetch_rows(u + 2);             // this tile's staging buffer is free again
      };
      if (do_front) {
        prefetch_rows(0);
        prefetch_rows(1);
        if (n_tiles > 0) front(0);
      }
      __threadfence_block();
      nb_sync(kBarProd, kWsProd);
      WST_DECL;
      for (int t = 0; t < n_tiles; ++t) {
        const int buf = t & 1;
        const int64_t tile0 = lo + (int64_t)t * TP;
        const int count = (int)((hi - tile0 < TP) ? hi - tile0 : TP);
        WST(2);
        if (do_front && t + 1 < n_tiles) front(t + 1);
        WST(0);
        __threadfence_block();
        nb_sync(kBarProd, kWsProd);       // phase A of tile t is complete (and the front end of tile t + 1 is visible to the phase-A warps)
        WST(1);
        if (do_scale && li < count) {
          T Z = 0, S1 = 0, xmax = 0;
          for (int a = 0; a < TA; ++a) {
            const T* zp = s_zp + (((size_t)buf * TA + a) * TP + li) * 3;
            Z += zp[0]; S1 += zp[1]; xmax = zp[2] > xmax ? zp[2] : xmax;
          }
          const double rZ = 1.0 / Z;
          c_ent += log(Z) - Num<T>::kToNat * S1 * rZ - (double)B * C.eps_mass;   // binning.py:70-72
          const double mr = xmax * rZ;
          c_max = mr > c_max ? mr : c_max;
          const T* pt = s_pt + (size_t)(t % 3) * 7 * TP + li;
          const T q0 = pt[0], q1 = pt[TP], q2 = pt[2 * TP], d0 = pt[3 * TP], d1 = pt[4 * TP], d2 = pt[5 * TP], wq = pt[6 * TP];
          const T w = wq / Z;
          T f[kRowPitch];
          f[0] = w;
          f[1] = w * d0; f[2] = w * d1; f[3] = w * d2;
          f[4] = f[1] * d0; f[5] = f[1] * d1; f[6] = f[1] * d2; f[7] = f[2] * d1; f[8] = f[2] * d2; f[9] = f[3] * d2;
          f[10] = w * q0; f[11] = w * q1; f[12] = w * q2;
          f[13] = f[10] * q0; f[14] = f[10] * q1; f[15] = f[10] * q2; f[16] = f[11] * q1; f[17] = f[11] * q2; f[18] = f[12] * q2;
          f[19] = 0;
          double2* r2 = reinterpret_cast<double2*>(s_rows + ((size_t)buf * TP + li) * kRowPitch);
#pragma unroll
          for (int k = 0; k < 10; ++k) r2[k] = make_double2(f[2 * k], f[2 * k + 1]);
        }
        if (do_scale) {
          __threadfence_block();
          nb_arrive(kBarFull0 + buf, kWsCons + 32);
        }
      }
      WST(2);
      if (do_front) WST_OUT(8); else WST_OUT(12);   // point warp(s): [0] front end, [1] wait PROD, [2] scale
      if (do_front) cp_async_wait_group<0>();
      // certificate scalars of this CTA: fixed shuffle tree over the 32 lanes
      double cv[GCL_NCERT];
      cv[GCL_CS_SUM_W_RAW] = (stride == 1) ? c_sel : c_wraw;   // at stride 1 the raw-weight sum IS the selected-weight sum
      cv[GCL_CS_SUM_W_SEL] = c_sel; cv[GCL_CS_SUM_W_SEL_SQ] = c_sel2; cv[GCL_CS_SUM_W_DESKEW] = c_desk;
      cv[GCL_CS_SUM_ENTROPY] = c_ent; cv[GCL_CS_MAX_RESP] = c_max; cv[GCL_CS_ROWS] = c_rows; cv[GCL_CS_SPARE] = 0.0;
#pragma unroll
      for (int k = 0; k < GCL_NCERT; ++k) {
        double v = cv[k];
        for (int o = 16; o > 0; o >>= 1) {
          const double other = __shfl_xor_sync(0xffffffffu, v, o);
          v = (k == GCL_CS_MAX_RESP) ? fmax(v, other) : v + other;
        }
        const bool scale_owns = (k == GCL_CS_SUM_ENTROPY || k == GCL_CS_MAX_RESP);
        if (li == 0 && (NPW == 1 || scale_owns == do_scale)) part[(size_t)GCL_NM * B + k] = v;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// launch
// ------------------------------------------------------------------------------------------------------------
static inline int pow2_ceil(int x) { int p = 1; while (p < x) p <<= 1; return p; }

// G bins per phase-B thread, L = pow2 bin lanes with G*L >= B; least padding first, then the larger G.
// (Two CTAs per SM at <= 128 registers -- G = 4 fp32 / 2 fp64 -- were measured slower than one CTA with the larger tile:
// 470 vs 420 us fp32 and 1470 vs 1200 us fp64 at 1M x 256; the smaller tile doubles the row traffic per multiply-add.)
struct FusedCfg { int G, minb; };
void plan_fused_bins(int n_bins, int precision, int* G_out, int* L_out, int* minb_out) {
  static const FusedCfg c32[] = {{8, 1}, {6, 1}, {4, 1}}, c64[] = {{4, 1}, {3, 1}, {2, 1}};
  const FusedCfg* cand = precision == GCL_F64 ? c64 : c32;
  int force_g = 0;
  if (const char* e = getenv("GCL_K1_G")) force_g = atoi(e);   // tuning aid
  int bestG = cand[0].G, bestL = kNT, bestM = cand[0].minb, bestWaste = 1 << 30;
  for (int i = 0; i < 3; ++i) {
    const int G = cand[i].G;
    if (force_g && G != force_g) continue;
    const int L = pow2_ceil((n_bins + G - 1) / G);
    if (L > kNT) continue;
    const int waste = G * L - n_bins;
    if (waste < bestWaste) { bestWaste = waste; bestG = G; bestL = L; bestM = cand[i].minb; }
  }
  *G_out = bestG; *L_out = bestL; *minb_out = bestM;
}

// points per tile: as many as fit in shared memory (the kernel lowers it for CTAs that see few rows), at least 256/16
int plan_fused_tile(int n_bins, int precision, int G, int L, int minb, int mma = 0, int nt = 256) {
  const int tsz = precision == GCL_F64 ? 8 : 4;
  // 228 KB per SM, 1 KB reserved per CTA, ~17 KB static (certificate partials, reduction scratch; 16 KB more with 512 threads)
  const size_t budget = (minb == 2 ? 94 : (nt == 512 ? 190 : 200)) * 1024;
  // GCL_K1_TPMAX = 16..256 (a power of two): upper limit of the tile, for A/B runs (smaller tiles at few bins let the two thread groups
  // alternate and put several threads on a point in phase A)
  static const int tp_max = [] { const char* e = getenv("GCL_K1_TPMAX"); const int v = e ? atoi(e) : 0; return (v >= 16 && v <= kNT && !(v & (v - 1))) ? v : kNT; }();
  int TP = tp_max;
  while (TP > 16 && fused_plan(n_bins, tsz, G, L, TP, mma, nt).total > budget) TP >>= 1;
  return TP;
}

template <typename T, int G, int MINB, int MMA = 0, int NT_ = 256>
static cudaError_t launch_fused_one(const ScanParams& P, dim3 grid, size_t smem, cudaStream_t stream) {
  auto kern = scan_moments_fused_kernel<T, G, MINB, MMA, NT_>;
  static std::atomic<uint64_t> optin{0};  // per instantiation, one bit per device
  if (cudaError_t e = optin_dynamic_smem(kern, (NT_ == 512 ? 192 : 204) * 1024, optin)) return e;   // + the static part (34 / 17 KB) <= 227 KB
  kern<<<grid, NT_, smem, stream>>>(P);
  return cudaGetLastError();
}

int fused_ctas_per_sm(int n_bins, int precision) {
  int G, L, M;
  plan_fused_bins(n_bins, precision, &G, &L, &M);
  return M;
}

cudaError_t launch_scan_moments_fused(ScanParams& P, int n_jobs, int ctas_per_job, int precision, cudaStream_t stream) {
  // fp32c, up to 256 bins, fixed softmax shift: the tensor-core kernel (gcl_scan_tc.cu); GCL_K1_TC=0 keeps everything on the CUDA cores
  if (tc_path_supported(P.n_bins, precision, P.fixed_shift)) return launch_scan_moments_tc(P, n_jobs, ctas_per_job, stream);
  int G, L, M;
  plan_fused_bins(P.n_bins, precision, &G, &L, &M);
  // fp64, 129..256 bins (G = 4 on 64 bin lanes): the DMMA variants of the kernel (mma.sync.m8n8k4.f64).  Measured at 1 M x 256 on 148 CTAs /
  // on the 131 072-row shard one rank of eight sees, same box, alternating runs:
  //   DFMA loops (MMA = 0)                                   889.3 us / 125.1 us
  //   phase B on DMMA, 8 warps (GCL_K1_MMA=1)                894.2 us / 121.1 us   (phase B -11 %, flush a plain store; phase A unchanged)
  //   phase B on DMMA, 16 warps (GCL_K1_MMA=2)               919 us
  //   phases A and B on DMMA (GCL_K1_MMA=3, the default)     841.1 us / 114.5 us   (the similarity dot as a K = 4 product: phase A is
  //                                                                                 bound by dispatch slots, and this takes 6 of 41 per pair)
  // The default applies with the fixed softmax shift only (tiny tau takes the row-maximum path of the DFMA kernel); GCL_K1_MMA=0 = never.
  static const int mma_mode = [] { const char* e = getenv("GCL_K1_MMA"); return e ? atoi(e) : 3; }();
  const int mma0 = (mma_mode > 0 && precision == GCL_F64 && G == 4 && L == 64) ? (mma_mode >= 3 ? ((P.fixed_shift & 1) ? 2 : 0) : 1) : 0;
  int mma = mma0;
  const int nt = (mma && mma_mode == 2) ? 512 : 256;
  int TP = plan_fused_tile(P.n_bins, precision, G, L, M, mma, nt);
  if (mma == 2 && TP != 64) { mma = 1; TP = plan_fused_tile(P.n_bins, precision, G, L, M, mma, nt); }   // (its phase A is laid out for 64-point tiles)
  P.lanes = L; P.tile_points = TP; P.n_chunks = 1;
  const size_t smem = fused_plan(P.n_bins, precision == GCL_F64 ? 8 : 4, G, L, TP, mma, nt).total;
  dim3 grid(ctas_per_job, n_jobs);
  // fp64, 129..256 bins (G = 4 on 64 bin lanes), float64 rows, fixed shift: the warp-specialised kernel on request (GCL_K1_WS = 1, 2, 3:
  // measured 873 / 937 / 961 us against 894 us at 1 M x 256 -- see DESIGN.md; not the default)
  static const int ws_mode = [] { const char* e = getenv("GCL_K1_WS"); return e ? atoi(e) : 0; }();
  if (ws_mode && precision == GCL_F64 && G == 4 && L == 64 && (P.fixed_shift & 1) && P.layout == GCL_LAYOUT_REF_F64) {
    static std::atomic<uint64_t> optin_ws{0}, optin_ws2{0};
    if (ws_mode == 2) {        // 8 accumulating warps, 3 phase-A warps + the point warp in one warpgroup
      auto kern = scan_moments_ws64_kernel<4, 8, 3, 1, 200, 104, 4>;
      if (cudaError_t e = optin_dynamic_smem(kern, 204 * 1024, optin_ws2)) return e;
      kern<<<grid, 32 * (8 + 3 + 1), ws_plan(P.n_bins, G, L, 256, 3, 4).total, stream>>>(P);
    } else if (ws_mode == 3) { // 4 accumulating warps, 6 phase-A warps, a front-end warp and a scale warp
      static std::atomic<uint64_t> optin_ws3{0};
      auto kern = scan_moments_ws64_kernel<4, 4, 6, 2, 208, 144, 4>;
      if (cudaError_t e = optin_dynamic_smem(kern, 204 * 1024, optin_ws3)) return e;
      kern<<<grid, 32 * (4 + 6 + 2), ws_plan(P.n_bins, G, L, 128, 6, 4).total, stream>>>(P);
    } else {                   // 4 accumulating warps (one per scheduler), 7 phase-A warps + the point warp (two producer warps per scheduler)
      auto kern = scan_moments_ws64_kernel<4, 4, 7, 1, 208, 144, 4>;
      if (cudaError_t e = optin_dynamic_smem(kern, 204 * 1024, optin_ws)) return e;
      kern<<<grid, 32 * (4 + 7 + 1), ws_plan(P.n_bins, G, L, 128, 7, 4).total, stream>>>(P);
    }
    return cudaGetLastError();
  }
  if (mma == 2 && mma_mode == 4) {
    // two CTAs of four warps per SM on 32-point tiles (a warp owns 64 bins in phase B): the phases of the two CTAs drift apart, so the
    // dispatch-bound phase A of one runs beside the pipe-bound phase B of the other.  The caller plans 2 x SMs CTAs per job for it.
    P.tile_points = 32;
    static const int stagger = [] { const char* e = getenv("GCL_K1_STAGGER_NS"); return e ? atoi(e) : 1800; }();
    P.stagger_ns = stagger;
    return launch_fused_one<double, 4, 2, 2, 128>(P, grid, fused_plan(P.n_bins, 8, G, L, 32, 2, 128).total, stream);
  }
  if (mma == 2) return launch_fused_one<double, 4, 1, 2, 256>(P, grid, smem, stream);
  if (mma) return nt == 512 ? launch_fused_one<double, 4, 1, 1, 512>(P, grid, smem, stream) : launch_fused_one<double, 4, 1, 1, 256>(P, grid, smem, stream);
  if (precision == GCL_F64) {
    switch (G) {
      case 2: return launch_fused_one<double, 2, 1>(P, grid, smem, stream);
      case 3: return launch_fused_one<double, 3, 1>(P, grid, smem, stream);
      default: return launch_fused_one<double, 4, 1>(P, grid, smem, stream);
    }
  }
  switch (G) {
    case 4: return launch_fused_one<float, 4, 1>(P, grid, smem, stream);
    case 6: return launch_fused_one<float, 6, 1>(P, grid, smem, stream);
    default: return launch_fused_one<float, 8, 1>(P, grid, smem, stream);
  }
}

}  // namespace gcl
