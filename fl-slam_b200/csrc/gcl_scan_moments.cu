// gcl_scan_moments.cu -- moment sums from MATERIALISED responsibilities (the stand-alone ScanBinMomentMatch operator,
// archive/legacy_operators/binning.py:212-324, with per-point covariances).  The fused hot path is gcl_scan_fused.cu; only the
// FROM_RESP instantiations of the kernel below are built.
//
// One pass over the raw points fuses: PointBudget gather (stride), constant-twist deskew + time-window weights,
// ray directions, softmax responsibilities over the bin atlas and the per-bin weighted moment sums
//   M_b[19] = sum_i w_i r_ib [1, d, ddT(6), p, ppT(6)]
// without materialising the (N,B) responsibilities or the (N,3,3) outer products the reference builds
// (archive/legacy_operators/binning.py:69,:171).
//
// Structure per CTA (NT threads, persistent over its slice of the selected rows), tile = NT points:
//   phase A  thread-per-point: deskew, window weight, direction, softmax normaliser Z_i over all bins (fp32: packed
//            over bin PAIRS with fma.rn.f32x2), entropy and max-responsibility certificate partials; writes one
//            shared-memory row per point: d_i, the exponent offset and (w_i/Z_i) * phi_i.
//   phase B  "GEMM by hand" on CUDA cores: thread = (bin lane, point slice) owns G bins x 19 accumulators in
//            registers.  Per point it reads the row with warp-broadcast LDS.128, recomputes
//            e_ib = exp2(d_i . b'_b + off_i) (atlas pre-scaled by log2(e)/tau) and does the 19 multiply-adds per bin as
//            nine packed fma.rn.f32x2 over moment PAIRS plus one scalar FFMA (fp32), or 19 DFMA (fp64).  The exponentials
//            of the next row are computed while the current row accumulates (software pipeline).
//   flush    per-thread partials are combined across point slices in a FIXED order and carried in fp64 in this CTA's
//            record of the partials buffer (no atomics => bit-reproducible run to run).  In fp32 mode a flush happens
//            at least every ~512 accumulated points per thread, which keeps the fp32 chains short ("fp32c").
// The contraction depth of the similarity is 3, so tensor cores are not used (BASELINE.json north_star).
#include "gcl_internal.cuh"
#include "gcl_point.cuh"

namespace gcl {

// ------------------------------------------------------------------------------------------------------------
// softmax normaliser of one point over all bins (phase A)
// ------------------------------------------------------------------------------------------------------------
// Logits arrive pre-scaled: the shared atlas holds b' = c b with c = log2(e)/tau (fp32, ex2.approx) or 1/tau (fp64,
// exp), so x_b = d . b'_b + off <= ~0.  Returns Z = sum_b e_b, S1 = sum_b e_b x_b (entropy), xmax = max_b x_b.
// The fma order (z, y, x onto the offset) is the same as in phase B, so both phases see bit-identical exponents.
__device__ __forceinline__ void softmax_stats(const float* sbA, int B, float d0, float d1, float d2, float off, float& Z,
                                              float& S1, float& xmax) {
  // sbA: per bin PAIR two float4: {bx0,bx1,by0,by1} {bz0,bz1,-,-}
  const float2 dx = make_float2(d0, d0), dy = make_float2(d1, d1), dz = make_float2(d2, d2), o2 = make_float2(off, off);
  float2 Z2 = make_float2(0.f, 0.f), S2 = make_float2(0.f, 0.f);
  float m = -3.0e38f;
  const float4* p = reinterpret_cast<const float4*>(sbA);
  const int npair = B >> 1;
#pragma unroll 4
  for (int bp = 0; bp < npair; ++bp) {
    const float4 a = p[2 * bp], c = p[2 * bp + 1];
    float2 s = ffma2(dz, make_float2(c.x, c.y), o2);
    s = ffma2(dy, make_float2(a.z, a.w), s);
    s = ffma2(dx, make_float2(a.x, a.y), s);
    const float2 e = make_float2(fast_exp2(s.x), fast_exp2(s.y));
    Z2 = fadd2(Z2, e);
    S2 = ffma2(e, s, S2);
    m = fmaxf(m, fmaxf(s.x, s.y));
  }
  Z = Z2.x + Z2.y;
  S1 = S2.x + S2.y;
  if (B & 1) {  // odd bin count: last bin sits in lane x of the last pair slot
    const float4 a = p[2 * npair], c = p[2 * npair + 1];
    const float sx = fmaf(d0, a.x, fmaf(d1, a.z, fmaf(d2, c.x, off)));
    const float e = fast_exp2(sx);
    Z += e;
    S1 = fmaf(e, sx, S1);
    m = fmaxf(m, sx);
  }
  xmax = m;
}
__device__ __forceinline__ void softmax_stats(const double* sbA, int B, double d0, double d1, double d2, double off, double& Z,
                                              double& S1, double& xmax) {
  // sbA: per bin four doubles {bx, by, bz, -}
  double z = 0.0, s1 = 0.0, m = -1.0e300;
#pragma unroll 2
  for (int b = 0; b < B; ++b) {
    const double2 a = *reinterpret_cast<const double2*>(sbA + 4 * b);
    const double bz = sbA[4 * b + 2];
    const double x = fma(d0, a.x, fma(d1, a.y, fma(d2, bz, off)));
    const double e = exp(x);
    z += e;
    s1 = fma(e, x, s1);
    m = fmax(m, x);
  }
  Z = z; S1 = s1; xmax = m;
}
// row maximum of the scaled similarity (only when the fixed shift could underflow)
__device__ __forceinline__ float sim_rowmax(const float* sbA, int B, float d0, float d1, float d2) {
  const float4* p = reinterpret_cast<const float4*>(sbA);
  float m = -3.0e38f;
  for (int b = 0; b < B; ++b) {
    const float4 a = p[2 * (b >> 1)], c = p[2 * (b >> 1) + 1];
    const float s = (b & 1) ? fmaf(d0, a.y, fmaf(d1, a.w, d2 * c.y)) : fmaf(d0, a.x, fmaf(d1, a.z, d2 * c.x));
    m = fmaxf(m, s);
  }
  return m;
}
__device__ __forceinline__ double sim_rowmax(const double* sbA, int B, double d0, double d1, double d2) {
  double m = -1.0e300;
  for (int b = 0; b < B; ++b) m = fmax(m, fma(d0, sbA[4 * b], fma(d1, sbA[4 * b + 1], d2 * sbA[4 * b + 2])));
  return m;
}

// ------------------------------------------------------------------------------------------------------------
// register state of one thread in phase B: G bins (their pre-scaled directions) x NMOM accumulators
// ------------------------------------------------------------------------------------------------------------
template <typename T, int G, int NMOM> struct Acc;

// fp32: bins are handled in PAIRS (g = 2*gp + half): the similarity, the exponent and all 19 multiply-adds of a pair
// issue as packed fma.rn.f32x2 with the point's values as the broadcast operand (SASS: FFMA2 R, R.F32, R.F32x2, R).
template <int G>
struct Acc<float, G, 19> {
  static_assert(G % 2 == 0, "fp32 path pairs bins");
  static constexpr int STRIDE = 28;  // floats per tile row: 24 used; 112 B rows keep 128-bit accesses conflict-free
  float2 bx[G / 2], by[G / 2], bz[G / 2];
  float2 a[G / 2][19];
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int gp = 0; gp < G / 2; ++gp)
#pragma unroll
      for (int k = 0; k < 19; ++k) a[gp][k] = make_float2(0.f, 0.f);
  }
  __device__ __forceinline__ void load_bins(const float* s_bins, int Btot, int base, int L) {
#pragma unroll
    for (int gp = 0; gp < G / 2; ++gp) {
      const int b0 = base + (2 * gp) * L, b1 = b0 + L;
      bx[gp] = make_float2(s_bins[b0], s_bins[b1]);
      by[gp] = make_float2(s_bins[Btot + b0], s_bins[Btot + b1]);
      bz[gp] = make_float2(s_bins[2 * Btot + b0], s_bins[2 * Btot + b1]);
    }
  }
  // one point: e = exp2(d . b' + off) for the G bins, then acc += e * phi
  __device__ __forceinline__ void step(const float* row) {
    const float4* r4 = reinterpret_cast<const float4*>(row);
    const float4 d = r4[0], f0 = r4[1], f1 = r4[2], f2 = r4[3], f3 = r4[4], f4 = r4[5];
    const float f[19] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w, f2.x, f2.y, f2.z, f2.w, f3.x, f3.y, f3.z, f3.w, f4.x, f4.y, f4.z};
#pragma unroll
    for (int gp = 0; gp < G / 2; ++gp) {
      float2 s = ffma2(make_float2(d.z, d.z), bz[gp], make_float2(d.w, d.w));
      s = ffma2(make_float2(d.y, d.y), by[gp], s);
      s = ffma2(make_float2(d.x, d.x), bx[gp], s);
      const float2 e = make_float2(fast_exp2(s.x), fast_exp2(s.y));
#pragma unroll
      for (int k = 0; k < 19; ++k) a[gp][k] = ffma2(e, make_float2(f[k], f[k]), a[gp][k]);
    }
  }
  __device__ __forceinline__ void step_resp(const float*, const float*) {}
  __device__ __forceinline__ float get(int g, int k) const { return (g & 1) ? a[g >> 1][k].y : a[g >> 1][k].x; }
};

template <int G, int NMOM>
struct Acc<double, G, NMOM> {
  static constexpr int STRIDE = (NMOM == 19) ? 25 : 33;  // doubles per row; odd => conflict-free 64-bit stores
  double bx[G], by[G], bz[G];
  double a[G][NMOM];
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
      for (int k = 0; k < NMOM; ++k) a[g][k] = 0.0;
  }
  __device__ __forceinline__ void load_bins(const double* s_bins, int Btot, int base, int L) {
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const int b = base + g * L;
      bx[g] = s_bins[b]; by[g] = s_bins[Btot + b]; bz[g] = s_bins[2 * Btot + b];
    }
  }
  __device__ __forceinline__ void step(const double* row) {
    const double d0 = row[0], d1 = row[1], d2 = row[2], ov = row[3];
    double e[G];
#pragma unroll
    for (int g = 0; g < G; ++g) e[g] = exp(fma(d0, bx[g], fma(d1, by[g], fma(d2, bz[g], ov))));
    step_resp(row, e);
  }
  __device__ __forceinline__ void step_resp(const double* row, const double* e) {
#pragma unroll
    for (int k = 0; k < NMOM; ++k) {
      const double f = row[4 + k];
#pragma unroll
      for (int g = 0; g < G; ++g) a[g][k] = fma(e[g], f, a[g][k]);
    }
  }
  __device__ __forceinline__ double get(int g, int k) const { return a[g][k]; }
};

// ------------------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------------------
template <typename T> struct FlushCfg;
template <> struct FlushCfg<float> { static constexpr int KC = 19; };   // all moments in one staging pass (float staging)
template <> struct FlushCfg<double> { static constexpr int KC = 5; };   // five moments per pass (double staging)

template <typename T, int G, int NMOM, bool FROM_RESP, int NT>
__global__ void __launch_bounds__(NT, 1) scan_moments_kernel(const __grid_constant__ ScanParams P) {
  using AccT = Acc<T, G, NMOM>;
  constexpr int STRIDE = AccT::STRIDE;
  constexpr int NV = 4 + NMOM;  // values per tile row: d(3), offset, moments
  constexpr int kKC = FlushCfg<T>::KC;
  static_assert(NV <= STRIDE, "tile row too short");

  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int B = P.n_bins;
  const int L = P.lanes;             // bin lanes per point slice (power of two)
  const int Bp = G * L;              // bins per chunk
  const int n_chunks = P.n_chunks;   // ceil(B / Bp)
  const int Btot = Bp * n_chunks;
  const int PS = NT / L;             // point slices
  // shared memory: [phase-B atlas SoA 3 x Btot][phase-A atlas, packed][point tile][flush staging]
  T* s_bins = reinterpret_cast<T*>(smem_raw);
  size_t off = ((size_t)3 * Btot * sizeof(T) + 15) & ~(size_t)15;
  T* s_binsA = reinterpret_cast<T*>(smem_raw + off);               // fp32: 8 floats per bin pair; fp64: 4 doubles per bin
  off += (sizeof(T) == 4) ? (size_t)((B + 1) / 2) * 8 * sizeof(float) : (size_t)B * 4 * sizeof(double);
  off = (off + 15) & ~(size_t)15;
  T* s_tile = reinterpret_cast<T*>(smem_raw + off);                 // [NT][STRIDE]
  off += (size_t)NT * STRIDE * sizeof(T);
  off = (off + 15) & ~(size_t)15;
  T* s_stage = reinterpret_cast<T*>(smem_raw + off);                // [PS][kKC][Bp] flush staging (not aliased: with
                                                                    // several bin chunks the tile outlives a flush)
  __shared__ double s_red[NT / 32][GCL_NCERT];
  __shared__ __align__(16) double s_in[NT * 5];  // raw rows of the next tile (cp.async): f64 {x,y,z,t,w} or f32 {xyzt | w}

  const int tid = threadIdx.x;
  const gcl_job& J = P.jobs ? P.jobs[blockIdx.y] : P.job;
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
  const int64_t rpc = (rows + gridDim.x - 1) / gridDim.x;
  const int64_t lo0 = sel_begin + (int64_t)blockIdx.x * rpc;
  const int64_t lo = lo0 < sel_end ? lo0 : sel_end;
  const int64_t hi = (lo + rpc < sel_end) ? lo + rpc : sel_end;
  double* part = P.partials + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * P.rec_doubles;

  // exponent scale folded into the atlas: fp32 works in log2 units (ex2.approx), fp64 in natural units
  const double cscale_d = (sizeof(T) == 4) ? (1.4426950408889634 / P.tau) : (1.0 / P.tau);
  const T cscale = (T)cscale_d;

  // --- zero this CTA's record; atlas to shared (pre-scaled; padded bins = 0) ---------------------------------------
  for (int i = tid; i < NMOM * B + GCL_NCERT; i += NT) part[i] = 0.0;
  if (!FROM_RESP) {
    for (int i = tid; i < 3 * Btot; i += NT) {
      const int c = i / Btot, b = i - c * Btot;
      s_bins[i] = (b < B) ? (T)(P.bin_dirs[b * 3 + c] * cscale_d) : (T)0;
    }
    if (sizeof(T) == 4) {
      const int np = (B + 1) / 2;
      for (int i = tid; i < np * 8; i += NT) {
        const int bp = i >> 3, j = i & 7;          // {bx0,bx1,by0,by1,bz0,bz1,0,0}
        const int b = 2 * bp + (j & 1), c = j >> 1;
        s_binsA[i] = (c < 3 && b < B) ? (T)(P.bin_dirs[b * 3 + c] * cscale_d) : (T)0;
      }
    } else {
      for (int i = tid; i < B * 4; i += NT) {
        const int b = i >> 2, c = i & 3;
        s_binsA[i] = (c < 3) ? (T)(P.bin_dirs[b * 3 + c] * cscale_d) : (T)0;
      }
    }
  }
  const int binlane = tid & (L - 1);
  const int slice = tid / L;
  float xi32[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) xi32[k] = (float)J.xi_body[k];
  __syncthreads();

  AccT acc;
  acc.zero();
  if (!FROM_RESP) acc.load_bins(s_bins, Btot, binlane, L);

  double c_wraw = 0, c_wsel = 0, c_wsel2 = 0, c_wdesk = 0, c_ent = 0, c_maxr = 0, c_rows = 0;

  // --- sum of ALL raw weights of this CTA's raw range (mass preservation, point_budget.py:69) ------------------
  if (!FROM_RESP && stride > 1 && hi > lo) {
    const int64_t rlo = lo * stride;
    const int64_t rhi = (hi >= n_sel_total) ? n_raw : hi * stride;
    if (P.layout == GCL_LAYOUT_REF_F64) {
      const double* w = static_cast<const double*>(J.weights);
      for (int64_t r = rlo + tid; r < rhi; r += NT) c_wraw += w[r];
    } else {
      const float* w = static_cast<const float*>(J.weights);
      for (int64_t r = rlo + tid; r < rhi; r += NT) c_wraw += (double)w[r];
    }
  }

  // fixed-order cross-slice reduction, kKC moments at a time, carried in fp64 in this CTA's record
  auto flush = [&](int chunk) {
#pragma unroll
    for (int k0 = 0; k0 < NMOM; k0 += kKC) {
#pragma unroll
      for (int kc = 0; kc < kKC; ++kc) {
        if (k0 + kc < NMOM) {
#pragma unroll
          for (int g = 0; g < G; ++g)
            s_stage[((size_t)slice * kKC + kc) * Bp + g * L + binlane] = acc.get(g, (k0 + kc < NMOM) ? k0 + kc : 0);
        }
      }
      __syncthreads();
      for (int e = tid; e < kKC * Bp; e += NT) {
        const int kc = e / Bp, bl = e - kc * Bp, k = k0 + kc, b = chunk * Bp + bl;
        if (k < NMOM && b < B) {
          double s = 0.0;
          for (int sl = 0; sl < PS; ++sl) s += (double)s_stage[((size_t)sl * kKC + kc) * Bp + bl];
          part[(size_t)k * B + b] += s;
        }
      }
      __syncthreads();
    }
    acc.zero();
  };

  // raw rows of tile [t0, t0 + NT) -> s_in, asynchronously (one row per thread)
  auto prefetch_tile = [&](int64_t t0) {
    const int64_t srow = t0 + tid;
    if (FROM_RESP || srow >= hi) return;
    const int64_t r = srow * stride;
    if (P.layout == GCL_LAYOUT_REF_F64) {
      const double* pts = static_cast<const double*>(J.points) + r * 3;
      double* dst = s_in + (size_t)tid * 5;
      cp_async8(dst + 0, pts + 0); cp_async8(dst + 1, pts + 1); cp_async8(dst + 2, pts + 2);
      cp_async8(dst + 3, static_cast<const double*>(J.timestamps) + r);
      cp_async8(dst + 4, static_cast<const double*>(J.weights) + r);
    } else {
      float* base = reinterpret_cast<float*>(s_in);
      cp_async16(base + (size_t)tid * 4, static_cast<const float4*>(J.points) + r);
      cp_async4(base + (size_t)NT * 4 + tid, static_cast<const float*>(J.weights) + r);
    }
  };
  prefetch_tile(lo);

  int tiles_since_flush = 0;
  // fp32: flush once a thread's chains have seen ~512 points (NT / PS = L points per tile per thread)
  const int flush_every = (n_chunks > 1) ? 1 : ((sizeof(T) == 4) ? ((512 + L - 1) / L) : (1 << 30));

  for (int64_t tile0 = lo; tile0 < hi; tile0 += NT) {
    const int count = (int)((hi - tile0 < NT) ? hi - tile0 : NT);
    // ======================================= phase A ==========================================================
    if (!FROM_RESP) { cp_async_wait_all(); __syncthreads(); }
    {
      const int i = tid;  // one point per thread
      T* row = s_tile + (size_t)i * STRIDE;
      T v[NV + 1];
#pragma unroll
      for (int k = 0; k < NV + 1; ++k) v[k] = (T)0;
      if (i < count) {
        const int64_t srow = tile0 + i;   // selected-row index
        const int64_t r = srow * stride;  // raw index (PointBudget gather)
        T d0, d1, d2, q0, q1, q2;         // direction, deskewed point
        double w_sel, w_time;
        if (FROM_RESP) {
          // standalone ScanBinMomentMatch: points already deskewed, weights final (binning.py:157-164)
          const double* pts = static_cast<const double*>(J.points);
          const double px = pts[r * 3 + 0], py = pts[r * 3 + 1], pz = pts[r * 3 + 2];
          w_sel = static_cast<const double*>(J.weights)[r];
          const double r0 = px - C.origin[0], r1 = py - C.origin[1], r2 = pz - C.origin[2];
          const double inv = 1.0 / (sqrt(r0 * r0 + r1 * r1 + r2 * r2) + C.eps_mass);
          d0 = (T)(r0 * inv); d1 = (T)(r1 * inv); d2 = (T)(r2 * inv);
          q0 = (T)px; q1 = (T)py; q2 = (T)pz;
          w_time = 1.0;
        } else if (P.layout == GCL_LAYOUT_REF_F64) {
          const double* in = s_in + (size_t)i * 5;  // {x, y, z, t, w} staged by cp.async
          const double px = in[0], py = in[1], pz = in[2], tt = in[3];
          w_sel = in[4];
          if (sizeof(T) == 8) {
            PointOut64 o;
            point_front64(J, C, px, py, pz, tt, o);
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
          const float* base = reinterpret_cast<const float*>(s_in);
          const float4 xyzt = *reinterpret_cast<const float4*>(base + (size_t)i * 4);
          w_sel = (double)base[(size_t)NT * 4 + i];
          PointOut32 o;
          point_front32(J, C, xi32, xyzt.x, xyzt.y, xyzt.z, xyzt.w, o);
          d0 = (T)o.d[0]; d1 = (T)o.d[1]; d2 = (T)o.d[2];
          q0 = (T)o.p0[0]; q1 = (T)o.p0[1]; q2 = (T)o.p0[2];
          w_time = (double)o.w_time;
        }
        const double lam = J.point_lambda ? J.point_lambda[srow] : 1.0;
        const double w_desk = w_sel * w_time;
        c_wsel += w_sel; c_wsel2 += w_sel * w_sel; c_wdesk += w_desk; c_rows += 1.0;
        T wq = (T)(w_desk * lam);
        T offv = (T)0;
        if (!FROM_RESP) {
          // Shift = the largest possible logit (|d.b| <= 1) when that cannot underflow, else the true row maximum
          // (jax.nn.softmax, binning.py:69) from an extra pass.
          offv = (P.fixed_shift & 1) ? -cscale : -sim_rowmax(s_binsA, B, d0, d1, d2);
          T Z = (T)1, S1 = (T)0, xmax = (T)0;
          if (!(P.fixed_shift & 4)) softmax_stats(s_binsA, B, d0, d1, d2, offv, Z, S1, xmax);
          // H_i = -sum_b r log(r + eps) = log Z - S1/Z - B*eps + O(B eps^2 / r)   (binning.py:70-72)
          if (sizeof(T) == 4) {
            const float Zf = (float)Z, invZ = 1.f / Zf;
            c_ent += (double)(__logf(Zf) - 0.6931471805599453f * (float)S1 * invZ) - (double)B * C.eps_mass;
            const double mr = (double)(fast_exp2((float)xmax) * invZ);
            c_maxr = mr > c_maxr ? mr : c_maxr;
          } else {
            const double Zd = (double)Z;
            c_ent += log(Zd) - (double)S1 / Zd - (double)B * C.eps_mass;
            const double mr = exp((double)xmax) / Zd;
            c_maxr = mr > c_maxr ? mr : c_maxr;
          }
          wq = wq / Z;
        }
        v[0] = d0; v[1] = d1; v[2] = d2; v[3] = offv;
        T* f = v + 4;
        f[0] = wq;
        f[1] = wq * d0; f[2] = wq * d1; f[3] = wq * d2;
        f[4] = f[1] * d0; f[5] = f[1] * d1; f[6] = f[1] * d2; f[7] = f[2] * d1; f[8] = f[2] * d2; f[9] = f[3] * d2;
        f[10] = wq * q0; f[11] = wq * q1; f[12] = wq * q2;
        f[13] = f[10] * q0; f[14] = f[10] * q1; f[15] = f[10] * q2; f[16] = f[11] * q1; f[17] = f[11] * q2; f[18] = f[12] * q2;
        if (NMOM > 19) {  // per-point covariances, all nine entries as the einsum sums them (binning.py:173)
#pragma unroll
          for (int k = 0; k < 9; ++k) f[(19 + k < NMOM) ? 19 + k : 0] = P.point_cov ? wq * (T)P.point_cov[r * 9 + k] : (T)0;
        }
      }
      if (sizeof(T) == 4) {
        float4* r4 = reinterpret_cast<float4*>(row);
#pragma unroll
        for (int k = 0; k < 6; ++k) r4[k] = make_float4((float)v[4 * k], (float)v[4 * k + 1], (float)v[4 * k + 2], (float)v[4 * k + 3]);
      } else {
#pragma unroll
        for (int k = 0; k < NV; ++k) row[k] = v[k];
      }
    }
    __syncthreads();
    prefetch_tile(tile0 + NT);  // next tile's raw rows stream in while this tile accumulates
    // ======================================= phase B ==========================================================
    for (int chunk = 0; chunk < n_chunks; ++chunk) {
      if (n_chunks > 1 && !FROM_RESP) acc.load_bins(s_bins, Btot, chunk * Bp + binlane, L);
      const int npts = (P.fixed_shift & 2) ? 0 : count;  // (profiling toggle GCL_DEBUG_SKIP)
      for (int pt = slice; pt < npts; pt += PS) {
        const T* row = s_tile + (size_t)pt * STRIDE;
        if (!FROM_RESP) {
          acc.step(row);
        } else {
          const int64_t prow = tile0 + pt;
          T e[G];
#pragma unroll
          for (int g = 0; g < G; ++g) {
            const int b = chunk * Bp + g * L + binlane;
            e[g] = (b < B) ? (T)P.resp[prow * B + b] : (T)0;
          }
          acc.step_resp(row, e);
        }
      }
      if (n_chunks > 1) {
        __syncthreads();
        flush(chunk);
      }
    }
    if (n_chunks == 1) {
      __syncthreads();
      if (++tiles_since_flush >= flush_every) { flush(0); tiles_since_flush = 0; }
    }
  }
  if (n_chunks == 1 && tiles_since_flush > 0) flush(0);

  // --- certificate scalars of this CTA -------------------------------------------------------------------------
  double cs[GCL_NCERT] = {c_wraw, c_wsel, c_wsel2, c_wdesk, c_ent, c_maxr, c_rows, 0.0};
#pragma unroll
  for (int k = 0; k < GCL_NCERT; ++k) {
    double v = cs[k];
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
    for (int w = 1; w < NT / 32; ++w) v = (src == GCL_CS_MAX_RESP) ? fmax(v, s_red[w][src]) : v + s_red[w][src];
    part[(size_t)NMOM * B + tid] = v;
  }
}

// ------------------------------------------------------------------------------------------------------------
// launch
// ------------------------------------------------------------------------------------------------------------
static inline int pow2_ceil(int x) { int p = 1; while (p < x) p <<= 1; return p; }

int scan_moments_threads(int precision) { return (precision == GCL_F64) ? 256 : GCL_NT_F32; }

// Bins are laid out as n_chunks chunks of Bp = G*L (L a power of two in [8,256] lanes, G bins per thread).
// One chunk (accumulators stay in registers across tiles) whenever Gmax*256 >= B; fp64 keeps G <= 2 (registers).
void plan_bins(int n_bins, int precision, int* G_out, int* L_out, int* chunks_out) {
  const int Gmax = (precision == GCL_F64) ? 2 : GCL_GMAX_F32;
  const int Gstep = (precision == GCL_F64) ? 1 : 2;  // the fp32 kernel handles bins in pairs
  int bestG = Gstep, bestL = 256, bestChunks = 1 << 20, bestWaste = 1 << 30;
  for (int G = Gmax; G >= Gstep; G -= Gstep) {
    int L = pow2_ceil((n_bins + G - 1) / G);
    if (L < 8) L = 8;
    if (L > 256) L = 256;
    const int Bp = G * L;
    const int chunks = (n_bins + Bp - 1) / Bp;
    const int waste = chunks * Bp - n_bins;
    if (chunks < bestChunks || (chunks == bestChunks && waste < bestWaste)) {
      bestChunks = chunks; bestWaste = waste; bestG = G; bestL = L;
    }
  }
  *G_out = bestG; *L_out = bestL; *chunks_out = bestChunks;
}

template <typename T, int G, int NMOM, bool FROM_RESP, int NT>
static cudaError_t launch_one(const ScanParams& P, dim3 grid, size_t smem, cudaStream_t stream) {
  auto kern = scan_moments_kernel<T, G, NMOM, FROM_RESP, NT>;
  static bool attr_done = false;  // per instantiation
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  kern<<<grid, NT, smem, stream>>>(P);
  return cudaGetLastError();
}

size_t scan_moments_smem_bytes(int precision, int G, int L, int n_chunks, int nmom) {
  const size_t tsz = precision == GCL_F64 ? 8 : 4;
  const int nt = scan_moments_threads(precision);
  const int stride = precision == GCL_F64 ? (nmom == 19 ? 25 : 33) : 28;
  size_t off = ((size_t)3 * G * L * n_chunks * tsz + 15) & ~(size_t)15;
  off += (size_t)GCL_MAX_BINS * 4 * 8;  // phase-A atlas copy (upper bound)
  size_t tile = ((size_t)nt * stride * tsz + 15) & ~(size_t)15;
  const size_t stage = (size_t)(nt / L) * (precision == GCL_F64 ? 5 * sizeof(double) : 19 * sizeof(float)) * G * L;
  return off + tile + stage + 64;
}

cudaError_t launch_scan_moments(const ScanParams& P, int n_jobs, int ctas_per_job, int precision, int G, int nmom,
                                bool from_resp, cudaStream_t stream) {
  dim3 grid(ctas_per_job, n_jobs);
  const size_t smem = scan_moments_smem_bytes(from_resp ? GCL_F64 : precision, G, P.lanes, P.n_chunks, nmom);
  if (from_resp) {
    if (G == 1) return launch_one<double, 1, 28, true, 256>(P, grid, smem, stream);
    return launch_one<double, 2, 28, true, 256>(P, grid, smem, stream);
  }
  (void)precision;
  return cudaErrorNotSupported;  // the fused path lives in gcl_scan_fused.cu
}

}  // namespace gcl
