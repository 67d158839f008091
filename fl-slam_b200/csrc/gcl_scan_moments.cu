// gcl_scan_moments.cu -- K1: the one kernel that touches O(N) data.
//
// One pass over the raw points fuses: PointBudget gather (stride), constant-twist deskew + time-window weights,
// ray directions, softmax responsibilities over the bin atlas and the per-bin weighted moment sums
//   M_b[19] = sum_i w_i r_ib [1, d, ddT(6), p, ppT(6)]
// without materialising the (N,B) responsibilities or the (N,3,3) outer products the reference builds
// (archive/legacy_operators/binning.py:69,:171).
//
// Structure per CTA (256 threads, persistent over its slice of the selected rows), tile = 256 points:
//   phase A  thread-per-point: deskew, window weight, direction, softmax normaliser Z_i over all bins, entropy and
//            max-responsibility certificate partials; writes d_i, the exponent offset and w_i/Z_i * phi_i to shared.
//   phase B  "GEMM by hand" on CUDA cores: thread = (bin lane, point slice) owns G bins x 19 accumulators in
//            registers; per step it reads one shared-memory row (warp-broadcast LDS.128), recomputes
//            e_ib = exp2(c d_i.b_b + off_i) and does 19 fused multiply-adds per bin.  In fp32 mode a step is TWO
//            points and every accumulator is a float2, so the adds issue as packed fma.rn.f32x2 (FFMA2, sm_100).
//   flush    per-thread fp32 partials are combined across point slices in a FIXED order and carried in fp64 in this
//            CTA's record of the partials buffer (no atomics => bit-reproducible run to run).
// The contraction depth of the similarity is 3, so tensor cores are not used (BASELINE.json north_star).
#include "gcl_internal.cuh"

namespace gcl {

// ------------------------------------------------------------------------------------------------------------
// packed arithmetic
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
#if GCL_USE_FFMA2
  float2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;"
      : "=l"(reinterpret_cast<unsigned long long&>(d))
      : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)),
        "l"(reinterpret_cast<unsigned long long&>(c)));
  return d;
#else
  return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#endif
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
#if GCL_USE_FFMA2
  float2 d;
  asm("mul.rn.f32x2 %0, %1, %2;"
      : "=l"(reinterpret_cast<unsigned long long&>(d))
      : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)));
  return d;
#else
  return make_float2(a.x * b.x, a.y * b.y);
#endif
}
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <typename T> struct Vec;
template <> struct Vec<float> {
  using V = float2;
  static constexpr int W = 2;        // points per step
  static constexpr int STRIDE = 26;  // V per tile row (24 used): 208 B keeps LDS.128 aligned, 2-way STS conflicts
  static __device__ __forceinline__ V splat(float x) { return make_float2(x, x); }
  static __device__ __forceinline__ V zero() { return make_float2(0.f, 0.f); }
  static __device__ __forceinline__ V fma(V a, V b, V c) { return ffma2(a, b, c); }
  static __device__ __forceinline__ V mul(V a, V b) { return fmul2(a, b); }
  static __device__ __forceinline__ V expo(V s) { return make_float2(fast_exp2(s.x), fast_exp2(s.y)); }
  static __device__ __forceinline__ double hsum(V a) { return (double)a.x + (double)a.y; }
};
template <> struct Vec<double> {
  using V = double;
  static constexpr int W = 1;
  static constexpr int STRIDE = 25;  // 200 B: conflict-free 8-byte stores in phase A
  static __device__ __forceinline__ V splat(double x) { return x; }
  static __device__ __forceinline__ V zero() { return 0.0; }
  static __device__ __forceinline__ V fma(V a, V b, V c) { return ::fma(a, b, c); }
  static __device__ __forceinline__ V mul(V a, V b) { return a * b; }
  static __device__ __forceinline__ V expo(V s) { return exp(s); }
  static __device__ __forceinline__ double hsum(V a) { return a; }
};

template <typename T> __device__ __forceinline__ void set_lane(typename Vec<T>::V& v, int lane, T x);
template <> __device__ __forceinline__ void set_lane<float>(float2& v, int lane, float x) { (lane ? v.y : v.x) = x; }
template <> __device__ __forceinline__ void set_lane<double>(double& v, int, double x) { v = x; }

// ------------------------------------------------------------------------------------------------------------
// per-point front end (phase A, before the bin loop)
// ------------------------------------------------------------------------------------------------------------
struct PointOut64 { double p0[3], d[3], w_sel, w_time; };

// fp64: the reference formulas verbatim (gcl_math.cuh cites the lines)
__device__ __forceinline__ void point_front64(const gcl_job& J, const ScanConst& C, double px, double py, double pz,
                                              double t, PointOut64& o) {
  const double denom = fmax(J.scan_end_time - J.scan_start_time, 1e-12);
  const double alpha = (t - J.scan_start_time) / denom;
  const double p[3] = {px, py, pz};
  deskew_point64(J.xi_body, alpha, p, o.p0);
  const double sigma = fmax(0.1 * denom, 1e-6);  // GC_TIME_WARP_SIGMA_FRAC, constants.py:143; floor imu_preintegration.py:32
  o.w_time = window_weight64(t, J.scan_start_time, J.scan_end_time, sigma, 1e-12);
  const double r0 = o.p0[0] - C.origin[0], r1 = o.p0[1] - C.origin[1], r2 = o.p0[2] - C.origin[2];
  const double inv = 1.0 / (sqrt(r0 * r0 + r1 * r1 + r2 * r2) + C.eps_mass);
  o.d[0] = r0 * inv; o.d[1] = r1 * inv; o.d[2] = r2 * inv;
}

struct PointOut32 { float p0[3], d[3], w_time; };

// fp32: same maps; series for the Rodrigues coefficients below 0.5 rad (no cancellation), sincosf above
__device__ __forceinline__ void point_front32(const gcl_job& J, const ScanConst& C, const float* xi32, float px, float py,
                                              float pz, float trel /* t - scan_start */, PointOut32& o) {
  const float D = (float)(J.scan_end_time - J.scan_start_time);
  const float denom = fmaxf(D, 1e-12f);
  const float alpha = trel / denom;
  const float r0 = alpha * xi32[0], r1 = alpha * xi32[1], r2 = alpha * xi32[2];
  const float f0 = alpha * xi32[3], f1 = alpha * xi32[4], f2 = alpha * xi32[5];
  const float th2 = f0 * f0 + f1 * f1 + f2 * f2;
  float sinc, Bc, Cc;
  if (th2 < 0.25f) {
    sinc = 1.f + th2 * (-1.f / 6.f + th2 * (1.f / 120.f + th2 * (-1.f / 5040.f + th2 * (1.f / 362880.f))));
    Bc = 0.5f + th2 * (-1.f / 24.f + th2 * (1.f / 720.f + th2 * (-1.f / 40320.f + th2 * (1.f / 3628800.f))));
    Cc = 1.f / 6.f + th2 * (-1.f / 120.f + th2 * (1.f / 5040.f + th2 * (-1.f / 362880.f + th2 * (1.f / 39916800.f))));
  } else {
    const float th = sqrtf(th2);
    float sn, cs, sh, ch;
    sincosf(th, &sn, &cs);
    sincosf(0.5f * th, &sh, &ch);
    sinc = sn / th;
    Bc = 2.f * sh * sh / th2;
    Cc = (th - sn) / (th2 * th);
  }
  const float c0 = f1 * r2 - f2 * r1, c1 = f2 * r0 - f0 * r2, c2 = f0 * r1 - f1 * r0;
  const float d0 = f1 * c2 - f2 * c1, d1 = f2 * c0 - f0 * c2, d2 = f0 * c1 - f1 * c0;
  const float u0 = px - (r0 + Bc * c0 + Cc * d0), u1 = py - (r1 + Bc * c1 + Cc * d1), u2 = pz - (r2 + Bc * c2 + Cc * d2);
  const float g0 = f1 * u2 - f2 * u1, g1 = f2 * u0 - f0 * u2, g2 = f0 * u1 - f1 * u0;
  const float h0 = f1 * g2 - f2 * g1, h1 = f2 * g0 - f0 * g2, h2 = f0 * g1 - f1 * g0;
  o.p0[0] = u0 - sinc * g0 + Bc * h0;
  o.p0[1] = u1 - sinc * g1 + Bc * h1;
  o.p0[2] = u2 - sinc * g2 + Bc * h2;
  const float sigma = fmaxf(0.1f * denom, 1e-6f);
  const float a = trel / sigma, b = (D - trel) / sigma;
  const float sa = 1.f / (1.f + __expf(-a)), sb = 1.f / (1.f + __expf(-b));
  o.w_time = sa * sb * (1.f - 1e-12f) + 1e-12f;
  const float q0 = o.p0[0] - (float)C.origin[0], q1 = o.p0[1] - (float)C.origin[1], q2 = o.p0[2] - (float)C.origin[2];
  const float inv = 1.f / (sqrtf(q0 * q0 + q1 * q1 + q2 * q2) + (float)C.eps_mass);
  o.d[0] = q0 * inv; o.d[1] = q1 * inv; o.d[2] = q2 * inv;
}

// ------------------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------------------
constexpr int kThreads = 256;
constexpr int kTilePoints = 256;
constexpr int kKC = 5;  // moments per flush chunk

// two consecutive tile values with one shared-memory load where alignment allows it
__device__ __forceinline__ void load2(const float2* row, int k, float2& a, float2& b) {  // k even, row 16 B aligned
  const float4 v = *reinterpret_cast<const float4*>(row + k);
  a = make_float2(v.x, v.y);
  b = make_float2(v.z, v.w);
}
__device__ __forceinline__ void load2(const double* row, int k, double& a, double& b) { a = row[k]; b = row[k + 1]; }

template <typename T, int G, int NMOM, bool FROM_RESP>
__global__ void __launch_bounds__(kThreads, 1) scan_moments_kernel(const __grid_constant__ ScanParams P) {
  using VT = Vec<T>;
  using V = typename VT::V;
  constexpr int W = VT::W;
  constexpr int NV = 4 + NMOM;                                     // values per tile row: d(3), offset, moments
  constexpr int STRIDE = (NMOM == 19) ? VT::STRIDE : 33;           // 28-moment rows exist in fp64 only
  static_assert(NV <= STRIDE, "tile row too short");

  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int B = P.n_bins;
  const int L = P.lanes;             // bin lanes per point slice (power of two)
  const int Bp = G * L;              // bins per chunk
  const int n_chunks = P.n_chunks;   // ceil(B / Bp)
  const int Btot = Bp * n_chunks;
  const int PS = kThreads / L;
  T* s_bins = reinterpret_cast<T*>(smem_raw);                       // [3][Btot]
  size_t off = (size_t)3 * Btot * sizeof(T);
  off = (off + 15) & ~(size_t)15;
  V* s_tile = reinterpret_cast<V*>(smem_raw + off);                 // [kTilePoints / W][STRIDE]
  off += (size_t)(kTilePoints / W) * STRIDE * sizeof(V);
  off = (off + 15) & ~(size_t)15;
  double* s_stage = reinterpret_cast<double*>(smem_raw + off);      // [PS][kKC][Bp] flush staging (not aliased: with
                                                                    // several bin chunks the tile outlives a flush)
  __shared__ double s_red[kThreads / 32][GCL_NCERT];

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
  int64_t rpc = (rows + gridDim.x - 1) / gridDim.x;
  rpc = (rpc + W - 1) / W * W;
  const int64_t lo0 = sel_begin + (int64_t)blockIdx.x * rpc;
  const int64_t lo = lo0 < sel_end ? lo0 : sel_end;
  const int64_t hi = (lo + rpc < sel_end) ? lo + rpc : sel_end;
  double* part = P.partials + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * P.rec_doubles;

  // --- zero this CTA's record; atlas to shared (SoA, padded bins = 0) -------------------------------------------
  for (int i = tid; i < NMOM * B + GCL_NCERT; i += kThreads) part[i] = 0.0;
  for (int i = tid; i < 3 * Btot; i += kThreads) {
    const int c = i / Btot, b = i - c * Btot;
    s_bins[i] = (!FROM_RESP && b < B) ? (T)P.bin_dirs[b * 3 + c] : (T)0;  // no atlas on the materialised-r path
  }
  const int binlane = tid & (L - 1);
  const int slice = tid / L;
  float xi32[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) xi32[k] = (float)J.xi_body[k];
  __syncthreads();

  V bx[G], by[G], bz[G];
  auto load_bins = [&](int chunk) {
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const int b = chunk * Bp + g * L + binlane;
      bx[g] = VT::splat(s_bins[0 * Btot + b]);
      by[g] = VT::splat(s_bins[1 * Btot + b]);
      bz[g] = VT::splat(s_bins[2 * Btot + b]);
    }
  };
  load_bins(0);
  V acc[G][NMOM];
#pragma unroll
  for (int g = 0; g < G; ++g)
#pragma unroll
    for (int k = 0; k < NMOM; ++k) acc[g][k] = VT::zero();

  // exponent scale: fp32 works in log2 units (ex2.approx), fp64 in natural units
  const T cscale = (sizeof(T) == 4) ? (T)(1.4426950408889634 / P.tau) : (T)(1.0 / P.tau);
  const V cscale_v = VT::splat(cscale);

  double c_wraw = 0, c_wsel = 0, c_wsel2 = 0, c_wdesk = 0, c_ent = 0, c_maxr = 0, c_rows = 0;

  // --- sum of ALL raw weights of this CTA's raw range (mass preservation, point_budget.py:69) ------------------
  if (!FROM_RESP && stride > 1 && hi > lo) {
    const int64_t rlo = lo * stride;
    const int64_t rhi = (hi >= n_sel_total) ? n_raw : hi * stride;
    if (P.layout == GCL_LAYOUT_REF_F64) {
      const double* w = static_cast<const double*>(J.weights);
      for (int64_t r = rlo + tid; r < rhi; r += kThreads) c_wraw += w[r];
    } else {
      const float* w = static_cast<const float*>(J.weights);
      for (int64_t r = rlo + tid; r < rhi; r += kThreads) c_wraw += (double)w[r];
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
            s_stage[((size_t)slice * kKC + kc) * Bp + g * L + binlane] = VT::hsum(acc[g][(k0 + kc < NMOM) ? k0 + kc : 0]);
        }
      }
      __syncthreads();
      for (int e = tid; e < kKC * Bp; e += kThreads) {
        const int kc = e / Bp, bl = e - kc * Bp, k = k0 + kc, b = chunk * Bp + bl;
        if (k < NMOM && b < B) {
          double s = 0.0;
          for (int sl = 0; sl < PS; ++sl) s += s_stage[((size_t)sl * kKC + kc) * Bp + bl];
          part[(size_t)k * B + b] += s;
        }
      }
      __syncthreads();
    }
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
      for (int k = 0; k < NMOM; ++k) acc[g][k] = VT::zero();
  };

  int tiles_since_flush = 0;
  const int flush_every = (n_chunks > 1) ? 1 : ((sizeof(T) == 4) ? (PS > 1 ? PS : 1) : (1 << 30));

  for (int64_t tile0 = lo; tile0 < hi; tile0 += kTilePoints) {
    const int count = (int)((hi - tile0 < kTilePoints) ? hi - tile0 : kTilePoints);
    // ======================================= phase A ==========================================================
    {
      const int i = tid;  // one point per thread
      V* row = s_tile + (size_t)(i / W) * STRIDE;
      T vals[NV];
#pragma unroll
      for (int k = 0; k < NV; ++k) vals[k] = (T)0;
      if (i < count) {
        const int64_t srow = tile0 + i;   // selected-row index
        const int64_t r = srow * stride;  // raw index (PointBudget gather)
        T d0, d1, d2, q0, q1, q2;         // direction, deskewed point
        double w_sel, w_time;
        if (sizeof(T) == 8 || P.layout == GCL_LAYOUT_REF_F64) {
          const double* pts = static_cast<const double*>(J.points);
          const double px = pts[r * 3 + 0], py = pts[r * 3 + 1], pz = pts[r * 3 + 2];
          w_sel = static_cast<const double*>(J.weights)[r];
          if (FROM_RESP) {
            // standalone ScanBinMomentMatch: points already deskewed, weights final (binning.py:157-164)
            const double r0 = px - C.origin[0], r1 = py - C.origin[1], r2 = pz - C.origin[2];
            const double inv = 1.0 / (sqrt(r0 * r0 + r1 * r1 + r2 * r2) + C.eps_mass);
            d0 = (T)(r0 * inv); d1 = (T)(r1 * inv); d2 = (T)(r2 * inv);
            q0 = (T)px; q1 = (T)py; q2 = (T)pz;
            w_time = 1.0;
          } else if (sizeof(T) == 8) {
            PointOut64 o;
            point_front64(J, C, px, py, pz, static_cast<const double*>(J.timestamps)[r], o);
            d0 = (T)o.d[0]; d1 = (T)o.d[1]; d2 = (T)o.d[2];
            q0 = (T)o.p0[0]; q1 = (T)o.p0[1]; q2 = (T)o.p0[2];
            w_time = o.w_time;
          } else {
            PointOut32 o;
            const float trel = (float)(static_cast<const double*>(J.timestamps)[r] - J.scan_start_time);
            point_front32(J, C, xi32, (float)px, (float)py, (float)pz, trel, o);
            d0 = (T)o.d[0]; d1 = (T)o.d[1]; d2 = (T)o.d[2];
            q0 = (T)o.p0[0]; q1 = (T)o.p0[1]; q2 = (T)o.p0[2];
            w_time = (double)o.w_time;
          }
        } else {
          const float4 xyzt = static_cast<const float4*>(J.points)[r];
          w_sel = (double)static_cast<const float*>(J.weights)[r];
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
          // softmax normaliser over all bins.  Shift = the largest possible logit (|d.b| <= 1) when that cannot
          // underflow, else the true row maximum (jax.nn.softmax, binning.py:69) from an extra pass.
          T m = (T)1;
          if (!P.fixed_shift) {
            m = (T)-2;
            for (int b = 0; b < B; ++b) {
              const T sdot = d0 * s_bins[b] + d1 * s_bins[Btot + b] + d2 * s_bins[2 * Btot + b];
              m = sdot > m ? sdot : m;
            }
          }
          offv = -m * cscale;
          T Z = (T)0, S1 = (T)0, emax = (T)0;
          for (int b = 0; b < B; ++b) {
            const T sdot = d0 * s_bins[b] + d1 * s_bins[Btot + b] + d2 * s_bins[2 * Btot + b];
            const T x = sdot * cscale + offv;  // <= 0 up to rounding
            T e;
            if (sizeof(T) == 4) e = (T)fast_exp2((float)x); else e = (T)exp((double)x);
            Z += e;
            S1 += e * x;
            emax = e > emax ? e : emax;
          }
          const double Zd = (double)Z;
          const double s1_nat = (sizeof(T) == 4) ? 0.6931471805599453 * (double)S1 : (double)S1;
          // H_i = -sum_b r log(r + eps) = log Z - S1/Z - B*eps + O(B eps^2 / r)   (binning.py:70-72)
          c_ent += log(Zd) - s1_nat / Zd - (double)B * C.eps_mass;
          const double mr = (double)emax / Zd;
          c_maxr = mr > c_maxr ? mr : c_maxr;
          wq = wq / Z;
        }
        vals[0] = d0; vals[1] = d1; vals[2] = d2; vals[3] = offv;
        T* f = vals + 4;
        f[0] = wq;
        f[1] = wq * d0; f[2] = wq * d1; f[3] = wq * d2;
        f[4] = f[1] * d0; f[5] = f[1] * d1; f[6] = f[1] * d2; f[7] = f[2] * d1; f[8] = f[2] * d2; f[9] = f[3] * d2;
        f[10] = wq * q0; f[11] = wq * q1; f[12] = wq * q2;
        f[13] = f[10] * q0; f[14] = f[10] * q1; f[15] = f[10] * q2; f[16] = f[11] * q1; f[17] = f[11] * q2; f[18] = f[12] * q2;
        if (NMOM > 19) {  // per-point covariances, all nine entries as the einsum sums them (binning.py:173)
          if (P.point_cov) {
            const double* cv = P.point_cov + r * 9;
#pragma unroll
            for (int k = 0; k < 9; ++k) f[(19 + k < NMOM) ? 19 + k : 0] = wq * (T)cv[k];
          }
        }
      }
#pragma unroll
      for (int k = 0; k < NV; ++k) set_lane<T>(row[k], i % W, vals[k]);
    }
    __syncthreads();
    // ======================================= phase B ==========================================================
    const int nsteps = (count + W - 1) / W;
    for (int chunk = 0; chunk < n_chunks; ++chunk) {
      if (n_chunks > 1) load_bins(chunk);
      for (int step = slice; step < nsteps; step += PS) {
        const V* row = s_tile + (size_t)step * STRIDE;
        V e[G];
        if (!FROM_RESP) {
          V dx, dy, dz, ov;
          load2(row, 0, dx, dy);
          load2(row, 2, dz, ov);
#pragma unroll
          for (int g = 0; g < G; ++g) {
            V s = VT::mul(dx, bx[g]);
            s = VT::fma(dy, by[g], s);
            s = VT::fma(dz, bz[g], s);
            s = VT::fma(s, cscale_v, ov);
            e[g] = VT::expo(s);
          }
        } else {
          const int64_t prow = tile0 + step;  // W == 1 on this path
#pragma unroll
          for (int g = 0; g < G; ++g) {
            const int b = chunk * Bp + g * L + binlane;
            set_lane<T>(e[g], 0, (b < B) ? (T)P.resp[prow * B + b] : (T)0);
          }
        }
#pragma unroll
        for (int k = 0; k + 1 < NMOM; k += 2) {
          V f0, f1;
          load2(row, 4 + k, f0, f1);
#pragma unroll
          for (int g = 0; g < G; ++g) {
            acc[g][k] = VT::fma(e[g], f0, acc[g][k]);
            acc[g][k + 1] = VT::fma(e[g], f1, acc[g][k + 1]);
          }
        }
        if (NMOM & 1) {
          const V f = row[4 + NMOM - 1];
#pragma unroll
          for (int g = 0; g < G; ++g) acc[g][NMOM - 1] = VT::fma(e[g], f, acc[g][NMOM - 1]);
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
    for (int w = 1; w < kThreads / 32; ++w) v = (src == GCL_CS_MAX_RESP) ? fmax(v, s_red[w][src]) : v + s_red[w][src];
    part[(size_t)NMOM * B + tid] = v;
  }
}

// ------------------------------------------------------------------------------------------------------------
// launch
// ------------------------------------------------------------------------------------------------------------
static inline int pow2_ceil(int x) { int p = 1; while (p < x) p <<= 1; return p; }

// Bins are laid out as n_chunks chunks of Bp = G*L (L a power of two in [8,256] lanes, G bins per thread).
// One chunk (accumulators stay in registers across tiles) whenever G*256 >= B; fp64 keeps G <= 2 (registers).
void plan_bins(int n_bins, int precision, int* G_out, int* L_out, int* chunks_out) {
  const int Gmax = (precision == GCL_F64) ? 2 : 4;
  int bestG = 1, bestL = 256, bestChunks = 1 << 20, bestWaste = 1 << 30;
  for (int G = Gmax; G >= 1; --G) {
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

template <typename T, int G, int NMOM, bool FROM_RESP>
static cudaError_t launch_one(const ScanParams& P, dim3 grid, size_t smem, cudaStream_t stream) {
  auto kern = scan_moments_kernel<T, G, NMOM, FROM_RESP>;
  static bool attr_done = false;  // per instantiation
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  kern<<<grid, kThreads, smem, stream>>>(P);
  return cudaGetLastError();
}

size_t scan_moments_smem_bytes(int precision, int G, int L, int n_chunks, int nmom) {
  const size_t tsz = precision == GCL_F64 ? 8 : 4;
  const int W = precision == GCL_F64 ? 1 : 2;
  const int stride = (nmom == 19) ? (precision == GCL_F64 ? 25 : 26) : 33;
  size_t off = (size_t)3 * G * L * n_chunks * tsz;
  off = (off + 15) & ~(size_t)15;
  size_t tile = (size_t)(kTilePoints / W) * stride * tsz * W;
  tile = (tile + 15) & ~(size_t)15;
  const size_t stage = (size_t)(kThreads / L) * kKC * G * L * sizeof(double);
  return off + tile + stage;
}

cudaError_t launch_scan_moments(const ScanParams& P, int n_jobs, int ctas_per_job, int precision, int G, int nmom,
                                bool from_resp, cudaStream_t stream) {
  dim3 grid(ctas_per_job, n_jobs);
  const size_t smem = scan_moments_smem_bytes(from_resp ? GCL_F64 : precision, G, P.lanes, P.n_chunks, nmom);
  if (from_resp) {
    if (G == 1) return launch_one<double, 1, 28, true>(P, grid, smem, stream);
    return launch_one<double, 2, 28, true>(P, grid, smem, stream);
  }
  if (precision == GCL_F64) {
    if (G == 1) return launch_one<double, 1, 19, false>(P, grid, smem, stream);
    return launch_one<double, 2, 19, false>(P, grid, smem, stream);
  }
  switch (G) {
    case 1: return launch_one<float, 1, 19, false>(P, grid, smem, stream);
    case 2: return launch_one<float, 2, 19, false>(P, grid, smem, stream);
    case 3: return launch_one<float, 3, 19, false>(P, grid, smem, stream);
    default: return launch_one<float, 4, 19, false>(P, grid, smem, stream);
  }
}

}  // namespace gcl
