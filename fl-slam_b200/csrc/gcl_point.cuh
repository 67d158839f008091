// gcl_point.cuh -- per-point device helpers shared by the O(N) kernels: packed fp32 arithmetic, cp.async staging and the
// per-point front end (constant-twist deskew, time-window weight, ray direction).
#pragma once
#include "gcl_internal.cuh"

namespace gcl {

// ------------------------------------------------------------------------------------------------------------
// packed arithmetic (sm_100: FFMA2 / FADD2)
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
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
#if GCL_USE_FFMA2
  float2 d;
  asm("add.rn.f32x2 %0, %1, %2;"
      : "=l"(reinterpret_cast<unsigned long long&>(d))
      : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)));
  return d;
#else
  return make_float2(a.x + b.x, a.y + b.y);
#endif
}
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// cp.async (LDGSTS): global -> shared without staging registers; used to prefetch the next tile's raw rows
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// mbarrier + bulk asynchronous copy (TMA engine, SASS UBLKCP): one thread moves a contiguous block global -> shared and the mbarrier
// counts the bytes; the waiting threads poll the barrier's phase parity
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t mbar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}\n"
               :: "r"(mbar), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t mbar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(mbar) : "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t mbar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* smem, const void* gmem, uint32_t bytes, uint32_t mbar) {   // 16-byte aligned, bytes % 16 == 0
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(smem)), "l"(gmem), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ float expo(float x) { return fast_exp2(x); }
__device__ __forceinline__ double expo(double x) { return exp(x); }

// ---- fp64 exp for arguments in units of 1/128 octave ---------------------------------------------------------------------
// 2^(t/128) = 2^k * T[j] * exp(r ln2/128), n = rint(t) = 128 k + j, r = t - n in [-1/2, 1/2].  T lives in shared memory with
// kExpTab = 1024 entries, T[i] = 2^(i/1024) (the bin loop of scan_moments works in 1/1024 octaves with a shorter polynomial); this
// routine reads every eighth entry.  The degree-4 polynomial has a truncation error of 1.2e-15 relative (the parity bar for the
// moments is 1e-9).  GUARD: arguments below 2^-1000 return 0 and arguments above 2^1015 are clamped; without it the caller
// guarantees |t| < 128000 (fixed softmax shift with 2/tau <= 600).
constexpr int kExpTabLog2 = 10, kExpTab = 1 << kExpTabLog2;
template <bool GUARD>
__device__ __forceinline__ double exp2_128(double t, const double* __restrict__ tab) {
  const double magic = 6755399441055744.0;  // 2^52 + 2^51
  if (GUARD) t = fmin(t, 130000.0);
  const double tm = t + magic;
  const int n = __double2loint(tm);
  const double r = t - (tm - magic);
  double p = 3.583032305400251e-11;
  p = fma(p, r, 2.6466421444330967e-08);
  p = fma(p, r, 1.4662262387640423e-05);
  p = fma(p, r, 0.0054152123481245725);
  p = fma(p, r, 1.0);
  const double y = tab[(n & 127) << (kExpTabLog2 - 7)] * p;
  const double z = __hiloint2double(__double2hiint(y) + ((n >> 7) << 20), __double2loint(y));
  if (GUARD) return (t < -128000.0) ? 0.0 : z;
  return z;
}

// ------------------------------------------------------------------------------------------------------------
// per-point front end (phase A, before the bin loop)
// ------------------------------------------------------------------------------------------------------------
struct PointOut64 { double p0[3], d[3], w_time; };

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

// fp64, same maps as point_front64 with cheaper evaluation (results agree to ~1e-15 relative; the parity bar is 1e-9):
// reciprocals of the per-scan constants hoisted out (TimeConst), power series for the Rodrigues coefficients below
// |alpha phi| = 0.2 rad (truncation < 1e-18), the two sigmoids as one division over table-driven exponentials, and
// 1/(|r| + eps) = rs - eps rs^2 with rs = rsqrt(|r|^2) (second-order term eps^2 rs^3 ~ 1e-24).
struct TimeConst { double inv_denom, c_sig; };  // 1/max(t1-t0,1e-12);  128 log2(e) / sigma
__device__ __forceinline__ TimeConst time_const(const gcl_job& J) {
  const double denom = fmax(J.scan_end_time - J.scan_start_time, 1e-12);
  const double sigma = fmax(0.1 * denom, 1e-6);  // GC_TIME_WARP_SIGMA_FRAC, constants.py:143; floor imu_preintegration.py:32
  TimeConst tc;
  tc.inv_denom = 1.0 / denom;
  tc.c_sig = 184.6649652337873 / sigma;
  return tc;
}
__device__ __forceinline__ void point_front64_fast(const gcl_job& J, const ScanConst& C, const TimeConst& tc, const double* tab,
                                                   double px, double py, double pz, double t, PointOut64& o) {
  const double alpha = (t - J.scan_start_time) * tc.inv_denom;
  const double r0 = alpha * J.xi_body[0], r1 = alpha * J.xi_body[1], r2 = alpha * J.xi_body[2];
  const double f0 = alpha * J.xi_body[3], f1 = alpha * J.xi_body[4], f2 = alpha * J.xi_body[5];
  const double th2 = f0 * f0 + f1 * f1 + f2 * f2;
  if (th2 < 0.04) {
    // sinc = sum (-1)^k th2^k/(2k+1)!, B = sum (-1)^k th2^k/(2k+2)!, C = sum (-1)^k th2^k/(2k+3)!   (se3_jax.py:259-301,473-504)
    double sinc = -1.0 / 39916800.0, Bv = 1.0 / 479001600.0 * -1.0, Cv = -1.0 / 6227020800.0;
    sinc = fma(sinc, th2, 1.0 / 362880.0);   Bv = fma(Bv, th2, 1.0 / 3628800.0);   Cv = fma(Cv, th2, 1.0 / 39916800.0);
    sinc = fma(sinc, th2, -1.0 / 5040.0);    Bv = fma(Bv, th2, -1.0 / 40320.0);    Cv = fma(Cv, th2, -1.0 / 362880.0);
    sinc = fma(sinc, th2, 1.0 / 120.0);      Bv = fma(Bv, th2, 1.0 / 720.0);       Cv = fma(Cv, th2, 1.0 / 5040.0);
    sinc = fma(sinc, th2, -1.0 / 6.0);       Bv = fma(Bv, th2, -1.0 / 24.0);       Cv = fma(Cv, th2, -1.0 / 120.0);
    sinc = fma(sinc, th2, 1.0);              Bv = fma(Bv, th2, 0.5);               Cv = fma(Cv, th2, 1.0 / 6.0);
    const double c0 = f1 * r2 - f2 * r1, c1 = f2 * r0 - f0 * r2, c2 = f0 * r1 - f1 * r0;
    const double e0 = f1 * c2 - f2 * c1, e1 = f2 * c0 - f0 * c2, e2 = f0 * c1 - f1 * c0;
    const double u0 = px - (r0 + Bv * c0 + Cv * e0), u1 = py - (r1 + Bv * c1 + Cv * e1), u2 = pz - (r2 + Bv * c2 + Cv * e2);
    const double g0 = f1 * u2 - f2 * u1, g1 = f2 * u0 - f0 * u2, g2 = f0 * u1 - f1 * u0;
    const double h0 = f1 * g2 - f2 * g1, h1 = f2 * g0 - f0 * g2, h2 = f0 * g1 - f1 * g0;
    o.p0[0] = u0 - sinc * g0 + Bv * h0;
    o.p0[1] = u1 - sinc * g1 + Bv * h1;
    o.p0[2] = u2 - sinc * g2 + Bv * h2;
  } else {
    const double p[3] = {px, py, pz};
    deskew_point64(J.xi_body, alpha, p, o.p0);
  }
  // w_t = sigmoid(a) sigmoid(b) (1 - floor) + floor, a = (t-t0)/sigma, b = (t1-t)/sigma   (imu_preintegration.py:34-43)
  const double ea = exp2_128<true>(-(t - J.scan_start_time) * tc.c_sig, tab), eb = exp2_128<true>(-(J.scan_end_time - t) * tc.c_sig, tab);
  o.w_time = (1.0 - 1e-12) / ((1.0 + ea) * (1.0 + eb)) + 1e-12;
  const double q0 = o.p0[0] - C.origin[0], q1 = o.p0[1] - C.origin[1], q2 = o.p0[2] - C.origin[2];
  const double n2 = q0 * q0 + q1 * q1 + q2 * q2;
  double inv;
  if (n2 > 1e-200) {
    const double rs = rsqrt(n2);
    inv = rs - C.eps_mass * rs * rs;
  } else {
    inv = 1.0 / (sqrt(n2) + C.eps_mass);
  }
  o.d[0] = q0 * inv; o.d[1] = q1 * inv; o.d[2] = q2 * inv;
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

}  // namespace gcl
