// gcl_math.cuh -- small fp64 device math used by the finalisation / evidence kernels.
// 3x3 matrices are row-major double[9].  All routines are branch-light, fixed-cost and deterministic.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace gcl {

constexpr double kSmallAngle = 1e-7;  // pkg/common/geometry/se3_jax.py:32
constexpr double kNearPi = 1e-7;      // pkg/common/geometry/se3_jax.py:36
constexpr double kF64Eps = 2.220446049250313e-16;

__host__ __device__ inline void mat3_mul(const double* A, const double* B, double* C) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) C[i * 3 + j] = A[i * 3 + 0] * B[0 * 3 + j] + A[i * 3 + 1] * B[1 * 3 + j] + A[i * 3 + 2] * B[2 * 3 + j];
}
__host__ __device__ inline void mat3_mul_nt(const double* A, const double* B, double* C) {  // A * B^T
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) C[i * 3 + j] = A[i * 3 + 0] * B[j * 3 + 0] + A[i * 3 + 1] * B[j * 3 + 1] + A[i * 3 + 2] * B[j * 3 + 2];
}
__host__ __device__ inline void mat3_mul_tn(const double* A, const double* B, double* C) {  // A^T * B
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) C[i * 3 + j] = A[0 * 3 + i] * B[0 * 3 + j] + A[1 * 3 + i] * B[1 * 3 + j] + A[2 * 3 + i] * B[2 * 3 + j];
}
__host__ __device__ inline double mat3_det(const double* A) {
  return A[0] * (A[4] * A[8] - A[5] * A[7]) - A[1] * (A[3] * A[8] - A[5] * A[6]) + A[2] * (A[3] * A[7] - A[4] * A[6]);
}
__host__ __device__ inline double fro_norm9(const double* A) {
  double s = 0;
  for (int i = 0; i < 9; ++i) s += A[i] * A[i];
  return sqrt(s);
}
__host__ __device__ inline double sigmoid64(double x) { return 1.0 / (1.0 + exp(-x)); }

// ---- SO(3) -----------------------------------------------------------------------------------------------
// so3_exp, pkg/common/geometry/se3_jax.py:259-301.
__host__ __device__ inline void so3_exp(const double* w, double* R) {
  const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  const double th = sqrt(th2);
  const bool small = th < kSmallAngle;
  const double sth = small ? 1.0 : th;
  const double sth2 = (th2 < kSmallAngle * kSmallAngle) ? 1.0 : th2;
  const double a = small ? 1.0 : sin(sth) / sth;
  const double b = small ? 0.5 : (1.0 - cos(sth)) / sth2;
  const double K[9] = {0.0, -w[2], w[1], w[2], 0.0, -w[0], -w[1], w[0], 0.0};
  double K2[9];
  mat3_mul(K, K, K2);
  for (int i = 0; i < 9; ++i) R[i] = ((i % 4 == 0) ? 1.0 : 0.0) + a * K[i] + b * K2[i];
}

// so3_log, pkg/common/geometry/se3_jax.py:304-366 (softmax-mixture axis near pi).
__host__ __device__ inline void so3_log(const double* R, double* w) {
  double c = 0.5 * (R[0] + R[4] + R[8] - 1.0);
  c = fmin(1.0, fmax(-1.0, c));
  const double th = acos(c);
  const double v[3] = {0.5 * (R[7] - R[5]), 0.5 * (R[2] - R[6]), 0.5 * (R[3] - R[1])};
  const double s = sin(th);
  const double ss = (fabs(s) < kSmallAngle) ? 1.0 : s;
  const double kg = th / (2.0 * ss);
  if (th < kSmallAngle) {
    w[0] = v[0]; w[1] = v[1]; w[2] = v[2];
    return;
  }
  if (fabs(th - 3.141592653589793) < kNearPi) {
    const double z0 = 50.0 * (R[0] + 1.0), z1 = 50.0 * (R[4] + 1.0), z2 = 50.0 * (R[8] + 1.0);
    const double zm = fmax(z0, fmax(z1, z2));
    const double e0 = exp(z0 - zm), e1 = exp(z1 - zm), e2 = exp(z2 - zm);
    const double es = e0 + e1 + e2;
    const double w0 = e0 / es, w1 = e1 / es, w2 = e2 / es;
    double ax[3];
    for (int i = 0; i < 3; ++i)
      ax[i] = w0 * (R[i * 3 + 0] + (i == 0)) + w1 * (R[i * 3 + 1] + (i == 1)) + w2 * (R[i * 3 + 2] + (i == 2));
    const double n = sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
    const double sn = (n < kSmallAngle) ? 1.0 : n;
    for (int i = 0; i < 3; ++i) w[i] = ax[i] / sn * th;
    return;
  }
  for (int i = 0; i < 3; ++i) w[i] = kg * (2.0 * v[i]);
}

// ---- symmetric 3x3 eigen decomposition: cyclic Jacobi, fixed sweep budget -----------------------------------
// A (row-major, symmetric part used) -> ascending eigenvalues `val`, eigenvectors as COLUMNS of V (row-major).
// Matches LAPACK syevd on V diag(val) V^T to a few ulps of ||A|| (eigenvector signs are free; the rebuilt
// matrix is not).  Replaces jnp.linalg.eigh at pkg/common/primitives.py:101 and matrix_fisher_evidence.py:104.
__host__ __device__ inline void sym_eig3(const double* Ain, double* val, double* V) {
  double a00 = Ain[0], a11 = Ain[4], a22 = Ain[8];
  double a01 = 0.5 * (Ain[1] + Ain[3]), a02 = 0.5 * (Ain[2] + Ain[6]), a12 = 0.5 * (Ain[5] + Ain[7]);
  double v[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
#define GCL_JROT(app, aqq, apq, arp, arq, P, Q)                                        \
  if (apq != 0.0) {                                                                     \
    /* t = tan(phi) of the annihilating rotation: t = 2b sgn(d) / (|d| + sqrt(d^2 + 4b^2)), d = aqq - app */ \
    const double dd = aqq - app, b2 = 2.0 * apq;                                        \
    const double t = (dd >= 0.0 ? b2 : -b2) / (fabs(dd) + sqrt(dd * dd + b2 * b2));     \
    const double cs = rsqrt(t * t + 1.0), sn = t * cs;                                  \
    app -= t * apq;                                                                     \
    aqq += t * apq;                                                                     \
    apq = 0.0;                                                                          \
    const double rp = arp, rq = arq;                                                    \
    arp = cs * rp - sn * rq;                                                            \
    arq = sn * rp + cs * rq;                                                            \
    for (int k = 0; k < 3; ++k) {                                                       \
      const double vp = v[k * 3 + P], vq = v[k * 3 + Q];                                \
      v[k * 3 + P] = cs * vp - sn * vq;                                                 \
      v[k * 3 + Q] = sn * vp + cs * vq;                                                 \
    }                                                                                   \
  }
  for (int sweep = 0; sweep < 12; ++sweep) {
    const double off = fabs(a01) + fabs(a02) + fabs(a12);
    const double dg = fabs(a00) + fabs(a11) + fabs(a22);
    if (off == 0.0 || off + dg == dg) break;  // off-diagonal below one ulp of the diagonal
    GCL_JROT(a00, a11, a01, a02, a12, 0, 1)
    GCL_JROT(a00, a22, a02, a01, a12, 0, 2)
    GCL_JROT(a11, a22, a12, a01, a02, 1, 2)
  }
#undef GCL_JROT
  double e[3] = {a00, a11, a22};
  int idx[3] = {0, 1, 2};
  // ascending sort of 3 (stable network)
  if (e[idx[0]] > e[idx[1]]) { int t = idx[0]; idx[0] = idx[1]; idx[1] = t; }
  if (e[idx[1]] > e[idx[2]]) { int t = idx[1]; idx[1] = idx[2]; idx[2] = t; }
  if (e[idx[0]] > e[idx[1]]) { int t = idx[0]; idx[0] = idx[1]; idx[1] = t; }
  for (int j = 0; j < 3; ++j) {
    val[j] = e[idx[j]];
    for (int k = 0; k < 3; ++k) V[k * 3 + j] = v[k * 3 + idx[j]];
  }
}

// domain_projection_psd_core, pkg/common/primitives.py:80-123.
// cert = [projection_delta, sym_delta, eig_min, eig_max, cond, near_null_count]
__host__ __device__ inline void psd_project3(const double* M, double eps_psd, double* Mpsd, double* cert, double* vals_floored = nullptr) {
  double Ms[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) Ms[i * 3 + j] = 0.5 * (M[i * 3 + j] + M[j * 3 + i]);
  double sd = 0;
  for (int i = 0; i < 9; ++i) sd += (Ms[i] - M[i]) * (Ms[i] - M[i]);
  double val[3], V[9];
  sym_eig3(Ms, val, V);
  double vc[3];
  for (int k = 0; k < 3; ++k) vc[k] = fmax(val[k], eps_psd);
  if (vals_floored) { vals_floored[0] = vc[0]; vals_floored[1] = vc[1]; vals_floored[2] = vc[2]; }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      Mpsd[i * 3 + j] = V[i * 3 + 0] * vc[0] * V[j * 3 + 0] + V[i * 3 + 1] * vc[1] * V[j * 3 + 1] + V[i * 3 + 2] * vc[2] * V[j * 3 + 2];
  double pd = 0;
  for (int i = 0; i < 9; ++i) pd += (Mpsd[i] - Ms[i]) * (Mpsd[i] - Ms[i]);
  const double emin = fmin(vc[0], fmin(vc[1], vc[2])), emax = fmax(vc[0], fmax(vc[1], vc[2]));
  cert[0] = sqrt(pd);
  cert[1] = sqrt(sd);
  cert[2] = emin;
  cert[3] = emax;
  cert[4] = emax / emin;
  cert[5] = (double)((vc[0] < 10.0 * eps_psd) + (vc[1] < 10.0 * eps_psd) + (vc[2] < 10.0 * eps_psd));
}

// PSD projection with a shortcut: when the symmetrised matrix is positive definite with every eigenvalue provably above
// eps_psd (Sylvester minors > 0 and lambda_min >= 4 det / tr^2 > 2 eps_psd), flooring changes nothing and
// V diag(lambda) V^T == M_sym up to round-off, so the eigen decomposition is skipped.  Returns the projection delta
// (0 on the shortcut).  Used where only M_psd and the delta are consumed.
__host__ __device__ inline double psd_project3_fast(const double* M, double eps_psd, double* Mpsd) {
  const double a = M[0], d = M[4], f = M[8];
  const double b = 0.5 * (M[1] + M[3]), c = 0.5 * (M[2] + M[6]), e = 0.5 * (M[5] + M[7]);
  const double m2 = a * d - b * b;
  const double det = a * (d * f - e * e) - b * (b * f - e * c) + c * (b * e - d * c);
  const double tr = a + d + f;
  if (a > 0.0 && m2 > 0.0 && det > 0.0 && 4.0 * det > 2.0 * eps_psd * tr * tr) {
    Mpsd[0] = a; Mpsd[1] = b; Mpsd[2] = c; Mpsd[3] = b; Mpsd[4] = d; Mpsd[5] = e; Mpsd[6] = c; Mpsd[7] = e; Mpsd[8] = f;
    return 0.0;
  }
  double pc[6];
  psd_project3(M, eps_psd, Mpsd, pc);
  return pc[0];
}

// ---- 3x3 SVD: one-sided (Hestenes) Jacobi ------------------------------------------------------------------
// H = U diag(s) V^T, s descending, U and V orthogonal (full matrices), as jnp.linalg.svd(H, full_matrices=True)
// (matrix_fisher_evidence.py:215).  For a rank-deficient H the missing columns of U are completed orthonormally;
// the reflection fix in the caller makes R_mf independent of that choice when s1 > s2.
__host__ __device__ inline void svd3(const double* H, double* U, double* s, double* V) {
  double a[9], v[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  for (int i = 0; i < 9; ++i) a[i] = H[i];
  for (int sweep = 0; sweep < 30; ++sweep) {
    int rotated = 0;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        double al = 0, be = 0, ga = 0;
        for (int k = 0; k < 3; ++k) {
          al += a[k * 3 + p] * a[k * 3 + p];
          be += a[k * 3 + q] * a[k * 3 + q];
          ga += a[k * 3 + p] * a[k * 3 + q];
        }
        if (ga == 0.0 || ga * ga <= 1e-30 * (al * be)) continue;  // columns orthogonal to working precision
        rotated = 1;
        // t = tan of the rotation that orthogonalises columns p, q: t = 2 ga sgn(d) / (|d| + sqrt(d^2 + 4 ga^2)), d = be - al
        // (the same number as sgn(zeta) / (|zeta| + sqrt(1 + zeta^2)) with zeta = d / (2 ga), one division and one square root less)
        const double dd = be - al, g2 = 2.0 * ga;
        const double t = (dd >= 0.0 ? g2 : -g2) / (fabs(dd) + sqrt(dd * dd + g2 * g2));
        const double cs = rsqrt(1.0 + t * t), sn = cs * t;
        for (int k = 0; k < 3; ++k) {
          const double ap = a[k * 3 + p], aq = a[k * 3 + q];
          a[k * 3 + p] = cs * ap - sn * aq;
          a[k * 3 + q] = sn * ap + cs * aq;
          const double vp = v[k * 3 + p], vq = v[k * 3 + q];
          v[k * 3 + p] = cs * vp - sn * vq;
          v[k * 3 + q] = sn * vp + cs * vq;
        }
      }
    if (!rotated) break;
  }
  double nrm[3];
  for (int j = 0; j < 3; ++j) nrm[j] = sqrt(a[0 * 3 + j] * a[0 * 3 + j] + a[1 * 3 + j] * a[1 * 3 + j] + a[2 * 3 + j] * a[2 * 3 + j]);
  int idx[3] = {0, 1, 2};
  if (nrm[idx[0]] < nrm[idx[1]]) { int t = idx[0]; idx[0] = idx[1]; idx[1] = t; }
  if (nrm[idx[1]] < nrm[idx[2]]) { int t = idx[1]; idx[1] = idx[2]; idx[2] = t; }
  if (nrm[idx[0]] < nrm[idx[1]]) { int t = idx[0]; idx[0] = idx[1]; idx[1] = t; }
  const double tiny = 1e-300;
  double u[9];
  for (int j = 0; j < 3; ++j) {
    const int c = idx[j];
    s[j] = nrm[c];
    for (int k = 0; k < 3; ++k) {
      V[k * 3 + j] = v[k * 3 + c];
      u[k * 3 + j] = (nrm[c] > tiny) ? a[k * 3 + c] / nrm[c] : 0.0;
    }
  }
  // complete U when singular values vanish (relative to s0)
  const double thr = s[0] * 1e-15;
  if (!(s[0] > tiny)) {
    for (int i = 0; i < 9; ++i) u[i] = (i % 4 == 0) ? 1.0 : 0.0;
  } else {
    if (!(s[1] > thr)) {  // pick any unit vector orthogonal to u0
      const double x = u[0], y = u[3], z = u[6];
      double b[3];
      if (fabs(x) <= fabs(y) && fabs(x) <= fabs(z)) { b[0] = 0; b[1] = -z; b[2] = y; }
      else if (fabs(y) <= fabs(z)) { b[0] = -z; b[1] = 0; b[2] = x; }
      else { b[0] = -y; b[1] = x; b[2] = 0; }
      const double n = sqrt(b[0] * b[0] + b[1] * b[1] + b[2] * b[2]);
      u[1] = b[0] / n; u[4] = b[1] / n; u[7] = b[2] / n;
    }
    if (!(s[2] > thr)) {  // u2 = u0 x u1
      u[2] = u[3] * u[7] - u[6] * u[4];
      u[5] = u[6] * u[1] - u[0] * u[7];
      u[8] = u[0] * u[4] - u[3] * u[1];
      const double n = sqrt(u[2] * u[2] + u[5] * u[5] + u[8] * u[8]);
      if (n > tiny) { u[2] /= n; u[5] /= n; u[8] /= n; }
    }
  }
  for (int i = 0; i < 9; ++i) U[i] = u[i];
}

// ---- 3x3 LU with partial pivoting (the algorithm behind jnp.linalg.inv / solve = LAPACK getrf) --------------
__host__ __device__ inline bool lu3(double* a, int* piv) {
  for (int k = 0; k < 3; ++k) {
    int p = k;
    double mx = fabs(a[k * 3 + k]);
    for (int i = k + 1; i < 3; ++i)
      if (fabs(a[i * 3 + k]) > mx) { mx = fabs(a[i * 3 + k]); p = i; }
    piv[k] = p;
    if (p != k)
      for (int j = 0; j < 3; ++j) { const double t = a[k * 3 + j]; a[k * 3 + j] = a[p * 3 + j]; a[p * 3 + j] = t; }
    const double d = a[k * 3 + k];
    for (int i = k + 1; i < 3; ++i) {
      const double l = a[i * 3 + k] / d;
      a[i * 3 + k] = l;
      for (int j = k + 1; j < 3; ++j) a[i * 3 + j] -= l * a[k * 3 + j];
    }
  }
  return true;
}
__host__ __device__ inline void lu3_solve(const double* lu, const int* piv, double* b) {
  for (int k = 0; k < 3; ++k)
    if (piv[k] != k) { const double t = b[k]; b[k] = b[piv[k]]; b[piv[k]] = t; }
  b[1] -= lu[3] * b[0];
  b[2] -= lu[6] * b[0] + lu[7] * b[1];
  b[2] = b[2] / lu[8];
  b[1] = (b[1] - lu[5] * b[2]) / lu[4];
  b[0] = (b[0] - lu[1] * b[1] - lu[2] * b[2]) / lu[0];
}
__host__ __device__ inline void mat3_inv_lu(const double* A, double* Ainv) {
  double lu[9];
  int piv[3];
  for (int i = 0; i < 9; ++i) lu[i] = A[i];
  lu3(lu, piv);
  for (int c = 0; c < 3; ++c) {
    double e[3] = {c == 0 ? 1.0 : 0.0, c == 1 ? 1.0 : 0.0, c == 2 ? 1.0 : 0.0};
    lu3_solve(lu, piv, e);
    Ainv[0 * 3 + c] = e[0]; Ainv[1 * 3 + c] = e[1]; Ainv[2 * 3 + c] = e[2];
  }
}

// 3x3 inverse by cofactors: one division, no pivot search.  Used for the per-bin covariance sums of the translation evidence
// (matrix_fisher_evidence.py:462-466 calls jnp.linalg.inv): those matrices are symmetric positive definite up to round-off, for
// which the cofactor formula is as accurate as partially pivoted LU (relative error ~ cond * eps in both cases).
__host__ __device__ inline void mat3_inv_adj(const double* A, double* Ainv) {
  const double c00 = A[4] * A[8] - A[5] * A[7], c01 = A[5] * A[6] - A[3] * A[8], c02 = A[3] * A[7] - A[4] * A[6];
  const double det = A[0] * c00 + A[1] * c01 + A[2] * c02;
  const double id = 1.0 / det;
  Ainv[0] = c00 * id; Ainv[3] = c01 * id; Ainv[6] = c02 * id;
  Ainv[1] = (A[2] * A[7] - A[1] * A[8]) * id; Ainv[4] = (A[0] * A[8] - A[2] * A[6]) * id; Ainv[7] = (A[1] * A[6] - A[0] * A[7]) * id;
  Ainv[2] = (A[1] * A[5] - A[2] * A[4]) * id; Ainv[5] = (A[2] * A[3] - A[0] * A[5]) * id; Ainv[8] = (A[0] * A[4] - A[1] * A[3]) * id;
}

// ---- kappa, pkg/backend/operators/kappa.py:130-169 -----------------------------------------------------------
__host__ __device__ inline double kappa_from_resultant(double Rbar, double eps_r, double d, double r0, double tau) {
  const double R = fmin(fmax(Rbar, 0.0), 1.0 - eps_r);
  const double R2 = R * R;
  const double k_low = (R * (d - R2)) / (1.0 - R2 + eps_r);
  const double k_high = -log(fmax(1.0 - R2, eps_r));
  const double s = sigmoid64((R - r0) / fmax(tau, 1e-6));
  return (1.0 - s) * k_low + s * k_high;
}

// ---- constant-twist deskew of one point, fp64: deskew_constant_twist.py:47-56 + se3_jax.py:473-504,:259-301 --
// xi = [rho, phi]; alpha scalar; returns p0 = R(alpha phi)^T (p - V(alpha phi) alpha rho)
__host__ __device__ inline void deskew_point64(const double* xi, double alpha, const double* p, double* p0) {
  const double r0 = alpha * xi[0], r1 = alpha * xi[1], r2 = alpha * xi[2];
  const double f0 = alpha * xi[3], f1 = alpha * xi[4], f2 = alpha * xi[5];
  const double th2 = f0 * f0 + f1 * f1 + f2 * f2;
  const double th = sqrt(th2);
  const bool small = th < kSmallAngle;
  const double sth = small ? 1.0 : th;
  const double sth2 = (th2 < kSmallAngle * kSmallAngle) ? 1.0 : th2;
  double sn, cs;
  sincos(sth, &sn, &cs);
  const double Bv = small ? (0.5 - th2 / 24.0) : (1.0 - cs) / sth2;
  const double Cv = small ? (1.0 / 6.0 - th2 / 120.0) : (sth - sn) / (sth2 * sth);
  const double sinc = small ? 1.0 : sn / sth;
  const double Br = small ? 0.5 : (1.0 - cs) / sth2;
  // t = rho + Bv (phi x rho) + Cv (phi x (phi x rho))
  const double c0 = f1 * r2 - f2 * r1, c1 = f2 * r0 - f0 * r2, c2 = f0 * r1 - f1 * r0;
  const double d0 = f1 * c2 - f2 * c1, d1 = f2 * c0 - f0 * c2, d2 = f0 * c1 - f1 * c0;
  const double u0 = p[0] - (r0 + Bv * c0 + Cv * d0), u1 = p[1] - (r1 + Bv * c1 + Cv * d1), u2 = p[2] - (r2 + Bv * c2 + Cv * d2);
  // R^T u = u - sinc (phi x u) + Br (phi x (phi x u))
  const double g0 = f1 * u2 - f2 * u1, g1 = f2 * u0 - f0 * u2, g2 = f0 * u1 - f1 * u0;
  const double h0 = f1 * g2 - f2 * g1, h1 = f2 * g0 - f0 * g2, h2 = f0 * g1 - f1 * g0;
  p0[0] = u0 - sinc * g0 + Br * h0;
  p0[1] = u1 - sinc * g1 + Br * h1;
  p0[2] = u2 - sinc * g2 + Br * h2;
}

// smooth_window_weights for one stamp, imu_preintegration.py:34-43 (sigma already floored at 1e-6)
__host__ __device__ inline double window_weight64(double t, double t0, double t1, double sigma, double floor_w) {
  const double a = (t - t0) / sigma, b = (t1 - t) / sigma;
  return sigmoid64(a) * sigmoid64(b) * (1.0 - floor_w) + floor_w;
}

}  // namespace gcl
