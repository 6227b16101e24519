// exact_math.h — IEEE-only scalar math shared by host and device code.
//
// Every routine here is built from +, -, *, /, sqrt, rint and integer bit
// operations only, so it evaluates to the SAME bits on x86-64 (g++ with
// -ffp-contract=off) and on sm_100a (nvcc with -fmad=false).  The registration
// path makes discrete decisions from floating-point values (DoG extrema,
// histogram bins, RANSAC inlier votes, ICP iteration counts); routing the few
// transcendental functions through this header is what lets the CUDA path and
// the CPU checker agree on those decisions bit-for-bit instead of "almost".
//
// The functions are accurate to the last float bit in all but a vanishing
// fraction of inputs (tests/test_exact_math.py pins them against libm/numpy).
//
// Reference call sites these stand in for ([PCL-recall]):
//   expf   — pcl/keypoints/impl/sift_keypoint.hpp (Gaussian weights)
//   atan2f — pcl/features/impl/pfh.hpp computePairFeatures (f1)
//   atan2/cos/sin — pcl/common/impl/eigen.hpp computeRoots (normals, RANSAC)
//   JacobiSVD 3x3 — Eigen::umeyama via pcl::umeyama (RANSAC, ICP, final SVD)
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define MM_HD __host__ __device__ __forceinline__
#else
#define MM_HD inline
#endif

namespace mm3d {
namespace em {

MM_HD double bits_to_double(uint64_t u)
{
  double d;
  memcpy(&d, &u, sizeof(d));
  return d;
}

MM_HD float bits_to_float(uint32_t u)
{
  float f;
  memcpy(&f, &u, sizeof(f));
  return f;
}

// exp(x) in float only: Cody-Waite reduction x = n ln2 + r, |r| <= ln2/2, degree-7 polynomial,
// exact power-of-two scaling.  Within 1 ulp of the correctly rounded result (tests pin it against libm).
MM_HD float expf_(float x)
{
  if (x != x) return x;
  if (x < -104.0f) return 0.0f;
  if (x > 88.72f) return INFINITY;
  const float n = rintf(x * 1.44269504088896341f);
  // ln2 split: the high part has 9 significant bits, so n * hi is exact for |n| < 2^15
  float r = x - n * 0.693359375f;
  r = r - n * -2.12194440e-4f;
  float p = 1.9875691500e-4f;
  p = p * r + 1.3981999507e-3f;
  p = p * r + 8.3334519073e-3f;
  p = p * r + 4.1665795894e-2f;
  p = p * r + 1.6666665459e-1f;
  p = p * r + 5.0000001201e-1f;
  const float y = (p * (r * r) + r) + 1.0f;
  int ni = (int)n;  // in [-151, 128]
  // 2^ni in two exact factors so that sub-normal results and ni = 128 are handled
  const int h = ni / 2;
  const float s1 = bits_to_float((uint32_t)(127 + h) << 23);
  const float s2 = bits_to_float((uint32_t)(127 + (ni - h)) << 23);
  return (y * s1) * s2;
}

// atan on [0, inf) in float: three-interval reduction, odd degree-9 polynomial, the constants pi/2 and
// pi/4 carried as hi + lo pairs.  <= 2 ulp from the correctly rounded value (tests pin it against libm).
MM_HD float atanf_pos_(float x, float* lo)
{
  float y0, y0lo;
  if (x > 2.414213562373095f) {  // tan(3 pi / 8)
    y0 = 1.57079637050628662109375f;
    y0lo = -4.37113900018624283e-8f;
    x = -(1.0f / x);
  } else if (x > 0.4142135623730950f) {  // tan(pi / 8)
    y0 = 0.785398185253143310546875f;
    y0lo = -2.18556950009312142e-8f;
    x = (x - 1.0f) / (x + 1.0f);
  } else {
    y0 = 0.0f;
    y0lo = 0.0f;
  }
  const float z = x * x;
  float p = 8.05374449538e-2f;
  p = p * z - 1.38776856032e-1f;
  p = p * z + 1.99777106478e-1f;
  p = p * z - 3.33329491539e-1f;
  const float corr = p * z * x + y0lo;  // small terms first
  const float hi = y0 + x;              // exact or nearly so: x is small against y0 or y0 is 0
  *lo = ((y0 - hi) + x) + corr;         // rounding error of hi + the small terms
  return hi;
}

MM_HD float atan2f_(float y, float x)
{
  if (x != x || y != y) return x + y;
  const float ax = fabsf(x), ay = fabsf(y);
  float hi = 0.0f, lo = 0.0f;
  if (!(ax == 0.0f && ay == 0.0f)) hi = atanf_pos_(ay / ax, &lo);  // ay / 0 = inf -> pi/2
  float a;
  if (signbit(x)) {
    // pi - (hi + lo) with pi = pi_hi + pi_lo
    const float d = 3.1415927410125732421875f - hi;
    a = d + (-8.74227800037248566e-8f - lo);
  } else {
    a = hi + lo;
  }
  return signbit(y) ? -a : a;
}

// sin/cos for |t| <= ~1.2 (the eigen-solver only needs [0, pi/3]).
MM_HD double cos_small_(double t)
{
  const double t2 = t * t;
  double s = 1.0 / 6402373705728000.0;  // 1/18!
  s = 1.0 / 20922789888000.0 - t2 * s;  // 1/16!
  s = 1.0 / 87178291200.0 - t2 * s;     // 1/14!
  s = 1.0 / 479001600.0 - t2 * s;       // 1/12!
  s = 1.0 / 3628800.0 - t2 * s;         // 1/10!
  s = 1.0 / 40320.0 - t2 * s;           // 1/8!
  s = 1.0 / 720.0 - t2 * s;             // 1/6!
  s = 1.0 / 24.0 - t2 * s;              // 1/4!
  s = 0.5 - t2 * s;
  return 1.0 - t2 * s;
}

MM_HD double sin_small_(double t)
{
  const double t2 = t * t;
  double s = 1.0 / 121645100408832000.0;  // 1/19!
  s = 1.0 / 355687428096000.0 - t2 * s;   // 1/17!
  s = 1.0 / 1307674368000.0 - t2 * s;     // 1/15!
  s = 1.0 / 6227020800.0 - t2 * s;        // 1/13!
  s = 1.0 / 39916800.0 - t2 * s;          // 1/11!
  s = 1.0 / 362880.0 - t2 * s;            // 1/9!
  s = 1.0 / 5040.0 - t2 * s;              // 1/7!
  s = 1.0 / 120.0 - t2 * s;               // 1/5!
  s = 1.0 / 6.0 - t2 * s;                 // 1/3!
  return t * (1.0 - t2 * s);
}

MM_HD float cosf_small_(float t) { return (float)cos_small_((double)t); }
MM_HD float sinf_small_(float t) { return (float)sin_small_((double)t); }

template <typename T>
MM_HD T sqrt_(T v);
template <>
MM_HD float sqrt_<float>(float v) { return sqrtf(v); }
template <>
MM_HD double sqrt_<double>(double v) { return sqrt(v); }

// ---------------------------------------------------------------------------
// 3x3 helpers (row-major m[9]).
// ---------------------------------------------------------------------------
template <typename T>
MM_HD T det3(const T* m)
{
  return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) +
         m[2] * (m[3] * m[7] - m[4] * m[6]);
}

// One-sided (Hestenes) Jacobi SVD of a 3x3 matrix: A = U diag(s) V^T with
// s[0] >= s[1] >= s[2] >= 0.  Rank-deficient columns of U are completed to an
// orthonormal basis deterministically (zero matrix -> U = V = I, which is what
// Eigen::JacobiSVD returns for it).
template <typename T>
MM_HD void svd3(const T* A, T* U, T* s, T* V)
{
  T a[9];
  for (int i = 0; i < 9; ++i) {
    a[i] = A[i];
    V[i] = T(0);
  }
  V[0] = V[4] = V[8] = T(1);
  const T eps = sizeof(T) == 4 ? T(1.1920929e-7) : T(2.220446049250313e-16);
  for (int sweep = 0; sweep < 30; ++sweep) {
    bool rotated = false;
    for (int p = 0; p < 2; ++p) {
      for (int q = p + 1; q < 3; ++q) {
        T alpha = T(0), beta = T(0), gamma = T(0);
        for (int i = 0; i < 3; ++i) {
          alpha += a[i * 3 + p] * a[i * 3 + p];
          beta += a[i * 3 + q] * a[i * 3 + q];
          gamma += a[i * 3 + p] * a[i * 3 + q];
        }
        if (gamma == T(0)) continue;
        if (fabs(gamma) <= eps * sqrt_<T>(alpha * beta)) continue;
        rotated = true;
        const T zeta = (beta - alpha) / (T(2) * gamma);
        const T t = (zeta >= T(0) ? T(1) : T(-1)) / (fabs(zeta) + sqrt_<T>(T(1) + zeta * zeta));
        const T c = T(1) / sqrt_<T>(T(1) + t * t);
        const T sn = c * t;
        for (int i = 0; i < 3; ++i) {
          const T ap = a[i * 3 + p], aq = a[i * 3 + q];
          a[i * 3 + p] = c * ap - sn * aq;
          a[i * 3 + q] = sn * ap + c * aq;
          const T vp = V[i * 3 + p], vq = V[i * 3 + q];
          V[i * 3 + p] = c * vp - sn * vq;
          V[i * 3 + q] = sn * vp + c * vq;
        }
      }
    }
    if (!rotated) break;
  }
  T n[3];
  for (int j = 0; j < 3; ++j)
    n[j] = sqrt_<T>(a[j] * a[j] + a[3 + j] * a[3 + j] + a[6 + j] * a[6 + j]);
  // sort columns by descending norm (stable selection)
  int ord[3] = {0, 1, 2};
  for (int i = 0; i < 2; ++i)
    for (int j = i + 1; j < 3; ++j)
      if (n[ord[j]] > n[ord[i]]) {
        int tmp = ord[i];
        ord[i] = ord[j];
        ord[j] = tmp;
      }
  T Vs[9], As[9];
  for (int j = 0; j < 3; ++j) {
    s[j] = n[ord[j]];
    for (int i = 0; i < 3; ++i) {
      Vs[i * 3 + j] = V[i * 3 + ord[j]];
      As[i * 3 + j] = a[i * 3 + ord[j]];
    }
  }
  for (int i = 0; i < 9; ++i) V[i] = Vs[i];
  const T tiny = s[0] * eps * T(8);
  int rank = 0;
  for (int j = 0; j < 3; ++j)
    if (s[j] > tiny) ++rank;
  if (rank == 0) {
    for (int i = 0; i < 9; ++i) U[i] = T(0);
    U[0] = U[4] = U[8] = T(1);
    return;
  }
  for (int j = 0; j < rank; ++j)
    for (int i = 0; i < 3; ++i) U[i * 3 + j] = As[i * 3 + j] / s[j];
  if (rank == 1) {
    // pick the axis least aligned with u0, orthogonalise, normalise
    const T x = fabs(U[0]), y = fabs(U[3]), z = fabs(U[6]);
    T e[3] = {T(0), T(0), T(0)};
    if (x <= y && x <= z)
      e[0] = T(1);
    else if (y <= z)
      e[1] = T(1);
    else
      e[2] = T(1);
    const T d = e[0] * U[0] + e[1] * U[3] + e[2] * U[6];
    T w[3] = {e[0] - d * U[0], e[1] - d * U[3], e[2] - d * U[6]};
    const T wn = sqrt_<T>(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    U[1] = w[0] / wn;
    U[4] = w[1] / wn;
    U[7] = w[2] / wn;
  }
  if (rank <= 2) {
    U[2] = U[3] * U[7] - U[6] * U[4];
    U[5] = U[6] * U[1] - U[0] * U[7];
    U[8] = U[0] * U[4] - U[3] * U[1];
  }
}

// Rigid (no scale) Umeyama from the 3x3 cross-covariance sigma = E[(q-qm)(p-pm)^T]
// and the two means; writes a row-major 4x4.  Follows Eigen::umeyama
// (Eigen/src/Geometry/Umeyama.h) step by step [PCL-recall: pcl::umeyama].
template <typename T>
MM_HD void umeyama_from_sigma(const T* sigma, const T* src_mean, const T* dst_mean, T* Rt)
{
  T U[9], V[9], d[3];
  svd3<T>(sigma, U, d, V);
  T S[3] = {T(1), T(1), T(1)};
  if (det3<T>(sigma) < T(0)) S[2] = T(-1);
  const T prec = sizeof(T) == 4 ? T(1e-5) : T(1e-12);
  int rank = 0;
  for (int i = 0; i < 3; ++i)
    if (!(fabs(d[i]) <= fabs(d[0]) * prec)) ++rank;
  if (rank == 2) {
    if (det3<T>(U) * det3<T>(V) > T(0)) {
      S[2] = T(1);
    } else {
      S[2] = T(-1);
    }
  }
  T R[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      // (U * S.asDiagonal()) * V^T
      T acc = (U[i * 3 + 0] * S[0]) * V[j * 3 + 0];
      acc += (U[i * 3 + 1] * S[1]) * V[j * 3 + 1];
      acc += (U[i * 3 + 2] * S[2]) * V[j * 3 + 2];
      R[i * 3 + j] = acc;
    }
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) Rt[i * 4 + j] = R[i * 3 + j];
    T rs = R[i * 3 + 0] * src_mean[0];
    rs += R[i * 3 + 1] * src_mean[1];
    rs += R[i * 3 + 2] * src_mean[2];
    Rt[i * 4 + 3] = dst_mean[i] - rs;
  }
  Rt[12] = Rt[13] = Rt[14] = T(0);
  Rt[15] = T(1);
}

// ---------------------------------------------------------------------------
// pcl::eigen33 / computeRoots restated in float [PCL-recall
// pcl/common/impl/eigen.hpp].  m is a symmetric row-major 3x3.
// ---------------------------------------------------------------------------
MM_HD void compute_roots2_(float b, float c, float* roots)
{
  roots[0] = 0.0f;
  float d = (float)((double)(b * b) - 4.0 * (double)c);
  if (d < 0.0f) d = 0.0f;
  const float sd = sqrtf(d);
  roots[2] = 0.5f * (b + sd);
  roots[1] = 0.5f * (b - sd);
}

MM_HD void compute_roots_(const float* m, float* roots)
{
  const float c0 = m[0] * m[4] * m[8] + 2.0f * m[1] * m[2] * m[5] - m[0] * m[5] * m[5] -
                   m[4] * m[2] * m[2] - m[8] * m[1] * m[1];
  const float c1 = m[0] * m[4] - m[1] * m[1] + m[0] * m[8] - m[2] * m[2] + m[4] * m[8] - m[5] * m[5];
  const float c2 = m[0] + m[4] + m[8];
  if (fabsf(c0) < 1.1920929e-7f) {
    compute_roots2_(c2, c1, roots);
    return;
  }
  const float s_inv3 = (float)(1.0 / 3.0);
  const float s_sqrt3 = sqrtf(3.0f);
  const float c2_over_3 = c2 * s_inv3;
  float a_over_3 = (c1 - c2 * c2_over_3) * s_inv3;
  if (a_over_3 > 0.0f) a_over_3 = 0.0f;
  const float half_b = 0.5f * (c0 + c2_over_3 * (2.0f * c2_over_3 * c2_over_3 - c1));
  float q = half_b * half_b + a_over_3 * a_over_3 * a_over_3;
  if (q > 0.0f) q = 0.0f;
  const float rho = sqrtf(-a_over_3);
  const float theta = atan2f_(sqrtf(-q), half_b) * s_inv3;
  const float cos_theta = cosf_small_(theta);
  const float sin_theta = sinf_small_(theta);
  roots[0] = c2_over_3 + 2.0f * rho * cos_theta;
  roots[1] = c2_over_3 - rho * (cos_theta + s_sqrt3 * sin_theta);
  roots[2] = c2_over_3 - rho * (cos_theta - s_sqrt3 * sin_theta);
  if (roots[0] >= roots[1]) {
    float t = roots[0];
    roots[0] = roots[1];
    roots[1] = t;
  }
  if (roots[1] >= roots[2]) {
    float t = roots[1];
    roots[1] = roots[2];
    roots[2] = t;
    if (roots[0] >= roots[1]) {
      t = roots[0];
      roots[0] = roots[1];
      roots[1] = t;
    }
  }
  if (roots[0] <= 0.0f) compute_roots2_(c2, c1, roots);
}

// eigenvalues only (ascending), pcl::eigen33(mat, evals)
MM_HD void eigen33_values(const float* mat, float* evals)
{
  float scale = 0.0f;
  for (int i = 0; i < 9; ++i) scale = fmaxf(scale, fabsf(mat[i]));
  if (scale <= 1.17549435e-38f) scale = 1.0f;
  float sm[9];
  for (int i = 0; i < 9; ++i) sm[i] = mat[i] / scale;
  compute_roots_(sm, evals);
  for (int i = 0; i < 3; ++i) evals[i] *= scale;
}

// smallest eigenpair, pcl::eigen33(mat, eigenvalue, eigenvector)
MM_HD void eigen33_smallest(const float* mat, float* eigenvalue, float* vec)
{
  float scale = 0.0f;
  for (int i = 0; i < 9; ++i) scale = fmaxf(scale, fabsf(mat[i]));
  if (scale <= 1.17549435e-38f) scale = 1.0f;
  float sm[9];
  for (int i = 0; i < 9; ++i) sm[i] = mat[i] / scale;
  float ev[3];
  compute_roots_(sm, ev);
  *eigenvalue = ev[0] * scale;
  sm[0] -= ev[0];
  sm[4] -= ev[0];
  sm[8] -= ev[0];
  const float* r0 = sm;
  const float* r1 = sm + 3;
  const float* r2 = sm + 6;
  float v1[3] = {r0[1] * r1[2] - r0[2] * r1[1], r0[2] * r1[0] - r0[0] * r1[2], r0[0] * r1[1] - r0[1] * r1[0]};
  float v2[3] = {r0[1] * r2[2] - r0[2] * r2[1], r0[2] * r2[0] - r0[0] * r2[2], r0[0] * r2[1] - r0[1] * r2[0]};
  float v3[3] = {r1[1] * r2[2] - r1[2] * r2[1], r1[2] * r2[0] - r1[0] * r2[2], r1[0] * r2[1] - r1[1] * r2[0]};
  const float l1 = v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2];
  const float l2 = v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2];
  const float l3 = v3[0] * v3[0] + v3[1] * v3[1] + v3[2] * v3[2];
  const float* v;
  float l;
  if (l1 >= l2 && l1 >= l3) {
    v = v1;
    l = l1;
  } else if (l2 >= l1 && l2 >= l3) {
    v = v2;
    l = l2;
  } else {
    v = v3;
    l = l3;
  }
  const float inv = sqrtf(l);
  vec[0] = v[0] / inv;
  vec[1] = v[1] / inv;
  vec[2] = v[2] / inv;
}

// Symmetric 3x3 eigen-decomposition in double by cyclic Jacobi rotations (stands in for
// Eigen::SelfAdjointEigenSolver<Matrix3d> in the SHOT local reference frame).  a is row-major symmetric;
// eigenvalues come out ascending, vec[:, i] (column i, row-major storage) is the matching unit eigenvector.
MM_HD void eig3_sym_d(const double* a_in, double* val, double* vec)
{
  double a[9];
  for (int i = 0; i < 9; ++i) {
    a[i] = a_in[i];
    vec[i] = 0.0;
  }
  vec[0] = vec[4] = vec[8] = 1.0;
  for (int sweep = 0; sweep < 50; ++sweep) {
    const double off = a[1] * a[1] + a[2] * a[2] + a[5] * a[5];
    const double diag = a[0] * a[0] + a[4] * a[4] + a[8] * a[8];
    if (off <= 1e-32 * diag || off == 0.0) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        const double apq = a[p * 3 + q];
        if (apq == 0.0) continue;
        const double theta = (a[q * 3 + q] - a[p * 3 + p]) / (2.0 * apq);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
        for (int k = 0; k < 3; ++k) {  // A <- A J
          const double akp = a[k * 3 + p], akq = a[k * 3 + q];
          a[k * 3 + p] = c * akp - sn * akq;
          a[k * 3 + q] = sn * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) {  // A <- J^T A
          const double apk = a[p * 3 + k], aqk = a[q * 3 + k];
          a[p * 3 + k] = c * apk - sn * aqk;
          a[q * 3 + k] = sn * apk + c * aqk;
        }
        for (int k = 0; k < 3; ++k) {  // V <- V J
          const double vkp = vec[k * 3 + p], vkq = vec[k * 3 + q];
          vec[k * 3 + p] = c * vkp - sn * vkq;
          vec[k * 3 + q] = sn * vkp + c * vkq;
        }
      }
  }
  val[0] = a[0];
  val[1] = a[4];
  val[2] = a[8];
  // ascending order (selection sort on three values, columns follow)
  for (int i = 0; i < 2; ++i)
    for (int j = i + 1; j < 3; ++j)
      if (val[j] < val[i]) {
        const double tv = val[i];
        val[i] = val[j];
        val[j] = tv;
        for (int k = 0; k < 3; ++k) {
          const double tt = vec[k * 3 + i];
          vec[k * 3 + i] = vec[k * 3 + j];
          vec[k * 3 + j] = tt;
        }
      }
}

// acos / atan2 in double for SHOT's inclination and azimuth interpolation (IEEE ops only).
MM_HD double atan01_d_(double z)
{
  z = z / (1.0 + sqrt(1.0 + z * z));
  z = z / (1.0 + sqrt(1.0 + z * z));
  const double z2 = z * z;
  double s = 1.0 / 23.0;
  s = 1.0 / 21.0 - z2 * s;
  s = 1.0 / 19.0 - z2 * s;
  s = 1.0 / 17.0 - z2 * s;
  s = 1.0 / 15.0 - z2 * s;
  s = 1.0 / 13.0 - z2 * s;
  s = 1.0 / 11.0 - z2 * s;
  s = 1.0 / 9.0 - z2 * s;
  s = 1.0 / 7.0 - z2 * s;
  s = 1.0 / 5.0 - z2 * s;
  s = 1.0 / 3.0 - z2 * s;
  s = 1.0 - z2 * s;
  return 4.0 * (z * s);
}
MM_HD double atan2_d_(double y, double x)
{
  const double kPi = 3.14159265358979323846;
  if (x != x || y != y) return x + y;
  const double ax = fabs(x), ay = fabs(y);
  double a;
  if (ax == 0.0 && ay == 0.0) a = 0.0;
  else if (ay <= ax) a = atan01_d_(ay / ax);
  else a = 0.5 * kPi - atan01_d_(ax / ay);
  if (signbit(x)) a = kPi - a;
  return signbit(y) ? -a : a;
}
// acos(c) for c in [-1, 1]: atan2(sqrt((1 - c)(1 + c)), c)
MM_HD double acos_d_(double c)
{
  return atan2_d_(sqrt((1.0 - c) * (1.0 + c)), c);
}

// FLANN L2_Simple in 3-D: sequential float accumulation, no contraction.
MM_HD float dist2_3(float ax, float ay, float az, float bx, float by, float bz)
{
  const float dx = ax - bx, dy = ay - by, dz = az - bz;
  float r = dx * dx;
  r += dy * dy;
  r += dz * dz;
  return r;
}

// row-major 4x4 applied to a point the way Eigen/PCL evaluate it:
// ((m0*x + m1*y) + m2*z) + m3
MM_HD void transform_point(const float* m, float x, float y, float z, float* ox, float* oy, float* oz)
{
  *ox = ((m[0] * x + m[1] * y) + m[2] * z) + m[3];
  *oy = ((m[4] * x + m[5] * y) + m[6] * z) + m[7];
  *oz = ((m[8] * x + m[9] * y) + m[10] * z) + m[11];
}

// Order-independent accumulation: every term is rounded to a fixed-point grid
// and summed in int64, so any summation order gives the same bits.
#define MM3D_FIX1_SCALE 4294967296.0  /* 2^32: first-order sums (metres)   */
#define MM3D_FIX2_SCALE 16777216.0    /* 2^24: second-order sums (metres^2) */
#define MM3D_FIXD_SCALE 68719476736.0 /* 2^36: squared NN distances         */
MM_HD long long to_fix(double v, double scale)
{
#if defined(__CUDA_ARCH__)
  return __double2ll_rn(v * scale);
#else
  return (long long)rint(v * scale);
#endif
}

// rint(v * 2^32) for a finite v >= 0 below 2^31, in integer arithmetic (no FP64): the float's 24-bit significand shifted
// into place, round-to-nearest-even when bits fall off.  Same value as to_fix((double)v, MM3D_FIX1_SCALE).
MM_HD long long to_fix32_pos(float v)
{
  unsigned int b;
#if defined(__CUDA_ARCH__)
  b = __float_as_uint(v);
#else
  memcpy(&b, &v, 4);
#endif
  const int e = (int)((b >> 23) & 0xffu);
  const unsigned long long m = (unsigned long long)(b & 0x7fffffu) | (e ? 0x800000ull : 0ull);
  const int sh = (e ? e : 1) - 118;  // (e - 127 - 23) + 32
  if (sh >= 0) return (long long)(m << sh);
  const int s = -sh;
  if (s > 25) return 0;
  const unsigned long long r = m >> s, rem = m & ((1ull << s) - 1ull), half = 1ull << (s - 1);
  return (long long)(r + ((rem > half || (rem == half && (r & 1ull))) ? 1ull : 0ull));
}

}  // namespace em
}  // namespace mm3d
