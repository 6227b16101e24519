// pair_features.cuh — Darboux pair features and histogram binning shared by the FPFH and PFH kernels.
#pragma once
#include <cmath>
#include <limits>

#include "mm3d_internal.cuh"

namespace mm3d {

// t[feature][b] = smallest float whose (double-evaluated) bin index is >= b; t[.][0] = -inf, unused slots = +inf
struct BinTable {
  float t[3][12];
};

#if defined(__CUDACC__)
// pcl::computePairFeatures [PCL-recall pcl/features/impl/pfh.hpp]; false = degenerate pair (coincident points, or the
// normal parallel to the connecting line), which the callers skip like computePointSPFHSignature / computePointPFHSignature.
__device__ __forceinline__ bool pair_features(const float4& p1, const float4& n1, const float4& p2, const float4& n2, float* f1, float* f2,
                                              float* f3)
{
  float dx = p2.x - p1.x, dy = p2.y - p1.y, dz = p2.z - p1.z;
  const float f4 = sqrtf((dx * dx + dy * dy) + dz * dz);
  if (f4 == 0.0f) { *f1 = *f2 = *f3 = 0.0f; return false; }
  float ax = n1.x, ay = n1.y, az = n1.z, bx = n2.x, by = n2.y, bz = n2.z;
  const float angle1 = ((ax * dx + ay * dy) + az * dz) / f4;
  const float angle2 = ((bx * dx + by * dy) + bz * dz) / f4;
  const float fa1 = fabsf(angle1), fa2 = fabsf(angle2);
  // acos(|a1|) > acos(|a2|)  <=>  |a1| < |a2| with both inside [0, 1] (NaN otherwise)
  if ((fa1 <= 1.0f) && (fa2 <= 1.0f) && (fa1 < fa2)) {
    float t;
    t = ax; ax = bx; bx = t;
    t = ay; ay = by; by = t;
    t = az; az = bz; bz = t;
    dx *= -1.f; dy *= -1.f; dz *= -1.f;
    *f3 = -angle2;
  } else {
    *f3 = angle1;
  }
  float vx = dy * az - dz * ay, vy = dz * ax - dx * az, vz = dx * ay - dy * ax;
  const float v_norm = sqrtf((vx * vx + vy * vy) + vz * vz);
  if (v_norm == 0.0f) { *f1 = *f2 = *f3 = 0.0f; return false; }
  vx /= v_norm; vy /= v_norm; vz /= v_norm;
  const float wx = ay * vz - az * vy, wy = az * vx - ax * vz, wz = ax * vy - ay * vx;
  *f2 = (vx * bx + vy * by) + vz * bz;
  *f1 = em::atan2f_((wx * bx + wy * by) + wz * bz, (ax * bx + ay * by) + az * bz);
  return true;
}

// pcl::computeRGBPairFeatures [PCL-recall pcl/features/impl/pfhrgb.hpp]: the same Darboux frame WITHOUT the source/target
// swap; the colour ratios are handled by the caller (integer division, see pfh.cu).
__device__ __forceinline__ bool pair_features_noswap(const float4& p1, const float4& n1, const float4& p2, const float4& n2, float* f1,
                                                     float* f2, float* f3)
{
  const float dx = p2.x - p1.x, dy = p2.y - p1.y, dz = p2.z - p1.z;
  const float f4 = sqrtf((dx * dx + dy * dy) + dz * dz);
  if (f4 == 0.0f) return false;
  const float ax = n1.x, ay = n1.y, az = n1.z, bx = n2.x, by = n2.y, bz = n2.z;
  *f3 = ((ax * dx + ay * dy) + az * dz) / f4;
  float vx = dy * az - dz * ay, vy = dz * ax - dx * az, vz = dx * ay - dy * ax;
  const float v_norm = sqrtf((vx * vx + vy * vy) + vz * vz);
  if (v_norm == 0.0f) return false;
  vx /= v_norm; vy /= v_norm; vz /= v_norm;
  const float wx = ay * vz - az * vy, wy = az * vx - ax * vz, wz = ax * vy - ay * vx;
  *f2 = (vx * bx + vy * by) + vz * bz;
  *f1 = em::atan2f_((wx * bx + wy * by) + wz * bz, (ax * bx + ay * by) + az * bz);
  return true;
}


// FPFH bins are floor(11 * ((f + pi) / 2pi)) resp. floor(11 * ((f + 1) / 2)) evaluated in DOUBLE from a float feature
// (pcl/features/impl/fpfh.hpp).  Both are monotone in f, so the bin is fully described by 10 float thresholds:
// thr[b] = smallest float whose double formula reaches bin b.  The host derives the thresholds from the literal double
// expressions; the kernel only compares floats.

__device__ __forceinline__ int lookup_bin(const float* thr, int nb, float f, float scale, float shift)
{
  int g = (int)((f + shift) * scale);  // float guess, at most one bin off
  g = g < 0 ? 0 : (g > nb - 1 ? nb - 1 : g);
  while (g < nb - 1 && f >= thr[g + 1]) ++g;
  while (g > 0 && f < thr[g]) --g;
  return g;
}

#endif  // __CUDACC__

// thresholds derived on the host from the literal double expressions of pcl/features/impl/{fpfh,pfh}.hpp
inline int pf_bin_f1(float f, int nb)
{
  const float d_pi = 1.0f / (2.0f * (float)M_PI);
  int h = (int)std::floor(nb * (((double)f + M_PI) * (double)d_pi));
  return h < 0 ? 0 : (h > nb - 1 ? nb - 1 : h);
}
inline int pf_bin_f23(float f, int nb)
{
  int h = (int)std::floor(nb * (((double)f + 1.0) * 0.5));
  return h < 0 ? 0 : (h > nb - 1 ? nb - 1 : h);
}

inline BinTable make_bin_table(int nb)
{
  BinTable bt;
  auto next_up = [](float f) { return std::nextafter(f, std::numeric_limits<float>::infinity()); };
  for (int feat = 0; feat < 3; ++feat) {
    for (int b = 0; b < 12; ++b) bt.t[feat][b] = std::numeric_limits<float>::infinity();
    bt.t[feat][0] = -std::numeric_limits<float>::infinity();
    for (int b = 1; b <= nb - 1; ++b) {
      // smallest float whose bin is >= b: bisection over the ordered float line in [-8, 8]
      float lo = -8.0f, hi = 8.0f;
      while (next_up(lo) < hi) {
        float m = lo + (hi - lo) * 0.5f;
        if (m <= lo) m = next_up(lo);
        if (m >= hi) break;
        const int v = feat == 0 ? pf_bin_f1(m, nb) : pf_bin_f23(m, nb);
        if (v >= b) hi = m;
        else lo = m;
      }
      bt.t[feat][b] = hi;
    }
  }
  return bt;
}

}  // namespace mm3d
