// shot.cu — K8 SHOT-colour descriptors (1344 = 32 sectors x (11 shape + 31 colour bins)).
//   <- pcl::SHOTColorEstimation via map_merge_3d/src/dispatch_descriptors.h:46 and src/features.cpp:99-150
//   [PCL-recall pcl/features/impl/shot.hpp, pcl/features/impl/shot_lrf.hpp]
// One thread owns one keypoint: the local reference frame (weighted covariance in double, Jacobi eigen-solve,
// sign disambiguation) takes up to three neighbourhood walks, the quadrilinear vote accumulation one more.
// Votes land in the keypoint's private 1344-float row in ascending neighbour index, i.e. in the order the
// CPU checker adds them, so the histogram is bit-identical.
#include <algorithm>
#include <cmath>

#include "mm3d_internal.cuh"

namespace mm3d {

namespace {

constexpr int SB = 64;
constexpr int SHOT_D = 1344;

struct ShotJob {
  GridView g;
  const float4* normals;
  const float4* kp;
  int nk;
  float* desc_raw;  // nk x 1344
  float* rf;        // nk x 9
  uint32_t* valid;  // nk
};

struct LabLutDev {
  const float* srgb;  // 256
  const float* xyz;   // 4000
};

__device__ __forceinline__ void rgb2cielab(const LabLutDev& lut, uint32_t rgba, float* L, float* A, float* B2)
{
  const float fr = lut.srgb[(rgba >> 16) & 0xff], fg = lut.srgb[(rgba >> 8) & 0xff], fb = lut.srgb[rgba & 0xff];
  const float x = fr * 0.412453f + fg * 0.357580f + fb * 0.180423f;
  const float y = fr * 0.212671f + fg * 0.715160f + fb * 0.072169f;
  const float z = fr * 0.019334f + fg * 0.119193f + fb * 0.950227f;
  float vx = x / 0.95047f;
  float vy = y;
  float vz = z / 1.08883f;
  vx = lut.xyz[(int)(vx * 4000)];
  vy = lut.xyz[(int)(vy * 4000)];
  vz = lut.xyz[(int)(vz * 4000)];
  float l = 116.0f * vy - 16.0f;
  if (l > 100) l = 100.0f;
  float a = 500.0f * (vx - vy);
  if (a > 120) a = 120.0f;
  else if (a < -120) a = -120.0f;
  float b = 200.0f * (vy - vz);
  if (b > 120) b = 120.0f;
  else if (b < -120) b = -120.0f;
  *L = l;
  *A = a;
  *B2 = b;
}

__global__ void __launch_bounds__(SB) shot_kernel(const ShotJob* __restrict__ jobs, double radius, float r2, int rv, LabLutDev lut)
{
  const ShotJob& j = jobs[blockIdx.y];
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = t < j.nk;
  const float4 c = live ? j.kp[t] : make_float4(0.f, 0.f, 0.f, 0.f);

  // ---- SHOTLocalReferenceFrameEstimation::getLocalRF, pass 1: weighted covariance
  double cov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  double sum = 0.0;
  int valid = 0, n_nb = 0;
  for_each_in_radius(j.g, live, c.x, c.y, c.z, r2, rv, [&](int, const float4& pt, float d2) {
    ++n_nb;
    if (pt.x == c.x && pt.y == c.y && pt.z == c.z) return;
    const double v[3] = {(double)(pt.x - c.x), (double)(pt.y - c.y), (double)(pt.z - c.z)};
    const double distance = radius - sqrt((double)d2);
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) cov[r * 3 + cc] += distance * (v[r] * v[cc]);
    sum += distance;
    ++valid;
  });
  bool ok = live && valid >= 5;
  double v1[3] = {0, 0, 0}, v3[3] = {0, 0, 0};
  if (ok) {
    for (int k = 0; k < 9; ++k) cov[k] /= sum;
    double val[3], vec[9];
    em::eig3_sym_d(cov, val, vec);
    if (!isfinite(val[0]) || !isfinite(val[1]) || !isfinite(val[2])) ok = false;
    for (int k = 0; k < 3; ++k) {
      v1[k] = vec[k * 3 + 2];
      v3[k] = vec[k * 3 + 0];
    }
  }
  // pass 2: sign votes
  int plusNormal = 0, plusTangent = 0;
  for_each_in_radius(j.g, ok, c.x, c.y, c.z, r2, rv, [&](int, const float4& pt, float) {
    if (pt.x == c.x && pt.y == c.y && pt.z == c.z) return;
    const double v[3] = {(double)(pt.x - c.x), (double)(pt.y - c.y), (double)(pt.z - c.z)};
    if ((v[0] * v1[0] + v[1] * v1[1]) + v[2] * v1[2] >= 0) ++plusTangent;
    if ((v[0] * v3[0] + v[1] * v3[1]) + v[2] * v3[2] >= 0) ++plusNormal;
  });
  plusTangent = 2 * plusTangent - valid;
  plusNormal = 2 * plusNormal - valid;
  // pass 3 (ties only): the five rows around the median decide
  const bool tie = ok && (plusTangent == 0 || plusNormal == 0);
  int tieT = 0, tieN = 0, ord = 0;
  const int median = valid / 2;
  for_each_in_radius(j.g, tie, c.x, c.y, c.z, r2, rv, [&](int, const float4& pt, float) {
    if (pt.x == c.x && pt.y == c.y && pt.z == c.z) return;
    if (ord >= median - 2 && ord <= median + 2) {
      const double v[3] = {(double)(pt.x - c.x), (double)(pt.y - c.y), (double)(pt.z - c.z)};
      if ((v[0] * v1[0] + v[1] * v1[1]) + v[2] * v1[2] > 0) ++tieT;
      if ((v[0] * v3[0] + v[1] * v3[1]) + v[2] * v3[2] > 0) ++tieN;
    }
    ++ord;
  });
  if (ok) {
    if (plusTangent == 0) {
      if (tieT < 3) { v1[0] *= -1; v1[1] *= -1; v1[2] *= -1; }
    } else if (plusTangent < 0) {
      v1[0] *= -1; v1[1] *= -1; v1[2] *= -1;
    }
    if (plusNormal == 0) {
      if (tieN < 3) { v3[0] *= -1; v3[1] *= -1; v3[2] *= -1; }
    } else if (plusNormal < 0) {
      v3[0] *= -1; v3[1] *= -1; v3[2] *= -1;
    }
  }
  float rf[9];
  for (int k = 0; k < 3; ++k) {
    rf[k] = (float)v1[k];
    rf[6 + k] = (float)v3[k];
  }
  rf[3] = rf[7] * rf[2] - rf[8] * rf[1];
  rf[4] = rf[8] * rf[0] - rf[6] * rf[2];
  rf[5] = rf[6] * rf[1] - rf[7] * rf[0];
  ok = ok && n_nb >= 5;

  // ---- computePointSHOT + interpolateDoubleChannel
  float* shot = j.desc_raw + (size_t)(live ? t : 0) * SHOT_D;
  if (ok)
    for (int k = 0; k < SHOT_D; ++k) shot[k] = 0.f;
  const int nr_shape = 10, nr_color = 30, sectors = 32;
  const int shapeToColorStride = sectors * (nr_shape + 1);
  const double radius3_4 = (radius * 3) / 4, radius1_4 = radius / 4, radius1_2 = radius / 2;
  const double RAD_45 = 0.78539816339744830961566084581988, RAD_90 = 1.5707963267948966192313216916398,
               RAD_135 = 2.3561944901923449288469825374596, RAD_PI_7_8 = 2.7488935718910690836548129603691;
  float LRef = 0.f, aRef = 0.f, bRef = 0.f;
  if (ok) {
    rgb2cielab(lut, __float_as_uint(c.w), &LRef, &aRef, &bRef);
    LRef /= 100.0f; aRef /= 120.0f; bRef /= 120.0f;
  }
  for_each_in_radius(j.g, ok, c.x, c.y, c.z, r2, rv, [&](int s, const float4& sp, float d2) {
    const float4 nq = j.normals[j.g.orig ? j.g.orig[s] : s];
    if (!isfinite(nq.x) || !isfinite(nq.y) || !isfinite(nq.z)) return;
    double cosineDesc = (double)(((nq.x * rf[6] + nq.y * rf[7]) + nq.z * rf[8]) + 0.0f);
    if (cosineDesc > 1.0) cosineDesc = 1.0;
    if (cosineDesc < -1.0) cosineDesc = -1.0;
    double binDistanceShape = ((1.0 + cosineDesc) * nr_shape) / 2;
    float L, a, b;
    rgb2cielab(lut, __float_as_uint(sp.w), &L, &a, &b);
    L /= 100.0f; a /= 120.0f; b /= 120.0f;
    double colorDistance = (double)((fabsf(LRef - L) + ((fabsf(aRef - a) + fabsf(bRef - b)) / 2)) / 3);
    if (colorDistance > 1.0) colorDistance = 1.0;
    if (colorDistance < 0.0) colorDistance = 0.0;
    double binDistanceColor = colorDistance * nr_color;
    const float dl[3] = {sp.x - c.x, sp.y - c.y, sp.z - c.z};
    const double distance = sqrt((double)d2);
    if (fabs(distance - 0.0) < 1E-15) return;
    double xInFeatRef = (double)((dl[0] * rf[0] + dl[1] * rf[1]) + dl[2] * rf[2]);
    double yInFeatRef = (double)((dl[0] * rf[3] + dl[1] * rf[4]) + dl[2] * rf[5]);
    double zInFeatRef = (double)((dl[0] * rf[6] + dl[1] * rf[7]) + dl[2] * rf[8]);
    if (fabs(yInFeatRef) < 1E-30) yInFeatRef = 0;
    if (fabs(xInFeatRef) < 1E-30) xInFeatRef = 0;
    if (fabs(zInFeatRef) < 1E-30) zInFeatRef = 0;
    const int bit4 = ((yInFeatRef > 0) || ((yInFeatRef == 0.0) && (xInFeatRef < 0))) ? 1 : 0;
    const int bit3 = ((xInFeatRef > 0) || ((xInFeatRef == 0.0) && (yInFeatRef > 0))) ? !bit4 : bit4;
    int desc_index = (bit4 << 3) + (bit3 << 2);
    desc_index = desc_index << 1;
    if ((xInFeatRef * yInFeatRef > 0) || (xInFeatRef == 0.0)) desc_index += (fabs(xInFeatRef) >= fabs(yInFeatRef)) ? 0 : 4;
    else desc_index += (fabs(xInFeatRef) > fabs(yInFeatRef)) ? 4 : 0;
    desc_index += zInFeatRef > 0 ? 1 : 0;
    desc_index += (distance > radius1_2) ? 2 : 0;
    const int step_index_shape = (int)floor(binDistanceShape + 0.5);
    const int step_index_color = (int)floor(binDistanceColor + 0.5);
    const int volume_index_shape = desc_index * (nr_shape + 1);
    const int volume_index_color = shapeToColorStride + desc_index * (nr_color + 1);
    binDistanceShape -= step_index_shape;
    binDistanceColor -= step_index_color;
    double intWeightShape = (1 - fabs(binDistanceShape));
    double intWeightColor = (1 - fabs(binDistanceColor));
    if (binDistanceShape > 0) shot[volume_index_shape + ((step_index_shape + 1) % nr_shape)] += (float)binDistanceShape;
    else shot[volume_index_shape + ((step_index_shape - 1 + nr_shape) % nr_shape)] -= (float)binDistanceShape;
    if (binDistanceColor > 0) shot[volume_index_color + ((step_index_color + 1) % nr_color)] += (float)binDistanceColor;
    else shot[volume_index_color + ((step_index_color - 1 + nr_color) % nr_color)] -= (float)binDistanceColor;
    if (distance > radius1_2) {
      const double radiusDistance = (distance - radius3_4) / radius1_2;
      if (distance > radius3_4) {
        intWeightShape += 1 - radiusDistance;
        intWeightColor += 1 - radiusDistance;
      } else {
        intWeightShape += 1 + radiusDistance;
        intWeightColor += 1 + radiusDistance;
        shot[(desc_index - 2) * (nr_shape + 1) + step_index_shape] -= (float)radiusDistance;
        shot[shapeToColorStride + (desc_index - 2) * (nr_color + 1) + step_index_color] -= (float)radiusDistance;
      }
    } else {
      const double radiusDistance = (distance - radius1_4) / radius1_2;
      if (distance < radius1_4) {
        intWeightShape += 1 + radiusDistance;
        intWeightColor += 1 + radiusDistance;
      } else {
        intWeightShape += 1 - radiusDistance;
        intWeightColor += 1 - radiusDistance;
        shot[(desc_index + 2) * (nr_shape + 1) + step_index_shape] += (float)radiusDistance;
        shot[shapeToColorStride + (desc_index + 2) * (nr_color + 1) + step_index_color] += (float)radiusDistance;
      }
    }
    double inclinationCosine = zInFeatRef / distance;
    if (inclinationCosine < -1.0) inclinationCosine = -1.0;
    if (inclinationCosine > 1.0) inclinationCosine = 1.0;
    const double inclination = em::acos_d_(inclinationCosine);
    if (inclination > RAD_90 || (fabs(inclination - RAD_90) < 1e-30 && zInFeatRef <= 0)) {
      const double inclinationDistance = (inclination - RAD_135) / RAD_90;
      if (inclination > RAD_135) {
        intWeightShape += 1 - inclinationDistance;
        intWeightColor += 1 - inclinationDistance;
      } else {
        intWeightShape += 1 + inclinationDistance;
        intWeightColor += 1 + inclinationDistance;
        shot[(desc_index + 1) * (nr_shape + 1) + step_index_shape] -= (float)inclinationDistance;
        shot[shapeToColorStride + (desc_index + 1) * (nr_color + 1) + step_index_color] -= (float)inclinationDistance;
      }
    } else {
      const double inclinationDistance = (inclination - RAD_45) / RAD_90;
      if (inclination < RAD_45) {
        intWeightShape += 1 + inclinationDistance;
        intWeightColor += 1 + inclinationDistance;
      } else {
        intWeightShape += 1 - inclinationDistance;
        intWeightColor += 1 - inclinationDistance;
        shot[(desc_index - 1) * (nr_shape + 1) + step_index_shape] += (float)inclinationDistance;
        shot[shapeToColorStride + (desc_index - 1) * (nr_color + 1) + step_index_color] += (float)inclinationDistance;
      }
    }
    if (yInFeatRef != 0.0 || xInFeatRef != 0.0) {
      const double azimuth = em::atan2_d_(yInFeatRef, xInFeatRef);
      const int sel = desc_index >> 2;
      const double angularSectorSpan = RAD_45;
      const double angularSectorStart = -RAD_PI_7_8;
      double azimuthDistance = (azimuth - (angularSectorStart + angularSectorSpan * sel)) / angularSectorSpan;
      azimuthDistance = fmax(-0.5, fmin(azimuthDistance, 0.5));
      if (azimuthDistance > 0) {
        intWeightShape += 1 - azimuthDistance;
        intWeightColor += 1 - azimuthDistance;
        const int interp_index = (desc_index + 4) % sectors;
        shot[interp_index * (nr_shape + 1) + step_index_shape] += (float)azimuthDistance;
        shot[shapeToColorStride + interp_index * (nr_color + 1) + step_index_color] += (float)azimuthDistance;
      } else {
        const int interp_index = (desc_index - 4 + sectors) % sectors;
        intWeightShape += 1 + azimuthDistance;
        intWeightColor += 1 + azimuthDistance;
        shot[interp_index * (nr_shape + 1) + step_index_shape] -= (float)azimuthDistance;
        shot[shapeToColorStride + interp_index * (nr_color + 1) + step_index_color] -= (float)azimuthDistance;
      }
    }
    shot[volume_index_shape + step_index_shape] += (float)intWeightShape;
    shot[volume_index_color + step_index_color] += (float)intWeightColor;
  });
  if (!live) return;
  if (ok) {
    // normalizeHistogram
    double acc_norm = 0.0;
    for (int k = 0; k < SHOT_D; ++k) acc_norm += shot[k] * shot[k];
    acc_norm = sqrt(acc_norm);
    for (int k = 0; k < SHOT_D; ++k) {
      const float v = shot[k] / (float)acc_norm;
      shot[k] = v;
      if (!isfinite(v)) ok = false;
    }
  }
  j.valid[t] = ok ? 1u : 0u;
  for (int k = 0; k < 9; ++k) j.rf[(size_t)t * 9 + k] = rf[k];
}

struct ShotEmitJob {
  const float4* kp;
  const float* desc_raw;
  const float* rf_raw;
  const uint32_t* flags;
  const uint32_t* pos;
  float4* kp_out;
  float* desc_out;
  float* rf_out;
  int nk;
};
__global__ void __launch_bounds__(256) shot_emit_kernel(const ShotEmitJob* __restrict__ jobs)
{
  const ShotEmitJob& j = jobs[blockIdx.y];
  const int kp = blockIdx.x;
  if (kp >= j.nk || !j.flags[kp]) return;
  const uint32_t o = j.pos[kp];
  for (int b = threadIdx.x; b < SHOT_D; b += blockDim.x) j.desc_out[(size_t)o * SHOT_D + b] = j.desc_raw[(size_t)kp * SHOT_D + b];
  if (threadIdx.x == 0) j.kp_out[o] = j.kp[kp];
  if (threadIdx.x < 9 && j.rf_out) j.rf_out[(size_t)o * 9 + threadIdx.x] = j.rf_raw[(size_t)kp * 9 + threadIdx.x];
}

struct LabLutHost {
  float srgb[256];
  float xyz[4000];
  LabLutHost()
  {
    for (int i = 0; i < 256; i++) {
      const float f = (float)i / 255.0f;
      if (f > 0.04045) srgb[i] = powf((f + 0.055f) / 1.055f, 2.4f);
      else srgb[i] = f / 12.92f;
    }
    for (int i = 0; i < 4000; i++) {
      const float f = (float)i / 4000.0f;
      if (f > 0.008856) xyz[i] = (float)powf(f, 0.3333f);
      else xyz[i] = (float)((7.787 * f) + (16.0 / 116.0));
    }
  }
};

}  // namespace

void shot_batch(Ctx& c, const std::vector<CloudView>& clouds, const std::vector<DIndex>& idx, const std::vector<const float4*>& normals,
                std::vector<DCloud>& keypoints, double radius, std::vector<DBuf<float>>& desc, std::vector<DBuf<float>>* rf_dbg)
{
  const int M = (int)clouds.size();
  desc.clear();
  desc.resize(M);
  if (rf_dbg) { rf_dbg->clear(); rf_dbg->resize(M); }
  if (M == 0) return;
  std::vector<int> nks(M);
  int mxk = 0, totalk = 0;
  std::vector<Seg> segk(M);
  for (int m = 0; m < M; ++m) {
    nks[m] = keypoints[m].n;
    segk[m].off = totalk;
    segk[m].n = nks[m];
    totalk += nks[m];
    mxk = std::max(mxk, nks[m]);
  }
  if (totalk == 0) {
    for (int m = 0; m < M; ++m) { keypoints[m].n = 0; keypoints[m].pts.release(); }
    return;
  }
  static const LabLutHost hl;
  DBuf<float> dl(c, 4256);
  dl.upload(c, hl.srgb, 256);
  MM_CUDA(cudaMemcpyAsync(dl.p + 256, hl.xyz, 4000 * sizeof(float), cudaMemcpyHostToDevice, c.stream));
  LabLutDev lut{dl.p, dl.p + 256};
  DBuf<uint32_t> flags(c, totalk), pos(c, totalk);
  std::vector<DBuf<float>> raw(M), rfr(M);
  std::vector<ShotJob> jobs(M);
  for (int m = 0; m < M; ++m) {
    raw[m].alloc(c, (size_t)nks[m] * SHOT_D);
    rfr[m].alloc(c, (size_t)nks[m] * 9);
    jobs[m] = ShotJob{idx[m].v, normals[m], keypoints[m].pts.p, nks[m], raw[m].p, rfr[m].p, flags.p + segk[m].off};
  }
  DBuf<ShotJob> dj = to_device(c, jobs);
  const float r2 = (float)(radius * radius);
  const int rv = (int)std::ceil(radius / (double)idx[0].v.leaf) + 1;
  { double b = 0; for (int m = 0; m < M; ++m) b += 32.0 * clouds[m].n + (16.0 + 4.0 * SHOT_D) * nks[m]; MM_BYTES(c, b); }
  MM_LAUNCH(c, shot_kernel, dim3((mxk + SB - 1) / SB, M), SB, 0, dj.p, radius, r2, rv, lut);
  std::vector<int> totals;
  scan_flags_batch(c, flags.p, pos.p, segk, totals);
  std::vector<DCloud> kept(M);
  std::vector<ShotEmitJob> ej(M);
  for (int m = 0; m < M; ++m) {
    kept[m].n = totals[m];
    kept[m].pts.alloc(c, totals[m]);
    desc[m].alloc(c, (size_t)totals[m] * SHOT_D);
    if (rf_dbg) (*rf_dbg)[m].alloc(c, (size_t)totals[m] * 9);
    ej[m] = ShotEmitJob{keypoints[m].pts.p, raw[m].p, rfr[m].p, flags.p + segk[m].off, pos.p + segk[m].off, kept[m].pts.p, desc[m].p,
                        rf_dbg ? (*rf_dbg)[m].p : nullptr, nks[m]};
  }
  DBuf<ShotEmitJob> dej = to_device(c, ej);
  MM_LAUNCH(c, shot_emit_kernel, dim3(mxk, M), 256, 0, dej.p);
  for (int m = 0; m < M; ++m) keypoints[m] = std::move(kept[m]);
}

}  // namespace mm3d
