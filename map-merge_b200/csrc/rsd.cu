// rsd.cu — RSD descriptors (r_min, r_max)
//   <- pcl::RSDEstimation<PointXYZRGB, Normal, PrincipalRadiiRSD> via map_merge_3d/src/dispatch_descriptors.h:43,
//      src/features.cpp:99-150   [PCL-recall pcl/features/impl/rsd.hpp computeRSD]
// Warp per keypoint.  Pass 1: the nearest surface point (smallest (d^2, index)) is the centre of the patch.  Pass 2: for
// every other neighbour the angle between the normals and the distance to the centre update per-distance-bin min / max
// angles — min and max are order-free, so the lanes take one neighbour each and a shuffle reduction finishes.  The two
// least-squares radii are then computed by lane 0 exactly as written in PCL (double arithmetic).
#include <algorithm>
#include <cfloat>
#include <cmath>

#include "mm3d_internal.cuh"

namespace mm3d {

namespace {

constexpr int RB = 256;
constexpr int NR_SUBDIV = 5;

struct RsdJob {
  GridView g;
  const float4* normals;
  const float4* kp;
  int nk;
  float* desc;      // nk x 2
  uint32_t* valid;  // nk
};

__global__ void __launch_bounds__(RB) rsd_kernel(const RsdJob* __restrict__ jobs, float r2, int rv, double max_dist, double plane_radius)
{
  const RsdJob& j = jobs[blockIdx.y];
  const int lane = threadIdx.x & 31;
  const int k = blockIdx.x * (RB / 32) + (threadIdx.x >> 5);
  if (k >= j.nk) return;  // warp-uniform
  const unsigned full = 0xffffffffu;
  const float4 c = j.kp[k];
  // pass 1: count + nearest (d^2, original index)
  int n = 0;
  float bd = FLT_MAX;
  int bi = 0x7fffffff, bslot = -1;
  warp_radius_unordered(j.g, c.x, c.y, c.z, r2, rv, [&](bool valid, int s, const float4&, float d2) {
    if (valid) {
      ++n;
      const int oi = j.g.orig ? j.g.orig[s] : s;
      if (d2 < bd || (d2 == bd && oi < bi)) { bd = d2; bi = oi; bslot = s; }
    }
  });
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    n += __shfl_xor_sync(full, n, o);
    const float od = __shfl_xor_sync(full, bd, o);
    const int oi = __shfl_xor_sync(full, bi, o), os = __shfl_xor_sync(full, bslot, o);
    if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; bslot = os; }
  }
  float r_min = 0.0f, r_max = 0.0f;
  if (n >= 2) {
    const float4 pc = j.g.pts[bslot];
    const float4 nc = j.normals[bi];
    double mn[NR_SUBDIV], mx[NR_SUBDIV];
#pragma unroll
    for (int d = 0; d < NR_SUBDIV; ++d) { mn[d] = DBL_MAX; mx[d] = -DBL_MAX; }
    warp_radius_unordered(j.g, c.x, c.y, c.z, r2, rv, [&](bool valid, int s, const float4& p, float) {
      if (!valid || s == bslot) return;
      const float4 nq = j.normals[j.g.orig ? j.g.orig[s] : s];
      double cosine = (double)((nq.x * nc.x + nq.y * nc.y) + nq.z * nc.z);
      if (cosine > 1) cosine = 1;
      if (cosine < -1) cosine = -1;
      double angle = em::acos_d_(cosine);
      if (angle > M_PI / 2) angle = M_PI - angle;
      const float dx = p.x - pc.x, dy = p.y - pc.y, dz = p.z - pc.z;
      const double dist = sqrt((double)((dx * dx + dy * dy) + dz * dz));
      if (dist > max_dist) return;
      int bin_d = (int)floor(NR_SUBDIV * dist / max_dist);
      if (bin_d > NR_SUBDIV - 1) bin_d = NR_SUBDIV - 1;
#pragma unroll
      for (int d = 0; d < NR_SUBDIV; ++d)
        if (d == bin_d) {
          mn[d] = fmin(mn[d], angle);
          mx[d] = fmax(mx[d], angle);
        }
    });
#pragma unroll
    for (int d = 0; d < NR_SUBDIV; ++d)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        mn[d] = fmin(mn[d], __shfl_xor_sync(full, mn[d], o));
        mx[d] = fmax(mx[d], __shfl_xor_sync(full, mx[d], o));
      }
    // bin 0 starts at (0, 0) in PCL, the others at (+DBL_MAX, -DBL_MAX)
    mn[0] = fmin(mn[0], 0.0);
    mx[0] = fmax(mx[0], 0.0);
    double Amint_Amin = 0, Amint_d = 0, Amaxt_Amax = 0, Amaxt_d = 0;
#pragma unroll
    for (int d = 0; d < NR_SUBDIV; ++d) {
      if (mx[d] >= 0) {
        const double f = (d + 0.5) * max_dist / NR_SUBDIV;
        Amint_Amin += mn[d] * mn[d];
        Amint_d += mn[d] * f;
        Amaxt_Amax += mx[d] * mx[d];
        Amaxt_d += mx[d] * f;
      }
    }
    float min_radius = Amint_Amin == 0.0f ? (float)plane_radius : (float)fmin(Amint_d / Amint_Amin, plane_radius);
    float max_radius = Amaxt_Amax == 0.0f ? (float)plane_radius : (float)fmin(Amaxt_d / Amaxt_Amax, plane_radius);
    min_radius *= 1.1f;
    max_radius *= 0.9f;
    if (min_radius < max_radius) { r_min = min_radius; r_max = max_radius; }
    else { r_max = min_radius; r_min = max_radius; }
  }
  if (lane == 0) {
    j.desc[(size_t)k * 2] = r_min;
    j.desc[(size_t)k * 2 + 1] = r_max;
    j.valid[k] = (isfinite(r_min) && isfinite(r_max)) ? 1u : 0u;
  }
}

struct RsdEmitJob {
  const float4* kp;
  const float* desc_raw;
  const uint32_t* flags;
  const uint32_t* pos;
  float4* kp_out;
  float* desc_out;
  int nk;
};
__global__ void __launch_bounds__(256) rsd_emit_kernel(const RsdEmitJob* __restrict__ jobs)
{
  const RsdEmitJob& j = jobs[blockIdx.y];
  const int kp = blockIdx.x * blockDim.x + threadIdx.x;
  if (kp >= j.nk || !j.flags[kp]) return;
  const uint32_t o = j.pos[kp];
  j.desc_out[(size_t)o * 2] = j.desc_raw[(size_t)kp * 2];
  j.desc_out[(size_t)o * 2 + 1] = j.desc_raw[(size_t)kp * 2 + 1];
  j.kp_out[o] = j.kp[kp];
}

}  // namespace

void rsd_batch(Ctx& c, const std::vector<CloudView>& clouds, const std::vector<DIndex>& idx, const std::vector<const float4*>& normals,
               std::vector<DCloud>& keypoints, double radius, std::vector<DBuf<float>>& desc)
{
  const int M = (int)clouds.size();
  desc.clear();
  desc.resize(M);
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
  DBuf<uint32_t> flags(c, totalk), pos(c, totalk);
  std::vector<DBuf<float>> raw(M);
  std::vector<RsdJob> jobs(M);
  for (int m = 0; m < M; ++m) {
    raw[m].alloc(c, (size_t)nks[m] * 2);
    jobs[m] = RsdJob{idx[m].v, normals[m], keypoints[m].pts.p, nks[m], raw[m].p, flags.p + segk[m].off};
  }
  DBuf<RsdJob> dj = to_device(c, jobs);
  const float r2 = (float)(radius * radius);
  const int rv = (int)std::ceil(radius / (double)idx[0].v.leaf) + 1;
  { double b = 0; for (int m = 0; m < M; ++m) b += 32.0 * clouds[m].n + 24.0 * nks[m]; MM_BYTES(c, b); }
  MM_LAUNCH(c, rsd_kernel, dim3((mxk + RB / 32 - 1) / (RB / 32), M), RB, 0, dj.p, r2, rv, radius, 0.2);
  std::vector<int> totals;
  scan_flags_batch(c, flags.p, pos.p, segk, totals);
  std::vector<DCloud> kept(M);
  std::vector<RsdEmitJob> ej(M);
  for (int m = 0; m < M; ++m) {
    kept[m].n = totals[m];
    kept[m].pts.alloc(c, totals[m]);
    desc[m].alloc(c, (size_t)totals[m] * 2);
    ej[m] = RsdEmitJob{keypoints[m].pts.p, raw[m].p, flags.p + segk[m].off, pos.p + segk[m].off, kept[m].pts.p, desc[m].p, nks[m]};
  }
  DBuf<RsdEmitJob> dej = to_device(c, ej);
  MM_LAUNCH(c, rsd_emit_kernel, dim3((mxk + 255) / 256, M), 256, 0, dej.p);
  for (int m = 0; m < M; ++m) keypoints[m] = std::move(kept[m]);
}

}  // namespace mm3d
