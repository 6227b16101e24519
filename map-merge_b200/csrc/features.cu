// features.cu — per-map feature pipeline kernels (all maps per launch):
//   K3 radius outlier removal   <- pcl::RadiusOutlierRemoval  (map_merge_3d/src/features.cpp:31-43)
//   K4 surface normals          <- pcl::NormalEstimation      (src/features.cpp:168-179)
//   K5 SIFT3D keypoints         <- pcl::SIFTKeypoint          (src/features.cpp:45-62, 85-96)
//   K7 FPFH descriptors         <- pcl::FPFHEstimation        (src/features.cpp:99-166, src/dispatch_descriptors.h:40)
// One thread owns one query point and walks its neighbourhood in ascending
// point index, so every float accumulation happens in the canonical order.
#include <algorithm>
#include <cmath>
#include <limits>

#include "mm3d_internal.cuh"
#include "pair_features.cuh"

namespace mm3d {

namespace {

constexpr int FB = 128;  // block size of the neighbourhood kernels

static int radius_voxels(double radius, float leaf) { return (int)std::ceil(radius / (double)leaf) + 1; }

static int max_n(const std::vector<CloudView>& in)
{
  int mx = 0;
  for (const CloudView& v : in) mx = std::max(mx, v.n);
  return mx;
}

static std::vector<Seg> make_segs(const std::vector<int>& ns, int* total)
{
  std::vector<Seg> segs(ns.size());
  int off = 0;
  for (size_t m = 0; m < ns.size(); ++m) {
    segs[m].off = off;
    segs[m].n = ns[m];
    off += ns[m];
  }
  *total = off;
  return segs;
}

// ---------------------------------------------------------------- K3 outliers
struct OutlierJob {
  GridView g;
  uint32_t* flags;  // per original point
  int* counts;      // optional
};

// thread per point (a warp-per-point version with the flattened walk was measured: 1.9x slower, the visitor is too cheap)
__global__ void __launch_bounds__(FB) outlier_kernel(const OutlierJob* __restrict__ jobs, float r2, int rv, int min_nb)
{
  const OutlierJob& j = jobs[blockIdx.y];
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= j.g.n) return;
  const float4 q = j.g.pts[k];
  int cnt = 0;
  for_each_in_radius(j.g, true, q.x, q.y, q.z, r2, rv, [&](int, const float4&, float) { ++cnt; });
  const int oi = j.g.orig ? j.g.orig[k] : k;
  j.flags[oi] = cnt > min_nb ? 1u : 0u;  // "k <= min_pts_radius_" is an outlier
  if (j.counts) j.counts[oi] = cnt;
}

struct CompactJob {
  const float4* src;
  const uint32_t* flags;
  const uint32_t* pos;
  float4* dst;
  int n;
};
__global__ void __launch_bounds__(256) compact_points_kernel(const CompactJob* __restrict__ jobs)
{
  const CompactJob& j = jobs[blockIdx.y];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= j.n) return;
  if (j.flags[i]) j.dst[j.pos[i]] = j.src[i];
}

// ---------------------------------------------------------------- K4 normals
struct NormalJob {
  GridView g;
  float4* normals;  // per original point
};

__global__ void __launch_bounds__(FB) normals_kernel(const NormalJob* __restrict__ jobs, float r2, int rv)
{
  const NormalJob& j = jobs[blockIdx.y];
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = k < j.g.n;
  const float4 q = live ? j.g.pts[k] : make_float4(0.f, 0.f, 0.f, 0.f);
  // pcl::computeMeanAndCovarianceMatrix, single pass, float accumulators
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f, a5 = 0.f, a6 = 0.f, a7 = 0.f, a8 = 0.f;
  int cnt = 0;
  for_each_in_radius(j.g, live, q.x, q.y, q.z, r2, rv, [&](int, const float4& p, float) {
    a0 += p.x * p.x; a1 += p.x * p.y; a2 += p.x * p.z;
    a3 += p.y * p.y; a4 += p.y * p.z; a5 += p.z * p.z;
    a6 += p.x; a7 += p.y; a8 += p.z;
    ++cnt;
  });
  if (!live) return;
  const int oi = j.g.orig ? j.g.orig[k] : k;
  if (cnt < 3) {
    const float nanv = __int_as_float(0x7fc00000);
    j.normals[oi] = make_float4(nanv, nanv, nanv, nanv);
    return;
  }
  const float n = (float)cnt;
  a0 /= n; a1 /= n; a2 /= n; a3 /= n; a4 /= n; a5 /= n; a6 /= n; a7 /= n; a8 /= n;
  float cov[9];
  cov[0] = a0 - a6 * a6;
  cov[1] = a1 - a6 * a7;
  cov[2] = a2 - a6 * a8;
  cov[4] = a3 - a7 * a7;
  cov[5] = a4 - a7 * a8;
  cov[8] = a5 - a8 * a8;
  cov[3] = cov[1]; cov[6] = cov[2]; cov[7] = cov[5];
  float ev, vec[3];
  em::eigen33_smallest(cov, &ev, vec);
  float nx = vec[0], ny = vec[1], nz = vec[2];
  const float eig_sum = cov[0] + cov[4] + cov[8];
  const float curv = (eig_sum != 0.f) ? fabsf(ev / eig_sum) : 0.f;
  const float vx = 0.0f - q.x, vy = 0.0f - q.y, vz = 0.0f - q.z;  // viewpoint (0,0,0)
  const float cos_theta = (vx * nx + vy * ny + vz * nz);
  if (cos_theta < 0.f) { nx *= -1.f; ny *= -1.f; nz *= -1.f; }
  j.normals[oi] = make_float4(nx, ny, nz, curv);
}

// ---------------------------------------------------------------- K5 SIFT
struct SiftJob {
  GridView g;       // for the scale-space kernel g.pts holds (x, y, z, intensity)
  float* dog;       // n x 5
  uint32_t* flags;  // n x 3 (levels 1..3)
};
struct SiftScales {
  float sigma_sqr[6];
};

__device__ __forceinline__ float sift_intensity(float w)
{
  const uint32_t c = __float_as_uint(w);
  const int r = (c >> 16) & 0xff, g = (c >> 8) & 0xff, b = c & 0xff;
  return (float)(299 * r + 587 * g + 114 * b) / 1000.0f;
}

// (x, y, z, rgba) -> (x, y, z, intensity): the scale-space walk visits every point ~300 times, so the intensity
// (integer luma / 1000.0f, pcl::SIFTKeypointFieldSelector<PointXYZRGB>) is computed once per point
struct SiftPrepJob {
  const float4* src;
  float4* dst;
  int n;
};
__global__ void __launch_bounds__(256) sift_prep_kernel(const SiftPrepJob* __restrict__ jobs)
{
  const SiftPrepJob& j = jobs[blockIdx.y];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= j.n) return;
  const float4 p = j.src[i];
  j.dst[i] = make_float4(p.x, p.y, p.z, sift_intensity(p.w));
}

// computeScaleSpace, warp per point.  The Gaussian-weighted sums are accumulated in 2^-32 fixed point (int64), which makes
// them independent of the order the terms arrive in and the same bits as the CPU checker's int64 sums.  (PCL adds the
// same terms in float, nearest neighbour first; the fixed-point sum is the exactly rounded value of that.)
// A neighbour at distance d contributes to the scales with d^2 <= 9 sigma^2, i.e. to the LAST cnt of the six: the
// (neighbour, scale) pairs of a batch of 32 candidates are flattened with a warp scan so that every lane evaluates one
// exp() per round, instead of most lanes idling on the small scales.
__global__ void __launch_bounds__(FB) sift_scale_space_kernel(const SiftJob* __restrict__ jobs, SiftScales sc, float r2, int rv)
{
  __shared__ long long acc[FB / 32][12][32];  // [warp][scale: num 0..5, den 6..11][lane]
  const SiftJob& j = jobs[blockIdx.y];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int k = blockIdx.x * (FB / 32) + w;
  if (k >= j.g.n) return;  // warp-uniform
  const unsigned full = 0xffffffffu;
  const float4 q = j.g.pts[k];
#pragma unroll
  for (int s = 0; s < 12; ++s) acc[w][s][lane] = 0;
  warp_radius_unordered(j.g, q.x, q.y, q.z, r2, rv, [&](bool valid, int, const float4& p, float d2) {
    int cnt = 0;
    if (valid) {
#pragma unroll
      for (int s = 0; s < 6; ++s) cnt += (d2 <= 9 * sc.sigma_sqr[s]) ? 1 : 0;
    }
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(full, incl, o);
      if (lane >= o) incl += t;
    }
    const int total = __shfl_sync(full, incl, 31);
    const int excl = incl - cnt;
    for (int t0 = 0; t0 < total; t0 += 32) {
      const int t = t0 + lane;
      int L = 0;  // owner of pair t: smallest L with incl[L] > t
#pragma unroll
      for (int step = 16; step >= 1; step >>= 1) {
        const int v = __shfl_sync(full, incl, L + step - 1);
        if (v <= t) L += step;
      }
      L = min(L, 31);
      const float od2 = __shfl_sync(full, d2, L), oval = __shfl_sync(full, p.w, L);
      const int oex = __shfl_sync(full, excl, L), ocnt = __shfl_sync(full, cnt, L);
      if (t < total) {
        const int s = (6 - ocnt) + (t - oex);
        const float wgt = em::expf_(-0.5f * od2 / sc.sigma_sqr[s]);
        acc[w][s][lane] += em::to_fix32_pos(oval * wgt);  // intensity (sift_prep_kernel) times weight
        acc[w][6 + s][lane] += em::to_fix32_pos(wgt);
      }
    }
  });
  long long mine = 0;  // lanes 0..11 finish one sum each
  if (lane < 12)
    for (int l = 0; l < 32; ++l) mine += acc[w][lane][(l + lane) & 31];
  const long long den = __shfl_down_sync(full, mine, 6);
  float resp = 0.f;
  if (lane < 6) resp = (float)((double)mine / MM3D_FIX1_SCALE) / (float)((double)den / MM3D_FIX1_SCALE);
  const float prev = __shfl_up_sync(full, resp, 1);
  if (lane >= 1 && lane < 6) j.dog[(size_t)k * 5 + (lane - 1)] = resp - prev;
}

// findScaleSpaceExtrema, warp per point: the warp gathers the candidates of a small sphere into shared memory,
// ranks them by (d^2, index) with an all-pairs count (m is ~30-100), and reduces the DoG min / max of the 25 nearest
// with shuffles.  If the sphere holds fewer than 25 points the radius grows and the gather is repeated.
constexpr int EX_CAP = 384;  // candidates staged per warp

__global__ void __launch_bounds__(FB) sift_extrema_kernel(const SiftJob* __restrict__ jobs, float min_contrast)
{
  __shared__ int queues[FB / 32][64];
  __shared__ float cd[FB / 32][EX_CAP];
  __shared__ int ci[FB / 32][EX_CAP];
  const SiftJob& j = jobs[blockIdx.y];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = blockIdx.x * (FB / 32) + w;
  if (k >= j.g.n) return;  // warp-uniform
  const float4 q = j.g.pts[k];
  constexpr int K = 25;
  const int rv_max = max(max(j.g.div_v[0], j.g.div_v[1]), j.g.div_v[2]) + 2;
  int m = 0;
  bool overflow = false;
  for (int rv = 3;; rv = rv + max(1, rv / 3)) {
    const float rad = (float)rv * j.g.leaf;
    int fill = 0;  // warp-uniform: passing candidates arrive in dense groups
    m = warp_radius_query(j.g, q.x, q.y, q.z, rad * rad, rv + 1, queues[w], [&](int slot) {
      // called with lanes 0..g-1 active; slot order is ascending
      const int at = fill + lane;
      if (at < EX_CAP) {
        const float4 p = j.g.pts[slot];
        cd[w][at] = em::dist2_3(q.x, q.y, q.z, p.x, p.y, p.z);
        ci[w][at] = slot;
      }
      fill += 32;  // only exact for full groups; the final partial group is the last call
    });
    __syncwarp();
    if (m > EX_CAP) { overflow = true; break; }
    if (m >= K || rv > rv_max) break;
  }
  float mn[5], mx[5];
#pragma unroll
  for (int s = 0; s < 5; ++s) { mn[s] = 3.402823466e+38f; mx[s] = -3.402823466e+38f; }
  if (!overflow) {
    // rank = number of candidates that sort before mine; ranks < 25 are the 25 nearest
    for (int c = lane; c < m; c += 32) {
      const float d = cd[w][c];
      const int id = ci[w][c];
      int rank = 0;
      for (int o = 0; o < m; ++o) {
        const float od = cd[w][o];
        rank += (od < d || (od == d && ci[w][o] < id)) ? 1 : 0;
      }
      if (rank < K) {
        const float* dg = j.dog + (size_t)id * 5;
#pragma unroll
        for (int s = 0; s < 5; ++s) {
          const float v = dg[s];
          mn[s] = fminf(mn[s], v);
          mx[s] = fmaxf(mx[s], v);
        }
      }
    }
  } else if (lane == 0) {
    // rare: a neighbourhood too crowded for the staging buffer -> sequential insertion list on one lane
    float bd[K];
    int bi[K];
    int cnt = 0, rv = 3;
    for (;;) {
      cnt = 0;
      const float rad = (float)rv * j.g.leaf;
      for_each_in_radius(j.g, true, q.x, q.y, q.z, rad * rad, rv + 1, [&](int idx, const float4&, float d2) {
        if (cnt == K && !(d2 < bd[K - 1] || (d2 == bd[K - 1] && idx < bi[K - 1]))) return;
        int pos = (cnt < K) ? cnt : K - 1;
        while (pos > 0 && (d2 < bd[pos - 1] || (d2 == bd[pos - 1] && idx < bi[pos - 1]))) {
          bd[pos] = bd[pos - 1];
          bi[pos] = bi[pos - 1];
          --pos;
        }
        bd[pos] = d2;
        bi[pos] = idx;
        if (cnt < K) ++cnt;
      });
      if (cnt == K || rv > rv_max) break;
      rv *= 2;
    }
    for (int t = 0; t < cnt; ++t) {
      const float* dg = j.dog + (size_t)bi[t] * 5;
#pragma unroll
      for (int s = 0; s < 5; ++s) {
        mn[s] = fminf(mn[s], dg[s]);
        mx[s] = fmaxf(mx[s], dg[s]);
      }
    }
  }
#pragma unroll
  for (int s = 0; s < 5; ++s) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn[s] = fminf(mn[s], __shfl_xor_sync(0xffffffffu, mn[s], o));
      mx[s] = fmaxf(mx[s], __shfl_xor_sync(0xffffffffu, mx[s], o));
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int s = 1; s < 4; ++s) {
      const float val = j.dog[(size_t)k * 5 + s];
      uint32_t f = 0;
      if (fabsf(val) >= min_contrast) {
        if ((val == mn[s]) && (val < mn[s - 1]) && (val < mn[s + 1])) f = 1;
        else if ((val == mx[s]) && (val > mx[s - 1]) && (val > mx[s + 1])) f = 1;
      }
      j.flags[(size_t)k * 3 + (s - 1)] = f;
    }
  }
}

struct SiftEmitJob {
  const float4* pts;
  const uint32_t* flags;  // n x 3
  const uint32_t* pos;
  float4* dst;
  int n3;
};
__global__ void __launch_bounds__(256) sift_emit_kernel(const SiftEmitJob* __restrict__ jobs)
{
  const SiftEmitJob& j = jobs[blockIdx.y];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= j.n3) return;
  if (!j.flags[i]) return;
  const float4 p = j.pts[i / 3];
  j.dst[j.pos[i]] = make_float4(p.x, p.y, p.z, __uint_as_float(0xff000000u));  // copyPointCloud: default colour
}

// ---------------------------------------------------------------- K7 FPFH
struct FpfhJob {
  GridView g;             // surface index
  const float4* normals;  // per original surface point
  const float4* kp;       // keypoints
  int nk;
  uint32_t* need;         // per slot
  float* spfh;            // n x 33 per slot
  float* desc_raw;        // nk x 33
  uint32_t* valid3;       // nk x 3
};

// marks the surface points whose SPFH is needed (the neighbours of the keypoints): warp per keypoint, order-free walk
__global__ void __launch_bounds__(FB) fpfh_mark_kernel(const FpfhJob* __restrict__ jobs, float r2, int rv)
{
  const FpfhJob& j = jobs[blockIdx.y];
  const int t = blockIdx.x * (FB / 32) + (threadIdx.x >> 5);
  if (t >= j.nk) return;  // warp-uniform
  const float4 q = j.kp[t];
  warp_radius_unordered(j.g, q.x, q.y, q.z, r2, rv, [&](bool valid, int k, const float4&, float) {
    if (valid) j.need[k] = 1u;
  });
}

// warp per surface point: candidates that pass the radius test are queued so that the pair-feature block always
// runs with full warps; votes are shared-memory atomics on the warp's 33 counters (every vote of one histogram
// adds the same float, so a bin's value depends only on its vote count: the additions are replayed at the end).
__global__ void __launch_bounds__(FB) spfh_kernel(const FpfhJob* __restrict__ jobs, float r2, int rv, BinTable bins)
{
  __shared__ int queues[FB / 32][64];
  __shared__ unsigned short cnt[FB / 32][33][32];  // private counters per lane: no atomics, no bank conflicts
  __shared__ float thr[3][12];
  if (threadIdx.x < 36) (&thr[0][0])[threadIdx.x] = (&bins.t[0][0])[threadIdx.x];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int b = 0; b < 33; ++b) cnt[w][b][lane] = 0;
  __syncthreads();
  const FpfhJob& j = jobs[blockIdx.y];
  const int k = blockIdx.x * (FB / 32) + w;
  if (k >= j.g.n || !j.need[k]) return;  // warp-uniform
  const float4 p = j.g.pts[k];
  const float4 np = j.normals[j.g.orig ? j.g.orig[k] : k];
  const int n = warp_radius_query(j.g, p.x, p.y, p.z, r2, rv, queues[w], [&](int q) {
    if (q == k) return;
    const float4 pq = j.g.pts[q];
    const float4 nq = j.normals[j.g.orig ? j.g.orig[q] : q];
    float f1, f2, f3;
    if (!pair_features(p, np, pq, nq, &f1, &f2, &f3)) return;  // "if (!computePairFeatures (...)) continue;"
    const int h1 = lookup_bin(thr[0], 11, f1, 11.0f * 0.15915494f, 3.14159274f);
    const int h2 = lookup_bin(thr[1], 11, f2, 5.5f, 1.0f);
    const int h3 = lookup_bin(thr[2], 11, f3, 5.5f, 1.0f);
    cnt[w][h1][lane]++;
    cnt[w][11 + h2][lane]++;
    cnt[w][22 + h3][lane]++;
  });
  __syncwarp();
  const float hist_incr = 100.0f / (float)(n - 1);
  float* out = j.spfh + (size_t)k * 33;
  for (int b = lane; b < 33; b += 32) {
    unsigned int c = 0;
    for (int l = 0; l < 32; ++l) c += cnt[w][b][(l + lane) & 31];
    float h = 0.f;
    for (unsigned int t = 0; t < c; ++t) h += hist_incr;
    out[b] = h;
  }
}

// weightPointSPFHSignature, warp per keypoint: lane L owns bin L (lane 0 also bin 32), so one coalesced 132-byte row read
// per neighbour feeds all 33 float sums; the neighbours of a batch are replayed in ascending index order through shuffles.
// The three block sums are sequential double sums over (neighbour, bin) — lanes 0, 11 and 22 collect their block's eleven
// values in order with per-lane shuffle sources.
__global__ void __launch_bounds__(FB) fpfh_weight_kernel(const FpfhJob* __restrict__ jobs, float r2, int rv)
{
  __shared__ int queues[FB / 32][64];
  const FpfhJob& j = jobs[blockIdx.y];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kp = blockIdx.x * (FB / 32) + w;
  if (kp >= j.nk) return;  // warp-uniform
  const unsigned full = 0xffffffffu;
  const float4 q = j.kp[kp];
  const int f = lane < 11 ? 0 : (lane < 22 ? 1 : 2);  // feature block of bin `lane`
  const int base = f * 11;                             // collector lanes are 0, 11, 22
  float acc = 0.f, acc32 = 0.f;
  double sum = 0.0;
  const int found = warp_radius_query<true>(j.g, q.x, q.y, q.z, r2, rv, queues[w], [&](int slot) {
    float weight = 0.f;
    if (slot >= 0) {
      const float4 p = j.g.pts[slot];
      const float d2 = em::dist2_3(q.x, q.y, q.z, p.x, p.y, p.z);
      weight = d2 == 0.f ? 0.f : 1.0f / d2;  // a neighbour at distance 0 is skipped
    }
    const unsigned have = __ballot_sync(full, slot >= 0);
    const int m = __popc(have);
    for (int t = 0; t < m; ++t) {
      const int s = __shfl_sync(full, slot, t);
      const float wt = __shfl_sync(full, weight, t);
      if (wt == 0.f) continue;  // warp-uniform
      const float* h = j.spfh + (size_t)s * 33;
      const float val = h[lane] * wt;
      const float val32 = lane == 0 ? h[32] * wt : 0.f;
      acc += val;
      acc32 += val32;
      const float v32 = __shfl_sync(full, val32, 0);
#pragma unroll
      for (int i = 0; i < 11; ++i) {
        float v = __shfl_sync(full, val, (base + i) & 31);
        if (base + i == 32) v = v32;
        sum += (double)v;  // meaningful on lanes 0, 11, 22
      }
    }
  });
  // 100 / sum of the block, broadcast from its collector lane
  if (sum != 0.0) sum = 100.0 / sum;
  const float scale = (float)__shfl_sync(full, sum, base);
  const float scale2 = (float)__shfl_sync(full, sum, 22);
  const float v = acc * scale;
  const float v32 = acc32 * scale2;
  bool fin = isfinite(v);
  if (lane == 0) fin = fin && isfinite(v32);
  const bool ok = __all_sync(full, fin) && found > 0;
  j.desc_raw[(size_t)kp * 33 + lane] = v;
  if (lane == 0) j.desc_raw[(size_t)kp * 33 + 32] = v32;
  if (lane < 3) j.valid3[kp * 3 + lane] = ok ? 1u : 0u;
}

struct FpfhFlagJob {
  const uint32_t* valid3;
  uint32_t* flags;
  int nk;
};
__global__ void __launch_bounds__(256) fpfh_flag_kernel(const FpfhFlagJob* __restrict__ jobs)
{
  const FpfhFlagJob& j = jobs[blockIdx.y];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= j.nk) return;
  j.flags[i] = (j.valid3[3 * i] & j.valid3[3 * i + 1] & j.valid3[3 * i + 2]);
}

struct FpfhEmitJob {
  const float4* kp;
  const float* desc_raw;
  const uint32_t* flags;
  const uint32_t* pos;
  float4* kp_out;
  float* desc_out;
  int nk;
};
__global__ void __launch_bounds__(256) fpfh_emit_kernel(const FpfhEmitJob* __restrict__ jobs)
{
  const FpfhEmitJob& j = jobs[blockIdx.y];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= j.nk * 33) return;
  const int kp = i / 33, b = i - kp * 33;
  if (!j.flags[kp]) return;
  const uint32_t o = j.pos[kp];
  j.desc_out[(size_t)o * 33 + b] = j.desc_raw[i];
  if (b == 0) j.kp_out[o] = j.kp[kp];
}

// ---------------------------------------------------------------- K6 Harris3D
struct HarrisJob {
  GridView g;
  const float4* normals;  // per original point
  float* response;        // per slot
  uint32_t* flags;        // per original point
  const float4* cloud;    // original order
  float4* corners;
  int n_corners;
};

// responseHarris + calculateNormalCovar (SSE branch: sums divided by the count)
__global__ void __launch_bounds__(FB) harris_response_kernel(const HarrisJob* __restrict__ jobs, float r2, int rv)
{
  const HarrisJob& j = jobs[blockIdx.y];
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = k < j.g.n;
  const float4 q = live ? j.g.pts[k] : make_float4(0.f, 0.f, 0.f, 0.f);
  float c0 = 0.f, c1 = 0.f, c2 = 0.f, c5 = 0.f, c6 = 0.f, c7 = 0.f;
  unsigned count = 0;
  for_each_in_radius(j.g, live, q.x, q.y, q.z, r2, rv, [&](int s, const float4&, float) {
    const float4 nq = j.normals[j.g.orig ? j.g.orig[s] : s];
    if (!isfinite(nq.x)) return;
    c0 += nq.x * nq.x; c1 += nq.x * nq.y; c2 += nq.x * nq.z;
    c5 += nq.y * nq.y; c6 += nq.y * nq.z;
    c7 += nq.z * nq.z;
    ++count;
  });
  if (!live) return;
  if (count > 0) {
    const float cn = (float)count;
    c0 /= cn; c1 /= cn; c2 /= cn; c5 /= cn; c6 /= cn; c7 /= cn;
  } else {
    c0 = c1 = c2 = c5 = c6 = c7 = 0.f;
  }
  float resp = 0.0f;
  const float trace = c0 + c5 + c7;
  if (trace != 0.f) {
    const float det = c0 * c5 * c7 + 2.0f * c1 * c2 * c6 - c2 * c2 * c5 - c1 * c1 * c7 - c6 * c6 * c0;
    resp = 0.04f + det - 0.04f * trace * trace;
  }
  j.response[k] = resp;
}

// non-maximum suppression: keep a point unless a radius neighbour has a strictly larger response
__global__ void __launch_bounds__(FB) harris_nms_kernel(const HarrisJob* __restrict__ jobs, float r2, int rv, float threshold)
{
  const HarrisJob& j = jobs[blockIdx.y];
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = k < j.g.n;
  const float resp = live ? j.response[k] : 0.f;
  const bool cand = live && isfinite(resp) && !(resp < threshold);
  const float4 q = cand ? j.g.pts[k] : make_float4(0.f, 0.f, 0.f, 0.f);
  bool is_max = true;
  for_each_in_radius(j.g, cand, q.x, q.y, q.z, r2, rv, [&](int s, const float4&, float) {
    if (resp < j.response[s]) is_max = false;
  });
  if (!live) return;
  j.flags[j.g.orig ? j.g.orig[k] : k] = (cand && is_max) ? 1u : 0u;
}

struct HarrisEmitJob {
  const float4* cloud;
  const uint32_t* flags;
  const uint32_t* pos;
  float4* dst;
  int n;
};
__global__ void __launch_bounds__(256) harris_emit_kernel(const HarrisEmitJob* __restrict__ jobs)
{
  const HarrisEmitJob& j = jobs[blockIdx.y];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= j.n) return;
  if (!j.flags[i]) return;
  const float4 p = j.cloud[i];
  j.dst[j.pos[i]] = make_float4(p.x, p.y, p.z, __uint_as_float(0xff000000u));
}

// refineCorners: c <- (sum n n^T)^-1 sum n n^T p over the radius neighbours of c, at most 10 times
__global__ void __launch_bounds__(FB) harris_refine_kernel(const HarrisJob* __restrict__ jobs, float r2, int rv)
{
  const HarrisJob& j = jobs[blockIdx.y];
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = t < j.n_corners;
  float4 corner = live ? j.corners[t] : make_float4(0.f, 0.f, 0.f, 0.f);
  unsigned iterations = 0;
  float diff;
  bool working = live;
  while (__any_sync(0xffffffffu, working)) {
    float N[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, Np[3] = {0.f, 0.f, 0.f};
    const float cx = corner.x, cy = corner.y, cz = corner.z;
    for_each_in_radius(j.g, working, cx, cy, cz, r2, rv, [&](int s, const float4& p, float) {
      const float4 nq = j.normals[j.g.orig ? j.g.orig[s] : s];
      if (!isfinite(nq.x)) return;
      const float nv[3] = {nq.x, nq.y, nq.z};
      float nnT[9];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) nnT[r * 3 + c] = nv[r] * nv[c];
#pragma unroll
      for (int e = 0; e < 9; ++e) N[e] += nnT[e];
#pragma unroll
      for (int r = 0; r < 3; ++r) Np[r] += (nnT[r * 3 + 0] * p.x + nnT[r * 3 + 1] * p.y) + nnT[r * 3 + 2] * p.z;
    });
    if (!working) continue;
    // pcl::invert3x3SymMatrix
    const float fd_ee = N[4] * N[8] - N[7] * N[5];
    const float ce_bf = N[2] * N[5] - N[1] * N[8];
    const float be_cd = N[1] * N[5] - N[2] * N[4];
    const float det = N[0] * fd_ee + N[1] * ce_bf + N[2] * be_cd;
    if (det != 0.f) {
      float inv[9];
      inv[0] = fd_ee;
      inv[1] = inv[3] = ce_bf;
      inv[2] = inv[6] = be_cd;
      inv[4] = (N[0] * N[8] - N[2] * N[2]);
      inv[5] = inv[7] = (N[1] * N[2] - N[0] * N[5]);
      inv[8] = (N[0] * N[4] - N[1] * N[1]);
#pragma unroll
      for (int e = 0; e < 9; ++e) inv[e] /= det;
      corner.x = (inv[0] * Np[0] + inv[1] * Np[1]) + inv[2] * Np[2];
      corner.y = (inv[3] * Np[0] + inv[4] * Np[1]) + inv[5] * Np[2];
      corner.z = (inv[6] * Np[0] + inv[7] * Np[1]) + inv[8] * Np[2];
    }
    const float dx = corner.x - cx, dy = corner.y - cy, dz = corner.z - cz;
    diff = (dx * dx + dy * dy) + dz * dz;
    if (!((double)diff > 1e-6 && ++iterations < 10)) working = false;
  }
  if (live) j.corners[t] = corner;
}

}  // namespace

// ===========================================================================
void harris_batch(Ctx& c, const std::vector<CloudView>& clouds, const std::vector<DIndex>& idx, const std::vector<const float4*>& normals,
                  float threshold, float radius_f, std::vector<DCloud>& keypoints, std::vector<DBuf<float>>* response_dbg,
                  std::vector<DCloud>* unrefined_dbg)
{
  const int M = (int)clouds.size();
  keypoints.clear();
  keypoints.resize(M);
  if (response_dbg) { response_dbg->clear(); response_dbg->resize(M); }
  if (unrefined_dbg) { unrefined_dbg->clear(); unrefined_dbg->resize(M); }
  if (M == 0) return;
  std::vector<int> ns(M);
  for (int m = 0; m < M; ++m) ns[m] = clouds[m].n;
  int total = 0;
  std::vector<Seg> segs = make_segs(ns, &total);
  const int mx = max_n(clouds);
  if (total == 0) return;
  // search_radius_ is the float radius widened to double (HarrisKeypoint3D::setRadius(float))
  const double rd = (double)radius_f;
  const float r2 = (float)(rd * rd);
  const int rv = radius_voxels(rd, idx[0].v.leaf);
  DBuf<uint32_t> flags(c, total), pos(c, total);
  std::vector<DBuf<float>> resp(M);
  std::vector<HarrisJob> jobs(M);
  for (int m = 0; m < M; ++m) {
    resp[m].alloc(c, ns[m]);
    jobs[m].g = idx[m].v;
    jobs[m].normals = normals[m];
    jobs[m].response = resp[m].p;
    jobs[m].flags = flags.p + segs[m].off;
    jobs[m].cloud = clouds[m].pts;
    jobs[m].corners = nullptr;
    jobs[m].n_corners = 0;
  }
  DBuf<HarrisJob> dj = to_device(c, jobs);
  const dim3 grid((mx + FB - 1) / FB, M);
  { double b = 0; for (int m = 0; m < M; ++m) b += 36.0 * ns[m]; MM_BYTES(c, b); }
  MM_LAUNCH(c, harris_response_kernel, grid, FB, 0, dj.p, r2, rv);
  { double b = 0; for (int m = 0; m < M; ++m) b += 24.0 * ns[m]; MM_BYTES(c, b); }
  MM_LAUNCH(c, harris_nms_kernel, grid, FB, 0, dj.p, r2, rv, threshold);
  std::vector<int> totals;
  scan_flags_batch(c, flags.p, pos.p, segs, totals);
  std::vector<HarrisEmitJob> ej(M);
  int mxc = 0;
  for (int m = 0; m < M; ++m) {
    keypoints[m].n = totals[m];
    keypoints[m].pts.alloc(c, totals[m]);
    ej[m] = HarrisEmitJob{clouds[m].pts, flags.p + segs[m].off, pos.p + segs[m].off, keypoints[m].pts.p, ns[m]};
    jobs[m].corners = keypoints[m].pts.p;
    jobs[m].n_corners = totals[m];
    mxc = std::max(mxc, totals[m]);
  }
  DBuf<HarrisEmitJob> dej = to_device(c, ej);
  MM_LAUNCH(c, harris_emit_kernel, dim3((mx + 255) / 256, M), 256, 0, dej.p);
  if (unrefined_dbg)
    for (int m = 0; m < M; ++m) {
      (*unrefined_dbg)[m].n = totals[m];
      (*unrefined_dbg)[m].pts.alloc(c, totals[m]);
      if (totals[m]) MM_CUDA(cudaMemcpyAsync((*unrefined_dbg)[m].pts.p, keypoints[m].pts.p, (size_t)totals[m] * sizeof(float4), cudaMemcpyDeviceToDevice, c.stream));
    }
  if (mxc > 0) {
    DBuf<HarrisJob> dj2 = to_device(c, jobs);
    MM_LAUNCH(c, harris_refine_kernel, dim3((mxc + FB - 1) / FB, M), FB, 0, dj2.p, r2, rv);
  }
  if (response_dbg)
    for (int m = 0; m < M; ++m) (*response_dbg)[m] = std::move(resp[m]);
}

void remove_outliers_batch(Ctx& c, const std::vector<CloudView>& clouds, const std::vector<DIndex>& idx, double radius, int min_nb,
                           std::vector<DCloud>& out, std::vector<DBuf<int>>* counts)
{
  const int M = (int)clouds.size();
  out.clear();
  out.resize(M);
  if (counts) { counts->clear(); counts->resize(M); }
  if (M == 0) return;
  std::vector<int> ns(M);
  for (int m = 0; m < M; ++m) ns[m] = clouds[m].n;
  int total = 0;
  std::vector<Seg> segs = make_segs(ns, &total);
  if (total == 0) return;
  DBuf<uint32_t> flags(c, total), pos(c, total);
  std::vector<OutlierJob> jobs(M);
  for (int m = 0; m < M; ++m) {
    jobs[m].g = idx[m].v;
    jobs[m].flags = flags.p + segs[m].off;
    jobs[m].counts = nullptr;
    if (counts) {
      (*counts)[m].alloc(c, ns[m]);
      jobs[m].counts = (*counts)[m].p;
    }
  }
  DBuf<OutlierJob> dj = to_device(c, jobs);
  const float r2 = (float)(radius * radius);
  const int rv = radius_voxels(radius, idx[0].v.leaf);
  const int mx = max_n(clouds);
  { double b = 0; for (int m = 0; m < M; ++m) b += 20.0 * ns[m]; MM_BYTES(c, b); }
  MM_LAUNCH(c, outlier_kernel, dim3((mx + FB - 1) / FB, M), FB, 0, dj.p, r2, rv, min_nb);
  std::vector<int> totals;
  scan_flags_batch(c, flags.p, pos.p, segs, totals);
  std::vector<CompactJob> cj(M);
  for (int m = 0; m < M; ++m) {
    out[m].n = totals[m];
    out[m].pts.alloc(c, totals[m]);
    cj[m] = CompactJob{clouds[m].pts, flags.p + segs[m].off, pos.p + segs[m].off, out[m].pts.p, ns[m]};
  }
  DBuf<CompactJob> dcj = to_device(c, cj);
  MM_LAUNCH(c, compact_points_kernel, dim3((mx + 255) / 256, M), 256, 0, dcj.p);
}

void normals_batch(Ctx& c, const std::vector<CloudView>& clouds, const std::vector<DIndex>& idx, double radius,
                   std::vector<DBuf<float4>>& normals)
{
  const int M = (int)clouds.size();
  normals.clear();
  normals.resize(M);
  if (M == 0) return;
  std::vector<NormalJob> jobs(M);
  for (int m = 0; m < M; ++m) {
    normals[m].alloc(c, clouds[m].n);
    jobs[m].g = idx[m].v;
    jobs[m].normals = normals[m].p;
  }
  const int mx = max_n(clouds);
  if (mx == 0) return;
  DBuf<NormalJob> dj = to_device(c, jobs);
  const float r2 = (float)(radius * radius);
  const int rv = radius_voxels(radius, idx[0].v.leaf);
  { double b = 0; for (int m = 0; m < M; ++m) b += 32.0 * clouds[m].n; MM_BYTES(c, b); }
  MM_LAUNCH(c, normals_kernel, dim3((mx + FB - 1) / FB, M), FB, 0, dj.p, r2, rv);
}

void sift_batch(Ctx& c, const std::vector<CloudView>& clouds, float min_scale, int n_octaves, int n_scales, float min_contrast,
                std::vector<DCloud>& keypoints, std::vector<DBuf<float>>* dog0)
{
  const int M = (int)clouds.size();
  keypoints.clear();
  keypoints.resize(M);
  if (dog0) { dog0->clear(); dog0->resize(M); }
  if (M == 0) return;
  if (n_scales != 3) throw std::runtime_error("sift_batch: the reference path uses 3 scales per octave");
  std::vector<char> active(M, 1);
  std::vector<CloudView> cur = clouds;
  std::vector<std::vector<DCloud>> octave_kp(M);
  std::vector<std::vector<DCloud>> keep_alive;  // octave clouds stay alive until the stream drains
  float scale = min_scale;
  for (int oct = 0; oct < n_octaves; ++oct) {
    const float s = 1.0f * scale;
    // maps that dropped out keep an empty view
    std::vector<CloudView> in(M);
    for (int m = 0; m < M; ++m) in[m] = active[m] ? cur[m] : CloudView{nullptr, 0};
    keep_alive.emplace_back();
    std::vector<DCloud>& oc = keep_alive.back();
    std::vector<VoxGeom> ogeom;
    voxel_downsample_batch(c, in, s, oc, &ogeom);
    bool any = false;
    std::vector<CloudView> ov(M);
    for (int m = 0; m < M; ++m) {
      if (active[m] && oc[m].n < 25) active[m] = 0;  // "break": no further octaves for this cloud
      ov[m] = active[m] ? oc[m].view() : CloudView{nullptr, 0};
      cur[m] = oc[m].view();
      any = any || active[m];
    }
    if (!any) break;
    std::vector<DIndex> idx;
    build_index_batch(c, ov, s, 2, 0, 0, idx, nullptr, &ogeom);
    // scales[i] = base * 2^((i-1)/3), i = 0..5 ; sigma^2 = powf(scale, 2)
    float scales[6];
    SiftScales sc;
    for (int i = 0; i < 6; ++i) {
      scales[i] = scale * powf(2.0f, (1.0f * (float)i - 1.0f) / (float)n_scales);
      sc.sigma_sqr[i] = powf(scales[i], 2.0f);
    }
    const float max_radius = 3.0f * scales[5];
    const float r2 = (float)((double)max_radius * (double)max_radius);
    const int rv = radius_voxels((double)max_radius, s);
    std::vector<int> ns(M), n3(M);
    for (int m = 0; m < M; ++m) { ns[m] = ov[m].n; n3[m] = ov[m].n * 3; }
    int total3 = 0;
    std::vector<Seg> segs3 = make_segs(n3, &total3);
    DBuf<uint32_t> flags(c, total3), pos(c, total3);
    std::vector<DBuf<float>> dog(M);
    std::vector<SiftJob> jobs(M);
    for (int m = 0; m < M; ++m) {
      dog[m].alloc(c, (size_t)ns[m] * 5);
      jobs[m].g = idx[m].v;
      jobs[m].dog = dog[m].p;
      jobs[m].flags = flags.p + segs3[m].off;
    }
    DBuf<SiftJob> dj = to_device(c, jobs);
    const int mx = max_n(ov);
    const dim3 grid((mx + FB - 1) / FB, M);
    // intensity-carrying copy of the octave clouds for the scale-space walk (the index must be in identity order)
    std::vector<DBuf<float4>> pi(M);
    std::vector<SiftPrepJob> pj(M);
    std::vector<SiftJob> jobs_ss = jobs;
    for (int m = 0; m < M; ++m) {
      if (idx[m].v.orig) throw std::runtime_error("sift_batch: octave cloud is not in voxel order");
      pi[m].alloc(c, ns[m]);
      pj[m] = SiftPrepJob{ov[m].pts, pi[m].p, ns[m]};
      jobs_ss[m].g.pts = pi[m].p;
    }
    DBuf<SiftPrepJob> dpj = to_device(c, pj);
    MM_LAUNCH(c, sift_prep_kernel, dim3((mx + 255) / 256, M), 256, 0, dpj.p);
    DBuf<SiftJob> djs = to_device(c, jobs_ss);
    { double b = 0; for (int m = 0; m < M; ++m) b += 36.0 * ns[m]; MM_BYTES(c, b); }
    MM_LAUNCH(c, sift_scale_space_kernel, dim3((mx + FB / 32 - 1) / (FB / 32), M), FB, 0, djs.p, sc, r2, rv);
    { double b = 0; for (int m = 0; m < M; ++m) b += 48.0 * ns[m]; MM_BYTES(c, b); }
    MM_LAUNCH(c, sift_extrema_kernel, dim3((mx + FB / 32 - 1) / (FB / 32), M), FB, 0, dj.p, min_contrast);
    std::vector<int> totals;
    scan_flags_batch(c, flags.p, pos.p, segs3, totals);
    std::vector<SiftEmitJob> ej(M);
    for (int m = 0; m < M; ++m) {
      octave_kp[m].emplace_back();
      DCloud& k = octave_kp[m].back();
      k.n = totals[m];
      k.pts.alloc(c, totals[m]);
      ej[m] = SiftEmitJob{ov[m].pts, flags.p + segs3[m].off, pos.p + segs3[m].off, k.pts.p, n3[m]};
    }
    DBuf<SiftEmitJob> dej = to_device(c, ej);
    MM_LAUNCH(c, sift_emit_kernel, dim3((mx * 3 + 255) / 256, M), 256, 0, dej.p);
    if (dog0 && oct == 0)
      for (int m = 0; m < M; ++m) (*dog0)[m] = std::move(dog[m]);
    scale *= 2;
  }
  for (int m = 0; m < M; ++m) {
    int tot = 0;
    for (const DCloud& k : octave_kp[m]) tot += k.n;
    keypoints[m].n = tot;
    keypoints[m].pts.alloc(c, tot);
    int off = 0;
    for (const DCloud& k : octave_kp[m]) {
      if (k.n) MM_CUDA(cudaMemcpyAsync(keypoints[m].pts.p + off, k.pts.p, (size_t)k.n * sizeof(float4), cudaMemcpyDeviceToDevice, c.stream));
      off += k.n;
    }
  }
}

void fpfh_batch(Ctx& c, const std::vector<CloudView>& clouds, const std::vector<DIndex>& idx, const std::vector<const float4*>& normals,
                std::vector<DCloud>& keypoints, double radius, std::vector<DBuf<float>>& desc, std::vector<DBuf<float>>* spfh_dbg)
{
  const int M = (int)clouds.size();
  desc.clear();
  desc.resize(M);
  if (spfh_dbg) { spfh_dbg->clear(); spfh_dbg->resize(M); }
  if (M == 0) return;
  std::vector<int> nks(M);
  int mxk = 0;
  for (int m = 0; m < M; ++m) { nks[m] = keypoints[m].n; mxk = std::max(mxk, nks[m]); }
  int totalk = 0;
  std::vector<Seg> segk = make_segs(nks, &totalk);
  const int mx = max_n(clouds);
  if (totalk == 0 || mx == 0) {
    for (int m = 0; m < M; ++m) { keypoints[m].n = 0; keypoints[m].pts.release(); }
    return;
  }
  std::vector<DBuf<uint32_t>> need(M), valid3(M);
  std::vector<DBuf<float>> spfh(M), raw(M);
  std::vector<FpfhJob> jobs(M);
  for (int m = 0; m < M; ++m) {
    need[m].alloc(c, clouds[m].n);
    need[m].zero(c);
    spfh[m].alloc(c, (size_t)clouds[m].n * 33);
    if (spfh_dbg) spfh[m].zero(c);
    raw[m].alloc(c, (size_t)nks[m] * 33);
    valid3[m].alloc(c, (size_t)nks[m] * 3);
    jobs[m].g = idx[m].v;
    jobs[m].normals = normals[m];
    jobs[m].kp = keypoints[m].pts.p;
    jobs[m].nk = nks[m];
    jobs[m].need = need[m].p;
    jobs[m].spfh = spfh[m].p;
    jobs[m].desc_raw = raw[m].p;
    jobs[m].valid3 = valid3[m].p;
  }
  DBuf<FpfhJob> dj = to_device(c, jobs);
  const float r2 = (float)(radius * radius);
  const int rv = radius_voxels(radius, idx[0].v.leaf);
  MM_LAUNCH(c, fpfh_mark_kernel, dim3((mxk + FB / 32 - 1) / (FB / 32), M), FB, 0, dj.p, r2, rv);
  { double b = 0; for (int m = 0; m < M; ++m) b += (32.0 + 132.0 + 4.0) * clouds[m].n; MM_BYTES(c, b); }
  static const BinTable bins = make_bin_table(11);
  MM_LAUNCH(c, spfh_kernel, dim3((mx + FB / 32 - 1) / (FB / 32), M), FB, 0, dj.p, r2, rv, bins);
  { double b = 0; for (int m = 0; m < M; ++m) b += (16.0 + 132.0) * clouds[m].n + (16.0 + 132.0) * nks[m]; MM_BYTES(c, b); }
  MM_LAUNCH(c, fpfh_weight_kernel, dim3((mxk + FB / 32 - 1) / (FB / 32), M), FB, 0, dj.p, r2, rv);
  DBuf<uint32_t> flags(c, totalk), pos(c, totalk);
  std::vector<FpfhFlagJob> fj(M);
  for (int m = 0; m < M; ++m) fj[m] = FpfhFlagJob{valid3[m].p, flags.p + segk[m].off, nks[m]};
  DBuf<FpfhFlagJob> dfj = to_device(c, fj);
  MM_LAUNCH(c, fpfh_flag_kernel, dim3((mxk + 255) / 256, M), 256, 0, dfj.p);
  std::vector<int> totals;
  scan_flags_batch(c, flags.p, pos.p, segk, totals);
  std::vector<DCloud> kept(M);
  std::vector<FpfhEmitJob> ej(M);
  for (int m = 0; m < M; ++m) {
    kept[m].n = totals[m];
    kept[m].pts.alloc(c, totals[m]);
    desc[m].alloc(c, (size_t)totals[m] * 33);
    ej[m] = FpfhEmitJob{keypoints[m].pts.p, raw[m].p, flags.p + segk[m].off, pos.p + segk[m].off, kept[m].pts.p, desc[m].p, nks[m]};
  }
  DBuf<FpfhEmitJob> dej = to_device(c, ej);
  MM_LAUNCH(c, fpfh_emit_kernel, dim3((mxk * 33 + 255) / 256, M), 256, 0, dej.p);
  for (int m = 0; m < M; ++m) keypoints[m] = std::move(kept[m]);
  if (spfh_dbg)
    for (int m = 0; m < M; ++m) (*spfh_dbg)[m] = std::move(spfh[m]);
}

}  // namespace mm3d
