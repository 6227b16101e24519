// mm3d_oracle.cpp — CPU restatement of map_merge_3d's registration hot path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing in the product library links, includes or
// calls this file; only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs load liboracle.so, and there only as the
// checker or as the timed CPU baseline.
//
// PARITY STATUS: "parity unpinned" for every PCL-backed stage.  The reference
// (/root/reference/map_merge_3d) delegates all arithmetic to PCL >= 1.8, which
// is neither vendored in the reference tree nor installable here, and the
// reference's own tests (test/test_map_merging.cpp) hold no numeric vectors.
// Each stage below is therefore a restatement of upstream PCL 1.8.1 from its
// published algorithm ("[PCL-recall <upstream file>]"), anchored on the
// reference's call sites ("[REF file:line]").  What IS pinned by reference code
// compiled here (oracle/_ref, `make ref`; outputs committed under tests/golden/):
//   * the pose-graph step: the reference's graph.cpp unmodified (libgraph_ref.so);
//   * ALL of the reference's own sources: features.cpp, map_merging.cpp, matching.cpp
//     and graph.cpp unmodified, with stand-ins for the PCL classes they drive that
//     hand the work to this file's stage functions
//     (libmapmerging_ref.so, mapmerging_ref_shim.cpp) — estimateMapsTransforms /
//     computeGlobalTransforms / composeMaps control flow, how features.cpp configures
//     each PCL object and filters invalid descriptors, findFeatureCorrespondences
//     (the reciprocal k-NN cross-match), how matching.cpp configures RANSAC / SAC-IA /
//     ICP / validation, the command-line flag table and the params printout;
//   * MapMergingParams defaults and the enum layer: the reference's public headers
//     unmodified (params_ref);
//   * the five degenerate gtest cases.
//
// Canonical choices where PCL is implementation-defined:
//   * voxel centroid summation order = ascending original point index
//   * radius-search result order (FLANN traversal order) = ascending index
//   * kNN ties = lower index first; 1-NN ties = lower index
//   * outlier boundary = strict d^2 < r^2 (PCL 1.8.1 radiusSearch)
//   * ICP / score reductions (Eigen's vectorised redux order is unknowable):
//     order-independent fixed-point sums (exact_math.h to_fix)
//   * k clamped to the descriptor-set size in findFeatureCorrespondences
//   * libm transcendentals replaced by exact_math.h (IEEE-only, <=1 ulp from
//     libm); build with -DORACLE_LIBM to get the libm variant for comparison.
//
// Build: see oracle/Makefile (g++ -O2 -ffp-contract=off, no -march, no deps).
#include <algorithm>
#ifdef _OPENMP
#include <omp.h>
#endif
#include <cfloat>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <queue>
#include <random>
#include <set>
#include <vector>

#include "../map-merge_b200/csrc/exact_math.h"

namespace orc {

namespace em = mm3d::em;

#ifdef ORACLE_LIBM
static inline float m_expf(float x) { return ::expf(x); }
static inline float m_atan2f(float y, float x) { return ::atan2f(y, x); }
static inline float m_cosf(float x) { return ::cosf(x); }
static inline float m_sinf(float x) { return ::sinf(x); }
static inline double m_acos(double x) { return ::acos(x); }
static inline float m_acosf(float x) { return ::acosf(x); }
#else
static inline double m_acos(double x) { return em::acos_d_(x); }
static inline float m_acosf(float x) { return (float)em::acos_d_((double)x); }
static inline float m_expf(float x) { return em::expf_(x); }
static inline float m_atan2f(float y, float x) { return em::atan2f_(y, x); }
static inline float m_cosf(float x) { return em::cosf_small_(x); }
static inline float m_sinf(float x) { return em::sinf_small_(x); }
#endif

struct P4 {
  float x, y, z;
  uint32_t rgba;  // a<<24 | r<<16 | g<<8 | b  (pcl::PointXYZRGB::rgba)
};
struct N4 {
  float nx, ny, nz, curv;
};
typedef std::vector<P4> Cloud;
typedef std::vector<N4> Normals;

struct Mat4 {
  float m[16];  // row-major
  static Mat4 identity()
  {
    Mat4 r;
    for (int i = 0; i < 16; ++i) r.m[i] = (i % 5 == 0) ? 1.0f : 0.0f;
    return r;
  }
  static Mat4 zero()
  {
    Mat4 r;
    for (int i = 0; i < 16; ++i) r.m[i] = 0.0f;
    return r;
  }
};

// Eigen 4x4 float product, coefficient-wise ((a0b0 + a1b1) + a2b2) + a3b3
static Mat4 mul(const Mat4& a, const Mat4& b)
{
  Mat4 r;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      float acc = a.m[i * 4 + 0] * b.m[0 * 4 + j];
      acc += a.m[i * 4 + 1] * b.m[1 * 4 + j];
      acc += a.m[i * 4 + 2] * b.m[2 * 4 + j];
      acc += a.m[i * 4 + 3] * b.m[3 * 4 + j];
      r.m[i * 4 + j] = acc;
    }
  return r;
}

// general 4x4 float inverse by cofactors (Eigen::Matrix4f::inverse(), tolerance-level)
static Mat4 inverse(const Mat4& a)
{
  const float* m = a.m;
  float inv[16];
  inv[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
  inv[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
  inv[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
  inv[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
  inv[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
  inv[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
  inv[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
  inv[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
  inv[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
  inv[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
  inv[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
  inv[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
  inv[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
  inv[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
  inv[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
  inv[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
  const float det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
  Mat4 r;
  for (int i = 0; i < 16; ++i) r.m[i] = inv[i] / det;
  return r;
}

static inline void xform(const Mat4& t, float x, float y, float z, float& ox, float& oy, float& oz)
{
  ox = ((t.m[0] * x + t.m[1] * y) + t.m[2] * z) + t.m[3];
  oy = ((t.m[4] * x + t.m[5] * y) + t.m[6] * z) + t.m[7];
  oz = ((t.m[8] * x + t.m[9] * y) + t.m[10] * z) + t.m[11];
}

static inline float d2(const P4& a, float bx, float by, float bz)
{
  // flann::L2_Simple: sequential, query - data has the same square either way
  const float dx = bx - a.x, dy = by - a.y, dz = bz - a.z;
  float r = dx * dx;
  r += dy * dy;
  r += dz * dz;
  return r;
}

// ===========================================================================
// Uniform-grid neighbour search (stands in for pcl::search::KdTree / FLANN;
// results are exact, so only the enumeration order is a choice).
// ===========================================================================
struct Grid {
  const Cloud* c = nullptr;
  float cell = 1.0f, inv = 1.0f;
  int mn[3] = {0, 0, 0}, dim[3] = {0, 0, 0};
  std::vector<int> start, order;

  int coord(float v, int a) const { return (int)std::floor(v * inv) - mn[a]; }

  void build(const Cloud& cl, float cellsize)
  {
    c = &cl;
    cell = cellsize;
    inv = 1.0f / cellsize;
    start.clear();
    order.clear();
    if (cl.empty()) {
      dim[0] = dim[1] = dim[2] = 0;
      return;
    }
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (const P4& p : cl) {
      lo[0] = std::min(lo[0], p.x); hi[0] = std::max(hi[0], p.x);
      lo[1] = std::min(lo[1], p.y); hi[1] = std::max(hi[1], p.y);
      lo[2] = std::min(lo[2], p.z); hi[2] = std::max(hi[2], p.z);
    }
    for (int a = 0; a < 3; ++a) {
      mn[a] = (int)std::floor(lo[a] * inv);
      dim[a] = (int)std::floor(hi[a] * inv) - mn[a] + 1;
    }
    // keep the dense table bounded
    while ((int64_t)dim[0] * dim[1] * dim[2] > (int64_t)64 * 1000 * 1000) {
      cell *= 2.0f;
      inv = 1.0f / cell;
      for (int a = 0; a < 3; ++a) {
        mn[a] = (int)std::floor(lo[a] * inv);
        dim[a] = (int)std::floor(hi[a] * inv) - mn[a] + 1;
      }
    }
    const size_t ncell = (size_t)dim[0] * dim[1] * dim[2];
    start.assign(ncell + 1, 0);
    std::vector<int> cid(cl.size());
    for (size_t i = 0; i < cl.size(); ++i) {
      const int cx = coord(cl[i].x, 0), cy = coord(cl[i].y, 1), cz = coord(cl[i].z, 2);
      cid[i] = (cz * dim[1] + cy) * dim[0] + cx;
      ++start[cid[i] + 1];
    }
    for (size_t k = 0; k < ncell; ++k) start[k + 1] += start[k];
    order.resize(cl.size());
    std::vector<int> fill(start.begin(), start.end() - 1);
    for (size_t i = 0; i < cl.size(); ++i) order[fill[cid[i]]++] = (int)i;
  }

  // visit every point with d^2 < r2 (strict, FLANN RadiusResultSet), any order
  template <typename F>
  void for_radius(float qx, float qy, float qz, float r, float r2, F f) const
  {
    if (!c || c->empty()) return;
    int lo[3], hi[3];
    const float q[3] = {qx, qy, qz};
    for (int a = 0; a < 3; ++a) {
      lo[a] = std::max(coord(q[a] - r, a) - 1, 0);
      hi[a] = std::min(coord(q[a] + r, a) + 1, dim[a] - 1);
      if (lo[a] > hi[a]) return;
    }
    for (int cz = lo[2]; cz <= hi[2]; ++cz)
      for (int cy = lo[1]; cy <= hi[1]; ++cy) {
        const int base = (cz * dim[1] + cy) * dim[0];
        const int s = start[base + lo[0]], e = start[base + hi[0] + 1];
        for (int k = s; k < e; ++k) {
          const int i = order[k];
          const float dd = d2((*c)[i], qx, qy, qz);
          if (dd < r2) f(i, dd);
        }
      }
  }

  // canonical radius search: ascending index
  void radius_sorted(float qx, float qy, float qz, double radius, std::vector<int>& idx, std::vector<float>& dist) const
  {
    const float r2 = (float)(radius * radius);  // pcl::KdTreeFLANN::radiusSearch
    std::vector<std::pair<int, float>> tmp;
    for_radius(qx, qy, qz, (float)radius, r2, [&](int i, float dd) { tmp.emplace_back(i, dd); });
    std::sort(tmp.begin(), tmp.end());
    idx.resize(tmp.size());
    dist.resize(tmp.size());
    for (size_t k = 0; k < tmp.size(); ++k) {
      idx[k] = tmp[k].first;
      dist[k] = tmp[k].second;
    }
  }

  // shell-expanding k-NN, result sorted by (d2, index)
  void knn(float qx, float qy, float qz, int k, std::vector<std::pair<float, int>>& best) const
  {
    best.clear();
    if (!c || c->empty()) return;
    const int q[3] = {coord(qx, 0), coord(qy, 1), coord(qz, 2)};
    int maxshell = 0;
    for (int a = 0; a < 3; ++a) maxshell = std::max(maxshell, std::max(std::abs(q[a]), std::abs(q[a] - (dim[a] - 1))) + 1);
    for (int s = 0; s <= maxshell; ++s) {
      // after finishing shell s-1 every unvisited point is >= (s-1)*cell away
      if ((int)best.size() >= k && s >= 1) {
        const float guard = (float)(s - 1) * cell;
        if (best.back().first <= guard * guard) break;
      }
      for (int cz = q[2] - s; cz <= q[2] + s; ++cz) {
        if (cz < 0 || cz >= dim[2]) continue;
        for (int cy = q[1] - s; cy <= q[1] + s; ++cy) {
          if (cy < 0 || cy >= dim[1]) continue;
          const bool face = (std::abs(cz - q[2]) == s) || (std::abs(cy - q[1]) == s);
          for (int cx = q[0] - s; cx <= q[0] + s; ++cx) {
            if (cx < 0 || cx >= dim[0]) continue;
            if (!face && std::abs(cx - q[0]) != s) continue;
            const int cidx = (cz * dim[1] + cy) * dim[0] + cx;
            for (int t = start[cidx]; t < start[cidx + 1]; ++t) {
              const int i = order[t];
              const std::pair<float, int> cand(d2((*c)[i], qx, qy, qz), i);
              if ((int)best.size() < k) {
                best.insert(std::upper_bound(best.begin(), best.end(), cand), cand);
              } else if (cand < best.back()) {
                best.pop_back();
                best.insert(std::upper_bound(best.begin(), best.end(), cand), cand);
              }
            }
          }
        }
      }
    }
  }

  // nearest neighbour among points with (double)d2 <= bound; ties -> lower index
  bool nn1(float qx, float qy, float qz, double bound, int& idx, float& dist) const
  {
    if (!c || c->empty()) return false;
    const float r = (float)std::sqrt(bound);
    bool found = false;
    int bi = -1;
    float bd = 0.f;
    int lo[3], hi[3];
    const float q[3] = {qx, qy, qz};
    for (int a = 0; a < 3; ++a) {
      lo[a] = std::max(coord(q[a] - r, a) - 1, 0);
      hi[a] = std::min(coord(q[a] + r, a) + 1, dim[a] - 1);
      if (lo[a] > hi[a]) return false;
    }
    const int qc[3] = {coord(qx, 0), coord(qy, 1), coord(qz, 2)};
    int maxshell = 0;
    for (int a = 0; a < 3; ++a) maxshell = std::max(maxshell, std::max(qc[a] - lo[a], hi[a] - qc[a]));
    maxshell = std::max(maxshell, 0);
    for (int s = 0; s <= maxshell; ++s) {
      if (found && s >= 1) {
        const float guard = (float)(s - 1) * cell;
        if (bd <= guard * guard) break;
      }
      for (int cz = std::max(qc[2] - s, lo[2]); cz <= std::min(qc[2] + s, hi[2]); ++cz)
        for (int cy = std::max(qc[1] - s, lo[1]); cy <= std::min(qc[1] + s, hi[1]); ++cy) {
          const bool face = (std::abs(cz - qc[2]) == s) || (std::abs(cy - qc[1]) == s);
          for (int cx = std::max(qc[0] - s, lo[0]); cx <= std::min(qc[0] + s, hi[0]); ++cx) {
            if (!face && std::abs(cx - qc[0]) != s) continue;
            const int cidx = (cz * dim[1] + cy) * dim[0] + cx;
            for (int t = start[cidx]; t < start[cidx + 1]; ++t) {
              const int i = order[t];
              const float dd = d2((*c)[i], qx, qy, qz);
              if ((double)dd > bound) continue;
              if (!found || dd < bd || (dd == bd && i < bi)) {
                found = true;
                bd = dd;
                bi = i;
              }
            }
          }
        }
    }
    idx = bi;
    dist = bd;
    return found;
  }
};

// ===========================================================================
// a3  downSample -> pcl::VoxelGrid<PointXYZRGB>
// [REF src/features.cpp:17-27] [PCL-recall pcl/filters/impl/voxel_grid.hpp,
//  pcl/common/impl/accumulators.hpp (AccumulatorXYZ, AccumulatorRGBA)]
// ===========================================================================
struct VoxelInfo {
  int min_b[3];
  int div_b[3];
  int passthrough;  // 1 = overflow guard hit, input returned unchanged
};

static Cloud voxel_grid(const Cloud& in, float leaf, VoxelInfo* info = nullptr, std::vector<uint32_t>* keys_out = nullptr)
{
  VoxelInfo vi;
  memset(&vi, 0, sizeof(vi));
  Cloud out;
  if (in.empty()) {
    if (info) *info = vi;
    return out;
  }
  const float inv = 1.0f / leaf;  // Eigen::Array4f::Ones() / leaf_size_.array()
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  // Every input is treated as !is_dense: getMinMax3D and the filter loop skip points with a non-finite coordinate
  // [PCL-recall pcl/common/impl/common.hpp, pcl/filters/impl/voxel_grid.hpp].
  auto finite = [](const P4& p) { return std::isfinite(p.x) && std::isfinite(p.y) && std::isfinite(p.z); };
  size_t n_finite = 0;
  for (const P4& p : in) {  // getMinMax3D
    if (!finite(p)) continue;
    ++n_finite;
    mn[0] = std::min(mn[0], p.x); mx[0] = std::max(mx[0], p.x);
    mn[1] = std::min(mn[1], p.y); mx[1] = std::max(mx[1], p.y);
    mn[2] = std::min(mn[2], p.z); mx[2] = std::max(mx[2], p.z);
  }
  if (n_finite == 0) {
    if (info) *info = vi;
    return out;
  }
  bool pass = !(leaf > 0.0f);
  if (!pass) {
    const int64_t dx = (int64_t)((mx[0] - mn[0]) * inv) + 1;
    const int64_t dy = (int64_t)((mx[1] - mn[1]) * inv) + 1;
    const int64_t dz = (int64_t)((mx[2] - mn[2]) * inv) + 1;
    if (dx * dy * dz > (int64_t)INT32_MAX) pass = true;  // "Leaf size is too small" -> output = input
  }
  if (pass) {
    vi.passthrough = 1;
    if (info) *info = vi;
    return in;
  }
  int min_b[3], div_b[3];
  for (int a = 0; a < 3; ++a) {
    min_b[a] = (int)std::floor(mn[a] * inv);
    const int max_b = (int)std::floor(mx[a] * inv);
    div_b[a] = max_b - min_b[a] + 1;
    vi.min_b[a] = min_b[a];
    vi.div_b[a] = div_b[a];
  }
  const int mul1 = div_b[0], mul2 = div_b[0] * div_b[1];
  std::vector<std::pair<uint32_t, uint32_t>> iv;
  iv.reserve(n_finite);
  for (size_t i = 0; i < in.size(); ++i) {
    if (!finite(in[i])) continue;
    const int i0 = (int)(std::floor(in[i].x * inv) - (float)min_b[0]);
    const int i1 = (int)(std::floor(in[i].y * inv) - (float)min_b[1]);
    const int i2 = (int)(std::floor(in[i].z * inv) - (float)min_b[2]);
    iv.emplace_back((uint32_t)(i0 + i1 * mul1 + i2 * mul2), (uint32_t)i);
  }
  std::sort(iv.begin(), iv.end());  // canonical: ties by ascending point index
  size_t k = 0;
  while (k < iv.size()) {
    size_t e = k + 1;
    while (e < iv.size() && iv[e].first == iv[k].first) ++e;
    float sx = 0.f, sy = 0.f, sz = 0.f, sr = 0.f, sg = 0.f, sb = 0.f, sa = 0.f;
    for (size_t t = k; t < e; ++t) {
      const P4& p = in[iv[t].second];
      sx += p.x; sy += p.y; sz += p.z;
      sr += (float)((p.rgba >> 16) & 0xff);
      sg += (float)((p.rgba >> 8) & 0xff);
      sb += (float)(p.rgba & 0xff);
      sa += (float)((p.rgba >> 24) & 0xff);
    }
    const float n = (float)(e - k);
    P4 o;
    o.x = sx / n; o.y = sy / n; o.z = sz / n;
    o.rgba = ((uint32_t)(sa / n) << 24) | ((uint32_t)(sr / n) << 16) | ((uint32_t)(sg / n) << 8) | (uint32_t)(sb / n);
    out.push_back(o);
    if (keys_out) keys_out->push_back(iv[k].first);
    k = e;
  }
  if (info) *info = vi;
  return out;
}

// ===========================================================================
// a4  removeOutliers -> pcl::RadiusOutlierRemoval  [REF src/features.cpp:31-43]
// [PCL-recall pcl/filters/impl/radius_outlier_removal.hpp, 1.8.1 radiusSearch]
// ===========================================================================
static Cloud radius_outlier_removal(const Cloud& in, double radius, int min_nb, std::vector<int>* kept, std::vector<int>* counts)
{
  Grid g;
  g.build(in, (float)radius);
  const float r2 = (float)(radius * radius);
  Cloud out;
  if (kept) kept->clear();
  if (counts) counts->assign(in.size(), 0);
  std::vector<int> cnt(in.size(), 0);
#pragma omp parallel for schedule(dynamic, 512)
  for (long long i = 0; i < (long long)in.size(); ++i) {
    int k = 0;
    g.for_radius(in[i].x, in[i].y, in[i].z, (float)radius, r2, [&](int, float) { ++k; });
    cnt[i] = k;
  }
  for (size_t i = 0; i < in.size(); ++i) {
    const int k = cnt[i];
    if (counts) (*counts)[i] = k;
    if (k <= min_nb) continue;  // "k <= min_pts_radius_" -> outlier
    out.push_back(in[i]);
    if (kept) kept->push_back((int)i);
  }
  return out;
}

// ===========================================================================
// pcl::eigen33 / computeRoots, restated [PCL-recall pcl/common/impl/eigen.hpp]
// ===========================================================================
static void computeRoots2(float b, float c, float* roots)
{
  roots[0] = 0.0f;
  float d = (float)(b * b - 4.0 * c);
  if (d < 0.0) d = 0.0;
  const float sd = std::sqrt(d);
  roots[2] = 0.5f * (b + sd);
  roots[1] = 0.5f * (b - sd);
}

static void computeRoots(const float m[3][3], float* roots)
{
  const float c0 = m[0][0] * m[1][1] * m[2][2] + 2.0f * m[0][1] * m[0][2] * m[1][2] - m[0][0] * m[1][2] * m[1][2] -
                   m[1][1] * m[0][2] * m[0][2] - m[2][2] * m[0][1] * m[0][1];
  const float c1 = m[0][0] * m[1][1] - m[0][1] * m[0][1] + m[0][0] * m[2][2] - m[0][2] * m[0][2] + m[1][1] * m[2][2] -
                   m[1][2] * m[1][2];
  const float c2 = m[0][0] + m[1][1] + m[2][2];
  if (std::fabs(c0) < FLT_EPSILON) {
    computeRoots2(c2, c1, roots);
    return;
  }
  const float s_inv3 = (float)(1.0 / 3.0);
  const float s_sqrt3 = std::sqrt(3.0f);
  const float c2_over_3 = c2 * s_inv3;
  float a_over_3 = (c1 - c2 * c2_over_3) * s_inv3;
  if (a_over_3 > 0.0f) a_over_3 = 0.0f;
  const float half_b = 0.5f * (c0 + c2_over_3 * (2.0f * c2_over_3 * c2_over_3 - c1));
  float q = half_b * half_b + a_over_3 * a_over_3 * a_over_3;
  if (q > 0.0f) q = 0.0f;
  const float rho = std::sqrt(-a_over_3);
  const float theta = m_atan2f(std::sqrt(-q), half_b) * s_inv3;
  const float cos_theta = m_cosf(theta);
  const float sin_theta = m_sinf(theta);
  roots[0] = c2_over_3 + 2.0f * rho * cos_theta;
  roots[1] = c2_over_3 - rho * (cos_theta + s_sqrt3 * sin_theta);
  roots[2] = c2_over_3 - rho * (cos_theta - s_sqrt3 * sin_theta);
  if (roots[0] >= roots[1]) std::swap(roots[0], roots[1]);
  if (roots[1] >= roots[2]) {
    std::swap(roots[1], roots[2]);
    if (roots[0] >= roots[1]) std::swap(roots[0], roots[1]);
  }
  if (roots[0] <= 0) computeRoots2(c2, c1, roots);
}

static float scaleMat(const float in[3][3], float out[3][3])
{
  float scale = 0.0f;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) scale = std::max(scale, std::fabs(in[i][j]));
  if (scale <= FLT_MIN) scale = 1.0f;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) out[i][j] = in[i][j] / scale;
  return scale;
}

static void eigen33_values(const float mat[3][3], float* evals)
{
  float sm[3][3];
  const float scale = scaleMat(mat, sm);
  computeRoots(sm, evals);
  for (int i = 0; i < 3; ++i) evals[i] *= scale;
}

static void cross3(const float* a, const float* b, float* o)
{
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}

static void eigen33_smallest(const float mat[3][3], float& eigenvalue, float* eigenvector)
{
  float sm[3][3];
  const float scale = scaleMat(mat, sm);
  float ev[3];
  computeRoots(sm, ev);
  eigenvalue = ev[0] * scale;
  sm[0][0] -= ev[0];
  sm[1][1] -= ev[0];
  sm[2][2] -= ev[0];
  float v1[3], v2[3], v3[3];
  cross3(sm[0], sm[1], v1);
  cross3(sm[0], sm[2], v2);
  cross3(sm[1], sm[2], v3);
  const float len1 = v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2];
  const float len2 = v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2];
  const float len3 = v3[0] * v3[0] + v3[1] * v3[1] + v3[2] * v3[2];
  const float* v;
  float len;
  if (len1 >= len2 && len1 >= len3) { v = v1; len = len1; }
  else if (len2 >= len1 && len2 >= len3) { v = v2; len = len2; }
  else { v = v3; len = len3; }
  const float s = std::sqrt(len);
  for (int i = 0; i < 3; ++i) eigenvector[i] = v[i] / s;
}

// pcl::computeMeanAndCovarianceMatrix (float, single pass) [PCL-recall pcl/common/impl/centroid.hpp]
static void mean_and_cov(const Cloud& cl, const std::vector<int>& idx, float cov[3][3], float* centroid)
{
  float a[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int i : idx) {
    const P4& p = cl[i];
    a[0] += p.x * p.x; a[1] += p.x * p.y; a[2] += p.x * p.z;
    a[3] += p.y * p.y; a[4] += p.y * p.z; a[5] += p.z * p.z;
    a[6] += p.x; a[7] += p.y; a[8] += p.z;
  }
  const float n = (float)idx.size();
  for (int k = 0; k < 9; ++k) a[k] /= n;
  centroid[0] = a[6]; centroid[1] = a[7]; centroid[2] = a[8];
  cov[0][0] = a[0] - a[6] * a[6];
  cov[0][1] = a[1] - a[6] * a[7];
  cov[0][2] = a[2] - a[6] * a[8];
  cov[1][1] = a[3] - a[7] * a[7];
  cov[1][2] = a[4] - a[7] * a[8];
  cov[2][2] = a[5] - a[8] * a[8];
  cov[1][0] = cov[0][1]; cov[2][0] = cov[0][2]; cov[2][1] = cov[1][2];
}

// ===========================================================================
// a5  computeSurfaceNormals -> pcl::NormalEstimation  [REF src/features.cpp:168-179]
// [PCL-recall pcl/features/impl/normal_3d.hpp, normal_3d.h]
// ===========================================================================
static Normals surface_normals(const Cloud& in, double radius)
{
  Grid g;
  g.build(in, (float)radius);
  Normals out(in.size());
  const float nanv = std::numeric_limits<float>::quiet_NaN();
#pragma omp parallel
  {
  std::vector<int> idx;
  std::vector<float> dist;
#pragma omp for schedule(dynamic, 512)
  for (long long i = 0; i < (long long)in.size(); ++i) {
    g.radius_sorted(in[i].x, in[i].y, in[i].z, radius, idx, dist);
    if (idx.size() < 3) {
      out[i].nx = out[i].ny = out[i].nz = out[i].curv = nanv;
      continue;
    }
    float cov[3][3], cen[3];
    mean_and_cov(in, idx, cov, cen);
    float ev, vec[3];
    eigen33_smallest(cov, ev, vec);  // solvePlaneParameters
    float nx = vec[0], ny = vec[1], nz = vec[2];
    const float eig_sum = cov[0][0] + cov[1][1] + cov[2][2];
    const float curv = (eig_sum != 0) ? std::fabs(ev / eig_sum) : 0.0f;
    // flipNormalTowardsViewpoint, viewpoint (0,0,0)
    const float vx = 0.0f - in[i].x, vy = 0.0f - in[i].y, vz = 0.0f - in[i].z;
    const float cos_theta = (vx * nx + vy * ny + vz * nz);
    if (cos_theta < 0) { nx *= -1; ny *= -1; nz *= -1; }
    out[i].nx = nx; out[i].ny = ny; out[i].nz = nz; out[i].curv = curv;
  }
  }
  return out;
}

// ===========================================================================
// a6  detectKeypoints(SIFT) -> pcl::SIFTKeypoint<PointXYZRGB, PointWithScale>
// [REF src/features.cpp:45-62,85-96] [PCL-recall pcl/keypoints/impl/sift_keypoint.hpp]
// order_mode 0: the Gaussian-weighted sums are taken in 2^-32 fixed point (int64) — order-free, the exactly rounded
//               value of the sums PCL accumulates in float (canonical, what the CUDA path does)
// order_mode 1: float summation, nearest neighbour first, with PCL's early break (literal)
// ===========================================================================
static inline float sift_intensity(const P4& p)
{
  const int r = (p.rgba >> 16) & 0xff, g = (p.rgba >> 8) & 0xff, b = p.rgba & 0xff;
  return (float)(299 * r + 587 * g + 114 * b) / 1000.0f;
}

struct SiftDebug {
  std::vector<float> dog;       // octave 0, N x 5
  std::vector<int> octave_sizes;
};

static Cloud sift_keypoints(const Cloud& input, float min_scale, int nr_octaves, int nr_scales_per_octave, float min_contrast,
                            int order_mode, SiftDebug* dbg = nullptr, std::vector<float>* scales_out = nullptr)
{
  Cloud output;
  Cloud cloud = input;
  float scale = min_scale;
  for (int i_octave = 0; i_octave < nr_octaves; ++i_octave) {
    const float s = 1.0f * scale;
    cloud = voxel_grid(cloud, s);
    if (dbg) dbg->octave_sizes.push_back((int)cloud.size());
    const size_t min_nr_points = 25;
    if (cloud.size() < min_nr_points) break;
    Grid tree;
    tree.build(cloud, 2.0f * s);
    // detectKeypointsForOctave
    const int nscales = nr_scales_per_octave + 3;
    std::vector<float> scales(nscales);
    for (int i_scale = 0; i_scale <= nr_scales_per_octave + 2; ++i_scale)
      scales[i_scale] = scale * powf(2.0f, (1.0f * (float)i_scale - 1.0f) / (float)nr_scales_per_octave);
    // computeScaleSpace
    const size_t n = cloud.size();
    const int nd = nscales - 1;
    std::vector<float> dog(n * nd);
    const float max_radius = 3.0f * scales.back();
    std::vector<float> sigma_sqr(nscales);
    for (int i = 0; i < nscales; ++i) sigma_sqr[i] = powf(scales[i], 2.0f);
#pragma omp parallel
    {
    std::vector<int> nn_idx;
    std::vector<float> nn_dist;
#pragma omp for schedule(dynamic, 256)
    for (long long ip = 0; ip < (long long)n; ++ip) {
      tree.radius_sorted(cloud[ip].x, cloud[ip].y, cloud[ip].z, (double)max_radius, nn_idx, nn_dist);
      if (order_mode == 1) {
        std::vector<std::pair<float, int>> tmp(nn_idx.size());
        for (size_t k = 0; k < nn_idx.size(); ++k) tmp[k] = std::make_pair(nn_dist[k], nn_idx[k]);
        std::sort(tmp.begin(), tmp.end());
        for (size_t k = 0; k < tmp.size(); ++k) { nn_dist[k] = tmp[k].first; nn_idx[k] = tmp[k].second; }
      }
      float filter_response = 0.0f, previous_filter_response;
      for (int i_scale = 0; i_scale < nscales; ++i_scale) {
        const float ss = sigma_sqr[i_scale];
        float numerator = 0.0f, denominator = 0.0f;
        long long num_fix = 0, den_fix = 0;
        for (size_t k = 0; k < nn_idx.size(); ++k) {
          const float value = sift_intensity(cloud[nn_idx[k]]);
          const float dist_sqr = nn_dist[k];
          if (dist_sqr <= 9 * ss) {
            const float w = m_expf(-0.5f * dist_sqr / ss);
            if (order_mode == 1) {
              numerator += value * w;
              denominator += w;
            } else {
              num_fix += mm3d::em::to_fix32_pos(value * w);
              den_fix += mm3d::em::to_fix32_pos(w);
            }
          } else if (order_mode == 1) {
            break;  // sorted results: everything after is farther
          }
        }
        if (order_mode != 1) {
          numerator = (float)((double)num_fix / MM3D_FIX1_SCALE);
          denominator = (float)((double)den_fix / MM3D_FIX1_SCALE);
        }
        previous_filter_response = filter_response;
        filter_response = numerator / denominator;
        if (i_scale > 0) dog[ip * nd + (i_scale - 1)] = filter_response - previous_filter_response;
      }
    }
    }
    if (dbg && i_octave == 0) dbg->dog = dog;
    // findScaleSpaceExtrema (the per-point min / max tables in parallel, the emission loop below in point order)
    const int k = 25;
    std::vector<float> all_min(n * nd), all_max(n * nd);
#pragma omp parallel
    {
    std::vector<std::pair<float, int>> nn;
#pragma omp for schedule(dynamic, 256)
    for (long long ip = 0; ip < (long long)n; ++ip) {
      tree.knn(cloud[ip].x, cloud[ip].y, cloud[ip].z, k, nn);
      for (int is = 0; is < nd; ++is) {
        float mnv = FLT_MAX, mxv = -FLT_MAX;
        for (size_t t = 0; t < nn.size(); ++t) {
          const float d = dog[(size_t)nn[t].second * nd + is];
          mnv = std::min(mnv, d);
          mxv = std::max(mxv, d);
        }
        all_min[ip * nd + is] = mnv;
        all_max[ip * nd + is] = mxv;
      }
    }
    }
    for (size_t ip = 0; ip < n; ++ip) {
      const float* min_val = &all_min[ip * nd];
      const float* max_val = &all_max[ip * nd];
      for (int is = 1; is < nd - 1; ++is) {
        const float val = dog[ip * nd + is];
        if (std::fabs(val) >= min_contrast) {
          bool emit = false;
          if ((val == min_val[is]) && (val < min_val[is - 1]) && (val < min_val[is + 1])) emit = true;
          else if ((val == max_val[is]) && (val > max_val[is - 1]) && (val > max_val[is + 1])) emit = true;
          if (emit) {
            P4 kp;
            kp.x = cloud[ip].x; kp.y = cloud[ip].y; kp.z = cloud[ip].z;
            kp.rgba = 0xff000000u;  // copyPointCloud -> default-constructed PointXYZRGB colour
            output.push_back(kp);
            if (scales_out) scales_out->push_back(scales[is]);
          }
        }
      }
    }
    scale *= 2;
  }
  return output;
}

// ===========================================================================
// a7  detectKeypoints(HARRIS) -> pcl::HarrisKeypoint3D<PointXYZRGB, PointXYZI>
// [REF src/features.cpp:64-83: setNonMaxSupression(true), setRefine(true), setThreshold(float), setRadius(float)]
// [PCL-recall pcl/keypoints/impl/harris_3d.hpp (SSE branch of calculateNormalCovar), pcl/common/impl/eigen.hpp
//  invert3x3SymMatrix].  PCL's OpenMP loops make the output order nondeterministic; canonical = ascending index.
// ===========================================================================
static Cloud harris_keypoints(const Cloud& input, const Normals& normals, float threshold, float radius_f,
                              std::vector<float>* response_dbg = nullptr, Cloud* unrefined_dbg = nullptr)
{
  const double search_radius = (double)radius_f;  // search_radius_ = float radius (setRadius(float))
  Grid tree;
  tree.build(input, radius_f);
  const size_t n = input.size();
  std::vector<float> response(n, 0.0f);
  std::vector<int> nn_idx;
  std::vector<float> nn_dist;
  // responseHarris
  for (size_t p = 0; p < n; ++p) {
    tree.radius_sorted(input[p].x, input[p].y, input[p].z, search_radius, nn_idx, nn_dist);
    // calculateNormalCovar: xx xy xz | yy yz | zz
    float c0 = 0.f, c1 = 0.f, c2 = 0.f, c5 = 0.f, c6 = 0.f, c7 = 0.f;
    unsigned count = 0;
    for (int q : nn_idx) {
      const N4& nq = normals[q];
      if (!std::isfinite(nq.nx)) continue;
      c0 += nq.nx * nq.nx; c1 += nq.nx * nq.ny; c2 += nq.nx * nq.nz;
      c5 += nq.ny * nq.ny; c6 += nq.ny * nq.nz;
      c7 += nq.nz * nq.nz;
      ++count;
    }
    if (count > 0) {
      const float cn = (float)count;
      c0 /= cn; c1 /= cn; c2 /= cn; c5 /= cn; c6 /= cn; c7 /= cn;
    } else {
      c0 = c1 = c2 = c5 = c6 = c7 = 0.f;
    }
    const float trace = c0 + c5 + c7;
    if (trace != 0) {
      const float det = c0 * c5 * c7 + 2.0f * c1 * c2 * c6 - c2 * c2 * c5 - c1 * c1 * c7 - c6 * c6 * c0;
      response[p] = 0.04f + det - 0.04f * trace * trace;
    }
  }
  if (response_dbg) *response_dbg = response;
  // non-maximum suppression
  Cloud corners;
  for (size_t p = 0; p < n; ++p) {
    if (!std::isfinite(response[p]) || response[p] < threshold) continue;
    tree.radius_sorted(input[p].x, input[p].y, input[p].z, search_radius, nn_idx, nn_dist);
    bool is_maxima = true;
    for (int q : nn_idx)
      if (response[p] < response[q]) { is_maxima = false; break; }
    if (is_maxima) {
      P4 c = input[p];
      c.rgba = 0xff000000u;
      corners.push_back(c);
    }
  }
  if (unrefined_dbg) *unrefined_dbg = corners;
  // refineCorners
  for (P4& corner_out : corners) {
    unsigned iterations = 0;
    float diff;
    do {
      float NNT[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, NNTp[3] = {0, 0, 0};
      const float cx = corner_out.x, cy = corner_out.y, cz = corner_out.z;
      tree.radius_sorted(cx, cy, cz, search_radius, nn_idx, nn_dist);
      for (int q : nn_idx) {
        const N4& nq = normals[q];
        if (!std::isfinite(nq.nx)) continue;
        const float nv[3] = {nq.nx, nq.ny, nq.nz};
        float nnT[9];
        for (int r = 0; r < 3; ++r)
          for (int c = 0; c < 3; ++c) nnT[r * 3 + c] = nv[r] * nv[c];
        for (int k = 0; k < 9; ++k) NNT[k] += nnT[k];
        const float pv[3] = {input[q].x, input[q].y, input[q].z};
        for (int r = 0; r < 3; ++r) NNTp[r] += (nnT[r * 3 + 0] * pv[0] + nnT[r * 3 + 1] * pv[1]) + nnT[r * 3 + 2] * pv[2];
      }
      // invert3x3SymMatrix (column-major coeff(i) of a symmetric matrix == row-major)
      const float fd_ee = NNT[4] * NNT[8] - NNT[7] * NNT[5];
      const float ce_bf = NNT[2] * NNT[5] - NNT[1] * NNT[8];
      const float be_cd = NNT[1] * NNT[5] - NNT[2] * NNT[4];
      const float det = NNT[0] * fd_ee + NNT[1] * ce_bf + NNT[2] * be_cd;
      if (det != 0) {
        float inv[9];
        inv[0] = fd_ee;
        inv[1] = inv[3] = ce_bf;
        inv[2] = inv[6] = be_cd;
        inv[4] = (NNT[0] * NNT[8] - NNT[2] * NNT[2]);
        inv[5] = inv[7] = (NNT[1] * NNT[2] - NNT[0] * NNT[5]);
        inv[8] = (NNT[0] * NNT[4] - NNT[1] * NNT[1]);
        for (int k = 0; k < 9; ++k) inv[k] /= det;
        corner_out.x = (inv[0] * NNTp[0] + inv[1] * NNTp[1]) + inv[2] * NNTp[2];
        corner_out.y = (inv[3] * NNTp[0] + inv[4] * NNTp[1]) + inv[5] * NNTp[2];
        corner_out.z = (inv[6] * NNTp[0] + inv[7] * NNTp[1]) + inv[8] * NNTp[2];
      }
      const float dx = corner_out.x - cx, dy = corner_out.y - cy, dz = corner_out.z - cz;
      diff = (dx * dx + dy * dy) + dz * dz;
    } while (diff > 1e-6 && ++iterations < 10);
  }
  return corners;
}

// ===========================================================================
// a8-FPFH  computeLocalDescriptors(FPFH) -> pcl::FPFHEstimation
// [REF src/features.cpp:99-150, src/dispatch_descriptors.h:40]
// [PCL-recall pcl/features/impl/fpfh.hpp, pcl/features/impl/pfh.hpp computePairFeatures]
// ===========================================================================
// returns false for a degenerate pair (coincident points, or normal parallel to the connecting line); the callers skip it
// ("if (!computePairFeatures (...)) continue;" in computePointSPFHSignature / computePointPFHSignature)
static bool pair_features(const P4& p1, const N4& n1, const P4& p2, const N4& n2, float& f1, float& f2, float& f3, float& f4)
{
  float dp[3] = {p2.x - p1.x, p2.y - p1.y, p2.z - p1.z};
  f4 = std::sqrt((dp[0] * dp[0] + dp[1] * dp[1]) + dp[2] * dp[2]);
  if (f4 == 0.0f) { f1 = f2 = f3 = f4 = 0.0f; return false; }
  float a[3] = {n1.nx, n1.ny, n1.nz}, b[3] = {n2.nx, n2.ny, n2.nz};
  const float angle1 = ((a[0] * dp[0] + a[1] * dp[1]) + a[2] * dp[2]) / f4;
  const float angle2 = ((b[0] * dp[0] + b[1] * dp[1]) + b[2] * dp[2]) / f4;
  // acos(fabs(angle1)) > acos(fabs(angle2)); acos is strictly decreasing on [0,1]
  // and NaN outside, so this is |a1| < |a2| with both in range.
  const float fa1 = std::fabs(angle1), fa2 = std::fabs(angle2);
#ifdef ORACLE_LIBM
  const bool sw = std::acos((double)fa1) > std::acos((double)fa2);
#else
  const bool sw = (fa1 <= 1.0f) && (fa2 <= 1.0f) && (fa1 < fa2);
#endif
  if (sw) {
    for (int i = 0; i < 3; ++i) { std::swap(a[i], b[i]); dp[i] *= (-1); }
    f3 = -angle2;
  } else {
    f3 = angle1;
  }
  float v[3];
  cross3(dp, a, v);
  const float v_norm = std::sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]);
  if (v_norm == 0.0f) { f1 = f2 = f3 = f4 = 0.0f; return false; }
  for (int i = 0; i < 3; ++i) v[i] /= v_norm;
  float w[3];
  cross3(a, v, w);
  f2 = (v[0] * b[0] + v[1] * b[1]) + v[2] * b[2];
  f1 = m_atan2f((w[0] * b[0] + w[1] * b[1]) + w[2] * b[2], (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]);
  return true;
}

// pcl::computeRGBPairFeatures [PCL-recall pcl/features/impl/pfhrgb.hpp + pfh_tools]: the Darboux frame WITHOUT the
// source/target swap of computePairFeatures, plus three colour ratios taken with INTEGER division and folded into [-1, 1].
static bool rgb_pair_features(const P4& p1, const N4& n1, const P4& p2, const N4& n2, float f[7])
{
  for (int i = 0; i < 7; ++i) f[i] = 0.0f;
  const float dp[3] = {p2.x - p1.x, p2.y - p1.y, p2.z - p1.z};
  const float f4 = std::sqrt((dp[0] * dp[0] + dp[1] * dp[1]) + dp[2] * dp[2]);
  if (f4 == 0.0f) return false;
  const float a[3] = {n1.nx, n1.ny, n1.nz}, b[3] = {n2.nx, n2.ny, n2.nz};
  const float angle1 = ((a[0] * dp[0] + a[1] * dp[1]) + a[2] * dp[2]) / f4;
  float v[3];
  cross3(dp, a, v);
  const float v_norm = std::sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]);
  if (v_norm == 0.0f) return false;
  for (int i = 0; i < 3; ++i) v[i] /= v_norm;
  float w[3];
  cross3(a, v, w);
  f[3] = f4;
  f[2] = angle1;
  f[1] = (v[0] * b[0] + v[1] * b[1]) + v[2] * b[2];
  f[0] = m_atan2f((w[0] * b[0] + w[1] * b[1]) + w[2] * b[2], (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]);
  const int c1[3] = {(int)((p1.rgba >> 16) & 0xff), (int)((p1.rgba >> 8) & 0xff), (int)(p1.rgba & 0xff)};
  const int c2[3] = {(int)((p2.rgba >> 16) & 0xff), (int)((p2.rgba >> 8) & 0xff), (int)(p2.rgba & 0xff)};
  for (int i = 0; i < 3; ++i) {
    float r = (c2[i] != 0) ? (float)(c1[i] / c2[i]) : 1.0f;
    if (r > 1.0f) r = -1.0f / r;
    f[4 + i] = r;
  }
  return true;
}

static inline int clampbin(int h, int nb)
{
  if (h < 0) h = 0;
  if (h >= nb) h = nb - 1;
  return h;
}

// returns K' x 33 descriptors; keypoints is filtered in place (features.cpp:119-141)
static std::vector<float> fpfh_descriptors(const Cloud& surface, const Normals& normals, Cloud& keypoints, double radius,
                                           std::vector<float>* spfh_dbg = nullptr)
{
  const int NB = 11;
  Grid tree;
  tree.build(surface, (float)radius);
  const size_t N = surface.size(), K = keypoints.size();
  std::vector<int> nn_idx;
  std::vector<float> nn_dist;
  // computeSPFHSignatures: union of the keypoints' neighbours
  std::vector<char> need(N, 0);
#pragma omp parallel
  {
  std::vector<int> nn_idx;
  std::vector<float> nn_dist;
#pragma omp for schedule(dynamic, 64)
  for (long long k = 0; k < (long long)K; ++k) {
    tree.radius_sorted(keypoints[k].x, keypoints[k].y, keypoints[k].z, radius, nn_idx, nn_dist);
    for (int i : nn_idx) need[i] = 1;  // every writer stores the same value
  }
  }
  std::vector<float> hist(N * 33, 0.0f);  // rows at the surface index (lookup = identity on needed rows)
  const float d_pi = 1.0f / (2.0f * (float)M_PI);
#pragma omp parallel
  {
  std::vector<int> nn_idx;
  std::vector<float> nn_dist;
#pragma omp for schedule(dynamic, 256)
  for (long long p = 0; p < (long long)N; ++p) {
    if (!need[p]) continue;
    tree.radius_sorted(surface[p].x, surface[p].y, surface[p].z, radius, nn_idx, nn_dist);
    if (nn_idx.empty()) continue;
    const float hist_incr = 100.0f / (float)(nn_idx.size() - 1);
    float* h = &hist[p * 33];
    for (int q : nn_idx) {
      if ((int)p == q) continue;
      float f1, f2, f3, f4;
      if (!pair_features(surface[p], normals[p], surface[q], normals[q], f1, f2, f3, f4)) continue;
      int hi = (int)std::floor(NB * ((f1 + M_PI) * d_pi));
      h[clampbin(hi, NB)] += hist_incr;
      hi = (int)std::floor(NB * ((f2 + 1.0) * 0.5));
      h[NB + clampbin(hi, NB)] += hist_incr;
      hi = (int)std::floor(NB * ((f3 + 1.0) * 0.5));
      h[2 * NB + clampbin(hi, NB)] += hist_incr;
    }
  }
  }
  if (spfh_dbg) *spfh_dbg = hist;
  std::vector<float> desc;
  Cloud kept;
  for (size_t k = 0; k < K; ++k) {
    tree.radius_sorted(keypoints[k].x, keypoints[k].y, keypoints[k].z, radius, nn_idx, nn_dist);
    if (nn_idx.empty()) continue;  // NaN row -> dropped by the reference
    float f[33];
    for (int i = 0; i < 33; ++i) f[i] = 0.0f;
    double sum[3] = {0.0, 0.0, 0.0};
    for (size_t t = 0; t < nn_idx.size(); ++t) {
      if (nn_dist[t] == 0) continue;
      const float weight = 1.0f / nn_dist[t];
      const float* h = &hist[(size_t)nn_idx[t] * 33];
      for (int b = 0; b < 3; ++b)
        for (int i = 0; i < NB; ++i) {
          const float val = h[b * NB + i] * weight;
          sum[b] += val;
          f[b * NB + i] += val;
        }
    }
    bool finite = true;
    for (int b = 0; b < 3; ++b) {
      if (sum[b] != 0) sum[b] = 100.0 / sum[b];
      for (int i = 0; i < NB; ++i) {
        f[b * NB + i] *= (float)sum[b];
        if (!std::isfinite(f[b * NB + i])) finite = false;
      }
    }
    if (!finite) continue;  // DefaultPointRepresentation::isValid
    desc.insert(desc.end(), f, f + 33);
    kept.push_back(keypoints[k]);
  }
  keypoints.swap(kept);
  return desc;
}

// ===========================================================================
// a8-PFH  computeLocalDescriptors(PFH) -> pcl::PFHEstimation<PointXYZRGB, Normal, PFHSignature125>
// (the DEFAULT descriptor_type, map_merging.h:35) [REF src/dispatch_descriptors.h:38]
// [PCL-recall pcl/features/impl/pfh.hpp computePointPFHSignature: all pairs (i, j < i) of the neighbourhood,
//  5 x 5 x 5 bins over (f1, f2, f3), every pair adds 100 / (n (n-1) / 2)]
// ===========================================================================
static std::vector<float> pfh_descriptors(const Cloud& surface, const Normals& normals, Cloud& keypoints, double radius)
{
  const int nr_split = 5;
  Grid tree;
  tree.build(surface, (float)radius);
  const float d_pi = 1.0f / (2.0f * (float)M_PI);
  std::vector<int> idx;
  std::vector<float> sqd;
  std::vector<float> desc;
  Cloud kept;
  for (size_t k = 0; k < keypoints.size(); ++k) {
    tree.radius_sorted(keypoints[k].x, keypoints[k].y, keypoints[k].z, radius, idx, sqd);
    if (idx.empty()) continue;  // NaN histogram -> dropped
    float h[125];
    for (float& v : h) v = 0.f;
    const float hist_incr = 100.0f / (float)(idx.size() * (idx.size() - 1) / 2);
    for (size_t i = 0; i < idx.size(); ++i)
      for (size_t j = 0; j < i; ++j) {
        float f1, f2, f3, f4;
        if (!pair_features(surface[idx[i]], normals[idx[i]], surface[idx[j]], normals[idx[j]], f1, f2, f3, f4)) continue;
        int fi[3];
        fi[0] = (int)std::floor(nr_split * ((f1 + M_PI) * d_pi));
        fi[1] = (int)std::floor(nr_split * ((f2 + 1.0) * 0.5));
        fi[2] = (int)std::floor(nr_split * ((f3 + 1.0) * 0.5));
        int h_index = 0, h_p = 1;
        for (int d = 0; d < 3; ++d) {
          h_index += h_p * clampbin(fi[d], nr_split);
          h_p *= nr_split;
        }
        h[h_index] += hist_incr;
      }
    bool finite = true;
    for (float v : h)
      if (!std::isfinite(v)) finite = false;
    if (!finite) continue;
    desc.insert(desc.end(), h, h + 125);
    kept.push_back(keypoints[k]);
  }
  keypoints.swap(kept);
  return desc;
}

// ===========================================================================
// a8-PFHRGB  computeLocalDescriptors(PFHRGB) -> pcl::PFHRGBEstimation<PointXYZRGB, Normal, PFHRGBSignature250>
// [REF src/dispatch_descriptors.h:39] [PCL-recall pcl/features/impl/pfhrgb.hpp computePointPFHRGBSignature: all pairs
//  (i, j < i); 5^3 bins over (f1, f2, f3) in [0, 125) and 5^3 bins over the three colour ratios in [125, 250); every
//  surviving pair adds 100 / (n (n-1) / 2) to one bin of each half; no NaN marking — an empty neighbourhood leaves zeros]
// ===========================================================================
static std::vector<float> pfhrgb_descriptors(const Cloud& surface, const Normals& normals, Cloud& keypoints, double radius)
{
  const int nr_split = 5;
  Grid tree;
  tree.build(surface, (float)radius);
  const float d_pi = 1.0f / (2.0f * (float)M_PI);
  std::vector<int> idx;
  std::vector<float> sqd;
  std::vector<float> desc;
  Cloud kept;
  for (size_t k = 0; k < keypoints.size(); ++k) {
    tree.radius_sorted(keypoints[k].x, keypoints[k].y, keypoints[k].z, radius, idx, sqd);
    float h[250];
    for (float& v : h) v = 0.f;
    const float hist_incr = 100.0f / (float)(idx.size() * (idx.size() - 1) / 2);  // inf for n < 2, never added
    for (size_t i = 0; i < idx.size(); ++i)
      for (size_t j = 0; j < i; ++j) {
        float f[7];
        if (!rgb_pair_features(surface[idx[i]], normals[idx[i]], surface[idx[j]], normals[idx[j]], f)) continue;
        int fi[7];
        fi[0] = clampbin((int)std::floor(nr_split * ((f[0] + M_PI) * d_pi)), nr_split);
        fi[1] = clampbin((int)std::floor(nr_split * ((f[1] + 1.0) * 0.5)), nr_split);
        fi[2] = clampbin((int)std::floor(nr_split * ((f[2] + 1.0) * 0.5)), nr_split);
        fi[4] = clampbin((int)std::floor(nr_split * ((f[4] + 1.0) * 0.5)), nr_split);
        fi[5] = clampbin((int)std::floor(nr_split * ((f[5] + 1.0) * 0.5)), nr_split);
        fi[6] = clampbin((int)std::floor(nr_split * ((f[6] + 1.0) * 0.5)), nr_split);
        h[fi[0] + 5 * fi[1] + 25 * fi[2]] += hist_incr;
        h[125 + fi[4] + 5 * fi[5] + 25 * fi[6]] += hist_incr;
      }
    bool finite = true;
    for (float v : h)
      if (!std::isfinite(v)) finite = false;
    if (!finite) continue;
    desc.insert(desc.end(), h, h + 250);
    kept.push_back(keypoints[k]);
  }
  keypoints.swap(kept);
  return desc;
}

// ===========================================================================
// a8-RSD  computeLocalDescriptors(RSD) -> pcl::RSDEstimation<PointXYZRGB, Normal, PrincipalRadiiRSD>
// [REF src/dispatch_descriptors.h:43; the matcher sees 2 floats (r_min, r_max): DefaultPointRepresentation = sizeof / 4]
// [PCL-recall pcl/features/impl/rsd.hpp computeRSD (surface, normals, indices, max_dist, nr_subdiv, plane_radius, radii):
//  nr_subdiv_ = 5, plane_radius_ = 0.2; the centre of the patch is indices[0], i.e. the NEAREST surface point of the
//  keypoint (sorted radius search); for every other neighbour the angle between the normals (orientation ignored) and the
//  distance to the centre go into per-distance-bin min / max angles; radii from the least-squares lines through them.]
// Canonical choices: nearest = smallest (d^2, index); sqrt of the float distance sum taken in double (unqualified C sqrt);
// a neighbour at exactly max_dist would index bin nr_subdiv in PCL (one past the end) — it goes to the last bin here.
// ===========================================================================
static std::vector<float> rsd_descriptors(const Cloud& surface, const Normals& normals, Cloud& keypoints, double radius)
{
  const int nr_subdiv = 5;
  const double plane_radius = 0.2, max_dist = radius;
  Grid tree;
  tree.build(surface, (float)radius);
  std::vector<int> idx;
  std::vector<float> sqd;
  std::vector<float> desc;
  Cloud kept;
  for (size_t k = 0; k < keypoints.size(); ++k) {
    tree.radius_sorted(keypoints[k].x, keypoints[k].y, keypoints[k].z, radius, idx, sqd);
    float r_min = 0.0f, r_max = 0.0f;
    if (idx.size() >= 2) {
      size_t b = 0;
      for (size_t i = 1; i < idx.size(); ++i)
        if (sqd[i] < sqd[b]) b = i;  // ascending index: the first minimum is the lowest index
      const int c = idx[b];
      double mn[5], mx[5];
      mn[0] = mx[0] = 0.0;
      for (int d = 1; d < nr_subdiv; ++d) { mn[d] = DBL_MAX; mx[d] = -DBL_MAX; }
      for (size_t i = 0; i < idx.size(); ++i) {
        if (i == b) continue;
        const int q = idx[i];
        double cosine = (normals[q].nx * normals[c].nx + normals[q].ny * normals[c].ny) + normals[q].nz * normals[c].nz;
        if (cosine > 1) cosine = 1;
        if (cosine < -1) cosine = -1;
        double angle = m_acos(cosine);
        if (angle > M_PI / 2) angle = M_PI - angle;
        const float dx = surface[q].x - surface[c].x, dy = surface[q].y - surface[c].y, dz = surface[q].z - surface[c].z;
        const double dist = std::sqrt((double)((dx * dx + dy * dy) + dz * dz));
        if (dist > max_dist) continue;
        int bin_d = (int)std::floor(nr_subdiv * dist / max_dist);
        if (bin_d > nr_subdiv - 1) bin_d = nr_subdiv - 1;
        if (mn[bin_d] > angle) mn[bin_d] = angle;
        if (mx[bin_d] < angle) mx[bin_d] = angle;
      }
      double Amint_Amin = 0, Amint_d = 0, Amaxt_Amax = 0, Amaxt_d = 0;
      for (int d = 0; d < nr_subdiv; ++d) {
        if (mx[d] >= 0) {
          const double f = (d + 0.5) * max_dist / nr_subdiv;
          Amint_Amin += mn[d] * mn[d];
          Amint_d += mn[d] * f;
          Amaxt_Amax += mx[d] * mx[d];
          Amaxt_d += mx[d] * f;
        }
      }
      float min_radius = Amint_Amin == 0.0f ? (float)plane_radius : (float)std::min(Amint_d / Amint_Amin, plane_radius);
      float max_radius = Amaxt_Amax == 0.0f ? (float)plane_radius : (float)std::min(Amaxt_d / Amaxt_Amax, plane_radius);
      min_radius *= 1.1f;
      max_radius *= 0.9f;
      if (min_radius < max_radius) { r_min = min_radius; r_max = max_radius; }
      else { r_max = min_radius; r_min = max_radius; }
    }
    if (!std::isfinite(r_min) || !std::isfinite(r_max)) continue;
    desc.push_back(r_min);
    desc.push_back(r_max);
    kept.push_back(keypoints[k]);
  }
  keypoints.swap(kept);
  return desc;
}

// ===========================================================================
// a8-SC3D  computeLocalDescriptors(SC3D) -> pcl::ShapeContext3DEstimation<PointXYZRGB, Normal, ShapeContext1980>
// [REF src/dispatch_descriptors.h:47-48] [PCL-recall pcl/features/impl/3dsc.hpp initCompute + computePoint, 3dsc.h ctor:
//  azimuth 12 x elevation 11 x radius 15 = 1980 bins, min_radius 0.1, point_density_radius 0.2; log-spaced radii;
//  volume LUT 1 / cbrt(V); x axis = three draws of boost::uniform_01<mt19937> (seed 12345u, one estimator per call)
//  made orthogonal to the normal of the NEAREST surface point; every neighbour adds 1 / (local density * cbrt(V))
//  to its (azimuth, elevation, radius) bin.  The rf[9] of ShapeContext1980 is zeroed and is not part of the matcher's
//  point representation.]
// Canonical choices: neighbours in ascending index (PCL: ascending distance) for the float bin sums; nearest = smallest
// (d^2, index); Eigen's 3-vector reductions taken left to right; the LUTs use the host libm (both sides build them on the
// host with the same calls).
// ===========================================================================
struct Sc3dTables {
  float radii[16], theta_div[12], phi_div[13], volume_lut[1980];
};
static void sc3d_tables(double search_radius, Sc3dTables& t)
{
  const size_t azimuth_bins = 12, elevation_bins = 11, radius_bins = 15;
  const double min_radius = 0.1;
  const float azimuth_interval = 360.0f / (float)azimuth_bins;
  const float elevation_interval = 180.0f / (float)elevation_bins;
  for (size_t j = 0; j < radius_bins + 1; j++)
    t.radii[j] = (float)(exp(log(min_radius) + (((float)j / (float)radius_bins) * log(search_radius / min_radius))));
  for (size_t k = 0; k < elevation_bins + 1; k++) t.theta_div[k] = (float)k * elevation_interval;
  for (size_t l = 0; l < azimuth_bins + 1; l++) t.phi_div[l] = (float)l * azimuth_interval;
  const float integr_phi = (t.phi_div[1] * 0.017453293f) - (t.phi_div[0] * 0.017453293f);  // pcl::deg2rad (float)
  const float e = 1.0f / 3.0f;
  for (size_t j = 0; j < radius_bins; j++) {
    const float integr_r =
        (t.radii[j + 1] * t.radii[j + 1] * t.radii[j + 1] / 3.0f) - (t.radii[j] * t.radii[j] * t.radii[j] / 3.0f);
    for (size_t k = 0; k < elevation_bins; k++) {
      const float integr_theta = cosf(t.theta_div[k] * 0.017453293f) - cosf(t.theta_div[k + 1] * 0.017453293f);
      const float V = integr_phi * integr_theta * integr_r;
      for (size_t l = 0; l < azimuth_bins; l++) t.volume_lut[(l * elevation_bins * radius_bins) + k * radius_bins + j] = 1.0f / powf(V, e);
    }
  }
}

static std::vector<float> sc3d_descriptors(const Cloud& surface, const Normals& normals, Cloud& keypoints, double radius)
{
  const size_t azimuth_bins = 12, elevation_bins = 11, radius_bins = 15;
  const double point_density_radius = 0.2, min_radius = 0.1;
  std::vector<float> out;
  Cloud kept;
  if (radius < min_radius) {  // initCompute fails: "search_radius_ must be GREATER than min_radius_" -> empty output
    keypoints.swap(kept);
    return out;
  }
  Sc3dTables tb;
  sc3d_tables(radius, tb);
  Grid tree;
  tree.build(surface, (float)radius);
  // local point density of every surface point (computed on demand in PCL, identical values)
  std::vector<int> density(surface.size(), -1);
  const float dr2 = (float)(point_density_radius * point_density_radius);
  std::mt19937 rng(12345u);                                                  // == boost::mt19937
  auto rnd = [&]() { return (double)rng() * (1.0 / 4294967296.0); };         // boost::uniform_01<mt19937>
  auto is_zero = [](float v) { return std::fabs(v - 0.0f) < FLT_MIN; };      // pcl::utils::equal (v, 0.0f)
  std::vector<int> idx;
  std::vector<float> sqd;
  std::vector<float> desc(1980);
  for (size_t kp = 0; kp < keypoints.size(); ++kp) {
    const P4& o = keypoints[kp];
    if (!std::isfinite(o.x) || !std::isfinite(o.y) || !std::isfinite(o.z)) continue;  // NaN descriptor -> dropped
    tree.radius_sorted(o.x, o.y, o.z, radius, idx, sqd);
    if (idx.empty()) continue;  // NaN descriptor -> dropped; no random numbers drawn
    size_t b = 0;
    for (size_t i = 1; i < idx.size(); ++i)
      if (sqd[i] < sqd[b]) b = i;
    const float normal[3] = {normals[idx[b]].nx, normals[idx[b]].ny, normals[idx[b]].nz};
    float x_axis[3];
    x_axis[0] = (float)rnd();
    x_axis[1] = (float)rnd();
    x_axis[2] = (float)rnd();
    if (!is_zero(normal[2])) x_axis[2] = -(normal[0] * x_axis[0] + normal[1] * x_axis[1]) / normal[2];
    else if (!is_zero(normal[1])) x_axis[1] = -(normal[0] * x_axis[0] + normal[2] * x_axis[2]) / normal[1];
    else if (!is_zero(normal[0])) x_axis[0] = -(normal[1] * x_axis[1] + normal[2] * x_axis[2]) / normal[0];
    {
      const float z = (x_axis[0] * x_axis[0] + x_axis[1] * x_axis[1]) + x_axis[2] * x_axis[2];  // Eigen normalize()
      if (z > 0.0f) {
        const float nrm = std::sqrt(z);
        for (int a = 0; a < 3; ++a) x_axis[a] /= nrm;
      }
    }
    for (float& v : desc) v = 0.0f;
    for (size_t ne = 0; ne < idx.size(); ++ne) {
      if (is_zero(sqd[ne])) continue;
      const P4& q = surface[idx[ne]];
      const float r = std::sqrt(sqd[ne]);
      // pcl::geometry::project (neighbour, origin, normal, proj); proj -= origin
      const float po[3] = {q.x - o.x, q.y - o.y, q.z - o.z};
      const float lambda = (normal[0] * po[0] + normal[1] * po[1]) + normal[2] * po[2];
      float proj[3] = {q.x - lambda * normal[0], q.y - lambda * normal[1], q.z - lambda * normal[2]};
      proj[0] -= o.x; proj[1] -= o.y; proj[2] -= o.z;
      {
        const float z = (proj[0] * proj[0] + proj[1] * proj[1]) + proj[2] * proj[2];
        if (z > 0.0f) {
          const float nrm = std::sqrt(z);
          for (int a = 0; a < 3; ++a) proj[a] /= nrm;
        }
      }
      float cr[3];
      cross3(x_axis, proj, cr);
      const float cr_norm = std::sqrt((cr[0] * cr[0] + cr[1] * cr[1]) + cr[2] * cr[2]);
      float phi = m_atan2f(cr_norm, (x_axis[0] * proj[0] + x_axis[1] * proj[1]) + x_axis[2] * proj[2]) * 57.29578f;  // pcl::rad2deg
      phi = ((cr[0] * normal[0] + cr[1] * normal[1]) + cr[2] * normal[2]) < 0.f ? (360.0f - phi) : phi;
      float no[3] = {po[0], po[1], po[2]};
      {
        const float z = (no[0] * no[0] + no[1] * no[1]) + no[2] * no[2];
        if (z > 0.0f) {
          const float nrm = std::sqrt(z);
          for (int a = 0; a < 3; ++a) no[a] /= nrm;
        }
      }
      float theta = (normal[0] * no[0] + normal[1] * no[1]) + normal[2] * no[2];
      theta = m_acosf(std::min(1.0f, std::max(-1.0f, theta))) * 57.29578f;
      size_t j = 0, k = 0, l = 0;
      for (size_t rad = 1; rad < radius_bins + 1; rad++)
        if (r <= tb.radii[rad]) { j = rad - 1; break; }
      for (size_t ang = 1; ang < elevation_bins + 1; ang++)
        if (theta <= tb.theta_div[ang]) { k = ang - 1; break; }
      for (size_t ang = 1; ang < azimuth_bins + 1; ang++)
        if (phi <= tb.phi_div[ang]) { l = ang - 1; break; }
      int& dens = density[idx[ne]];
      if (dens < 0) {
        dens = 0;
        tree.for_radius(q.x, q.y, q.z, (float)point_density_radius, dr2, [&](int, float) { ++dens; });
      }
      if (dens == 0) continue;
      const float w = (1.0f / (float)dens) * tb.volume_lut[(l * elevation_bins * radius_bins) + (k * radius_bins) + j];
      desc[(l * elevation_bins * radius_bins) + (k * radius_bins) + j] += w;
    }
    bool finite = true;
    for (float v : desc)
      if (!std::isfinite(v)) finite = false;
    if (!finite) continue;
    out.insert(out.end(), desc.begin(), desc.end());
    kept.push_back(keypoints[kp]);
  }
  keypoints.swap(kept);
  return out;
}

// ===========================================================================
// a8-SHOT  computeLocalDescriptors(SHOT) -> pcl::SHOTColorEstimation<PointXYZRGB, Normal, SHOT1344>
// [REF src/dispatch_descriptors.h:46, src/features.cpp:99-150]
// [PCL-recall pcl/features/impl/shot.hpp (computeFeature, computePointSHOT, createBinDistanceShape,
//  interpolateDoubleChannel, normalizeHistogram, RGB2CIELAB), pcl/features/impl/shot_lrf.hpp (getLocalRF)]
// 32 spatial sectors x (10+1 shape bins, 30+1 colour bins) = 1344.  The keypoint cloud is what the reference
// passes as input_: after copyPointCloud its colour is the default (0,0,0), so the colour reference is black.
// ===========================================================================
static const double PST_PI = 3.1415926535897932384626433832795;
static const double PST_RAD_45 = 0.78539816339744830961566084581988;
static const double PST_RAD_90 = 1.5707963267948966192313216916398;
static const double PST_RAD_135 = 2.3561944901923449288469825374596;
static const double PST_RAD_PI_7_8 = 2.7488935718910690836548129603691;

struct LabLut {
  float srgb[256];
  float xyz[4000];
  LabLut()
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
static const LabLut& lab_lut()
{
  static LabLut l;
  return l;
}

static void rgb2cielab(unsigned char R, unsigned char G, unsigned char B, float& L, float& A, float& B2)
{
  const LabLut& lut = lab_lut();
  const float fr = lut.srgb[R], fg = lut.srgb[G], fb = lut.srgb[B];
  const float x = fr * 0.412453f + fg * 0.357580f + fb * 0.180423f;
  const float y = fr * 0.212671f + fg * 0.715160f + fb * 0.072169f;
  const float z = fr * 0.019334f + fg * 0.119193f + fb * 0.950227f;
  float vx = x / 0.95047f;
  float vy = y;
  float vz = z / 1.08883f;
  vx = lut.xyz[int(vx * 4000)];
  vy = lut.xyz[int(vy * 4000)];
  vz = lut.xyz[int(vz * 4000)];
  L = 116.0f * vy - 16.0f;
  if (L > 100) L = 100.0f;
  A = 500.0f * (vx - vy);
  if (A > 120) A = 120.0f;
  else if (A < -120) A = -120.0f;
  B2 = 200.0f * (vy - vz);
  if (B2 > 120) B2 = 120.0f;
  else if (B2 < -120) B2 = -120.0f;
}

// SHOTLocalReferenceFrameEstimation::getLocalRF; rf row-major (x axis, y axis, z axis); false = NaN frame
static bool shot_lrf(const Cloud& surface, const P4& center, double radius, const std::vector<int>& idx, const std::vector<float>& sqd, float* rf)
{
  std::vector<double> vij;
  vij.reserve(idx.size() * 3);
  double cov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  double sum = 0.0;
  int valid = 0;
  for (size_t i = 0; i < idx.size(); ++i) {
    const P4& pt = surface[idx[i]];
    if (pt.x == center.x && pt.y == center.y && pt.z == center.z) continue;
    const double v[3] = {(double)(pt.x - center.x), (double)(pt.y - center.y), (double)(pt.z - center.z)};
    const double distance = radius - std::sqrt((double)sqd[i]);
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) cov[r * 3 + c] += distance * (v[r] * v[c]);
    sum += distance;
    vij.push_back(v[0]); vij.push_back(v[1]); vij.push_back(v[2]);
    ++valid;
  }
  if (valid < 5) return false;
  for (int k = 0; k < 9; ++k) cov[k] /= sum;
  double val[3], vec[9];
  em::eig3_sym_d(cov, val, vec);
  if (!std::isfinite(val[0]) || !std::isfinite(val[1]) || !std::isfinite(val[2])) return false;
  double v1[3] = {vec[0 * 3 + 2], vec[1 * 3 + 2], vec[2 * 3 + 2]};  // largest eigenvalue
  double v3[3] = {vec[0 * 3 + 0], vec[1 * 3 + 0], vec[2 * 3 + 0]};  // smallest
  int plusNormal = 0, plusTangent = 0;
  for (int ne = 0; ne < valid; ++ne) {
    const double* r = &vij[ne * 3];
    if ((r[0] * v1[0] + r[1] * v1[1]) + r[2] * v1[2] >= 0) ++plusTangent;
    if ((r[0] * v3[0] + r[1] * v3[1]) + r[2] * v3[2] >= 0) ++plusNormal;
  }
  plusTangent = 2 * plusTangent - valid;
  if (plusTangent == 0) {
    const int points = 5, median = valid / 2;
    for (int i = -points / 2; i <= points / 2; ++i) {
      const double* r = &vij[(median - i) * 3];
      if ((r[0] * v1[0] + r[1] * v1[1]) + r[2] * v1[2] > 0) ++plusTangent;
    }
    if (plusTangent < points / 2 + 1) for (double& c : v1) c *= -1;
  } else if (plusTangent < 0) {
    for (double& c : v1) c *= -1;
  }
  plusNormal = 2 * plusNormal - valid;
  if (plusNormal == 0) {
    const int points = 5, median = valid / 2;
    for (int i = -points / 2; i <= points / 2; ++i) {
      const double* r = &vij[(median - i) * 3];
      if ((r[0] * v3[0] + r[1] * v3[1]) + r[2] * v3[2] > 0) ++plusNormal;
    }
    if (plusNormal < points / 2 + 1) for (double& c : v3) c *= -1;
  } else if (plusNormal < 0) {
    for (double& c : v3) c *= -1;
  }
  for (int k = 0; k < 3; ++k) { rf[k] = (float)v1[k]; rf[6 + k] = (float)v3[k]; }
  // rf.row(1) = rf.row(2).cross(rf.row(0))
  rf[3] = rf[7] * rf[2] - rf[8] * rf[1];
  rf[4] = rf[8] * rf[0] - rf[6] * rf[2];
  rf[5] = rf[6] * rf[1] - rf[7] * rf[0];
  return true;
}

static std::vector<float> shot_descriptors(const Cloud& surface, const Normals& normals, Cloud& keypoints, double radius,
                                           std::vector<float>* rf_dbg = nullptr)
{
  const int nr_shape = 10, nr_color = 30, sectors = 32, D = 1344;
  const int shapeToColorStride = sectors * (nr_shape + 1);
  Grid tree;
  tree.build(surface, (float)radius);
  const double radius3_4 = (radius * 3) / 4, radius1_4 = radius / 4, radius1_2 = radius / 2;
  std::vector<float> desc;
  Cloud kept;
  std::vector<int> idx;
  std::vector<float> sqd;
  std::vector<float> shot(D);
  for (size_t k = 0; k < keypoints.size(); ++k) {
    const P4& c = keypoints[k];
    tree.radius_sorted(c.x, c.y, c.z, radius, idx, sqd);
    float rf[9];
    if (!shot_lrf(surface, c, radius, idx, sqd, rf)) continue;  // NaN frame -> NaN descriptor -> dropped
    if (idx.empty() || idx.size() < 5) continue;                // "not sufficient for its description" -> NaN -> dropped
    for (float& v : shot) v = 0.f;
    // reference colour = the keypoint's own colour
    float LRef, aRef, bRef;
    rgb2cielab((c.rgba >> 16) & 0xff, (c.rgba >> 8) & 0xff, c.rgba & 0xff, LRef, aRef, bRef);
    LRef /= 100.0f; aRef /= 120.0f; bRef /= 120.0f;
    for (size_t i = 0; i < idx.size(); ++i) {
      const N4& nq = normals[idx[i]];
      // createBinDistanceShape
      if (!std::isfinite(nq.nx) || !std::isfinite(nq.ny) || !std::isfinite(nq.nz)) continue;
      double cosineDesc = (double)(((nq.nx * rf[6] + nq.ny * rf[7]) + nq.nz * rf[8]) + 0.0f);
      if (cosineDesc > 1.0) cosineDesc = 1.0;
      if (cosineDesc < -1.0) cosineDesc = -1.0;
      double binDistanceShape = ((1.0 + cosineDesc) * nr_shape) / 2;
      // colour
      const P4& sp = surface[idx[i]];
      float L, a, b;
      rgb2cielab((sp.rgba >> 16) & 0xff, (sp.rgba >> 8) & 0xff, sp.rgba & 0xff, L, a, b);
      L /= 100.0f; a /= 120.0f; b /= 120.0f;
      double colorDistance = (std::fabs(LRef - L) + ((std::fabs(aRef - a) + std::fabs(bRef - b)) / 2)) / 3;
      if (colorDistance > 1.0) colorDistance = 1.0;
      if (colorDistance < 0.0) colorDistance = 0.0;
      double binDistanceColor = colorDistance * nr_color;
      // interpolateDoubleChannel
      const float dl[3] = {sp.x - c.x, sp.y - c.y, sp.z - c.z};
      const double distance = std::sqrt((double)sqd[i]);
      if (std::fabs(distance - 0.0) < 1E-15) continue;
      double xInFeatRef = (double)((dl[0] * rf[0] + dl[1] * rf[1]) + dl[2] * rf[2]);
      double yInFeatRef = (double)((dl[0] * rf[3] + dl[1] * rf[4]) + dl[2] * rf[5]);
      double zInFeatRef = (double)((dl[0] * rf[6] + dl[1] * rf[7]) + dl[2] * rf[8]);
      if (std::fabs(yInFeatRef) < 1E-30) yInFeatRef = 0;
      if (std::fabs(xInFeatRef) < 1E-30) xInFeatRef = 0;
      if (std::fabs(zInFeatRef) < 1E-30) zInFeatRef = 0;
      const unsigned char bit4 = ((yInFeatRef > 0) || ((yInFeatRef == 0.0) && (xInFeatRef < 0))) ? 1 : 0;
      const unsigned char bit3 = (unsigned char)(((xInFeatRef > 0) || ((xInFeatRef == 0.0) && (yInFeatRef > 0))) ? !bit4 : bit4);
      int desc_index = (bit4 << 3) + (bit3 << 2);
      desc_index = desc_index << 1;
      if ((xInFeatRef * yInFeatRef > 0) || (xInFeatRef == 0.0)) desc_index += (std::fabs(xInFeatRef) >= std::fabs(yInFeatRef)) ? 0 : 4;
      else desc_index += (std::fabs(xInFeatRef) > std::fabs(yInFeatRef)) ? 4 : 0;
      desc_index += zInFeatRef > 0 ? 1 : 0;
      desc_index += (distance > radius1_2) ? 2 : 0;
      const int step_index_shape = (int)std::floor(binDistanceShape + 0.5);
      const int step_index_color = (int)std::floor(binDistanceColor + 0.5);
      const int volume_index_shape = desc_index * (nr_shape + 1);
      const int volume_index_color = shapeToColorStride + desc_index * (nr_color + 1);
      binDistanceShape -= step_index_shape;
      binDistanceColor -= step_index_color;
      double intWeightShape = (1 - std::fabs(binDistanceShape));
      double intWeightColor = (1 - std::fabs(binDistanceColor));
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
#ifdef ORACLE_LIBM
      const double inclination = std::acos(inclinationCosine);
#else
      const double inclination = em::acos_d_(inclinationCosine);
#endif
      if (inclination > PST_RAD_90 || (std::fabs(inclination - PST_RAD_90) < 1e-30 && zInFeatRef <= 0)) {
        const double inclinationDistance = (inclination - PST_RAD_135) / PST_RAD_90;
        if (inclination > PST_RAD_135) {
          intWeightShape += 1 - inclinationDistance;
          intWeightColor += 1 - inclinationDistance;
        } else {
          intWeightShape += 1 + inclinationDistance;
          intWeightColor += 1 + inclinationDistance;
          shot[(desc_index + 1) * (nr_shape + 1) + step_index_shape] -= (float)inclinationDistance;
          shot[shapeToColorStride + (desc_index + 1) * (nr_color + 1) + step_index_color] -= (float)inclinationDistance;
        }
      } else {
        const double inclinationDistance = (inclination - PST_RAD_45) / PST_RAD_90;
        if (inclination < PST_RAD_45) {
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
#ifdef ORACLE_LIBM
        const double azimuth = std::atan2(yInFeatRef, xInFeatRef);
#else
        const double azimuth = em::atan2_d_(yInFeatRef, xInFeatRef);
#endif
        const int sel = desc_index >> 2;
        const double angularSectorSpan = PST_RAD_45;
        const double angularSectorStart = -PST_RAD_PI_7_8;
        double azimuthDistance = (azimuth - (angularSectorStart + angularSectorSpan * sel)) / angularSectorSpan;
        azimuthDistance = std::max(-0.5, std::min(azimuthDistance, 0.5));
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
    }
    // normalizeHistogram
    double acc_norm = 0.0;
    for (int j = 0; j < D; ++j) acc_norm += shot[j] * shot[j];
    acc_norm = std::sqrt(acc_norm);
    bool finite = true;
    for (int j = 0; j < D; ++j) {
      shot[j] /= (float)acc_norm;
      if (!std::isfinite(shot[j])) finite = false;
    }
    if (!finite) continue;
    desc.insert(desc.end(), shot.begin(), shot.end());
    kept.push_back(c);
    if (rf_dbg) rf_dbg->insert(rf_dbg->end(), rf, rf + 9);
  }
  keypoints.swap(kept);
  return desc;
}

// ===========================================================================
// a9  findFeatureCorrespondences  [REF src/matching.cpp:31-93]
// distance = flann::L2_Simple over D floats [PCL-recall flann/algorithms/dist.h]
// ===========================================================================
struct Corr {
  int q, m;
  float d;
};

static void knn_bruteforce(const float* A, size_t na, const float* B, size_t nb, int D, int k, std::vector<int>& idx,
                           std::vector<float>& dist)
{
  // for each row of A: k nearest rows of B, sorted by (distance, index)
  idx.assign(na * k, -1);
  dist.assign(na * k, 0.f);
#pragma omp parallel
  {
  std::vector<std::pair<float, int>> best;
#pragma omp for schedule(dynamic, 32)
  for (long long i = 0; i < (long long)na; ++i) {
    best.clear();
    const float* a = A + i * D;
    for (size_t j = 0; j < nb; ++j) {
      const float* b = B + j * D;
      float r = 0.f;
      for (int t = 0; t < D; ++t) {
        const float diff = a[t] - b[t];
        r += diff * diff;
      }
      const std::pair<float, int> cand(r, (int)j);
      if ((int)best.size() < k) best.insert(std::upper_bound(best.begin(), best.end(), cand), cand);
      else if (cand < best.back()) {
        best.pop_back();
        best.insert(std::upper_bound(best.begin(), best.end(), cand), cand);
      }
    }
    for (size_t t = 0; t < best.size(); ++t) {
      idx[i * k + t] = best[t].second;
      dist[i * k + t] = best[t].first;
    }
  }
  }
}

static std::vector<Corr> find_correspondences(const float* S, size_t ns, const float* T, size_t nt, int D, size_t k_in)
{
  std::vector<Corr> out;
  if (ns == 0 || nt == 0 || k_in == 0) return out;
  const int kf = (int)std::min(k_in, nt);  // KdTreeFLANN clamps k to the cloud size
  const int kb = (int)std::min(k_in, ns);
  std::vector<int> fi, bi;
  std::vector<float> fd, bd;
  knn_bruteforce(S, ns, T, nt, D, kf, fi, fd);
  knn_bruteforce(T, nt, S, ns, D, kb, bi, bd);
  for (size_t i = 0; i < ns; ++i) {
    for (int j = 0; j < kf; ++j) {
      const int match = fi[i * kf + j];
      bool hit = false;
      for (int b = 0; b < kb; ++b)
        if (bi[(size_t)match * kb + b] == (int)i) { hit = true; break; }
      if (hit) {
        out.push_back(Corr{(int)i, match, fd[i * kf + j]});
        break;
      }
    }
  }
  return out;
}

// ===========================================================================
// pcl::umeyama (= Eigen::umeyama, no scaling), literal, sequential sums
// [PCL-recall Eigen/src/Geometry/Umeyama.h]
// ===========================================================================
template <typename T>
static void umeyama(const std::vector<T>& src, const std::vector<T>& dst, size_t n, T* Rt /*row-major 4x4*/)
{
  const T one_over_n = T(1) / (T)n;
  T sm[3] = {0, 0, 0}, dm[3] = {0, 0, 0};
  for (size_t i = 0; i < n; ++i)
    for (int a = 0; a < 3; ++a) { sm[a] += src[i * 3 + a]; dm[a] += dst[i * 3 + a]; }
  for (int a = 0; a < 3; ++a) { sm[a] *= one_over_n; dm[a] *= one_over_n; }
  T sigma[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (size_t i = 0; i < n; ++i) {
    T sd[3], dd[3];
    for (int a = 0; a < 3; ++a) { sd[a] = src[i * 3 + a] - sm[a]; dd[a] = dst[i * 3 + a] - dm[a]; }
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) sigma[r * 3 + c] += (one_over_n * dd[r]) * sd[c];
  }
  T U[9], V[9], d[3];
  em::svd3<T>(sigma, U, d, V);
  T S[3] = {1, 1, 1};
  if (em::det3<T>(sigma) < 0) S[2] = -1;
  const T prec = sizeof(T) == 4 ? T(1e-5) : T(1e-12);  // NumTraits::dummy_precision
  int rank = 0;
  for (int i = 0; i < 3; ++i)
    if (!(std::fabs(d[i]) <= std::fabs(d[0]) * prec)) ++rank;
  if (rank == 2) {
    if (em::det3<T>(U) * em::det3<T>(V) > 0) S[2] = 1;
    else S[2] = -1;
  }
  T R[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      T acc = (U[i * 3 + 0] * S[0]) * V[j * 3 + 0];
      acc += (U[i * 3 + 1] * S[1]) * V[j * 3 + 1];
      acc += (U[i * 3 + 2] * S[2]) * V[j * 3 + 2];
      R[i * 3 + j] = acc;
    }
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) Rt[i * 4 + j] = R[i * 3 + j];
    T rs = R[i * 3 + 0] * sm[0];
    rs += R[i * 3 + 1] * sm[1];
    rs += R[i * 3 + 2] * sm[2];
    Rt[i * 4 + 3] = dm[i] - rs;
  }
  Rt[12] = Rt[13] = Rt[14] = 0;
  Rt[15] = 1;
}

// ===========================================================================
// a10  estimateTransformFromCorrespondences  [REF src/matching.cpp:110-140]
// [PCL-recall pcl/registration/impl/correspondence_rejection_sample_consensus.hpp,
//  pcl/sample_consensus/impl/ransac.hpp, sac_model.h, impl/sac_model_registration.hpp,
//  boost::mt19937(12345) + uniform_int<>(0, INT_MAX) == mt() >> 1]
// ===========================================================================
struct RansacDebug {
  int iterations = 0;
  int best_count = 0;
  double sample_dist_thresh = 0;
  Mat4 best_model;
};

static bool is_identity(const Mat4& t)
{
  // Eigen isIdentity(prec = 1e-5)
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      const float v = t.m[i * 4 + j];
      if (i == j) {
        if (!(std::fabs(v - 1.0f) <= 1e-5f * std::min(std::fabs(v), 1.0f))) return false;
      } else {
        if (!(std::fabs(v) <= 1e-5f)) return false;
      }
    }
  return true;
}

// pcl::registration::CorrespondenceRejectorSampleConsensus::getRemainingCorrespondences: true = a model was found and it has
// at least three inliers (inliers = positions in corr, best = getBestTransformation()); false = PCL keeps the original
// correspondences and an identity best transformation.
static bool ransac_reject(const Cloud& skp, const Cloud& tkp, const std::vector<Corr>& corr, double inlier_threshold,
                          std::vector<int>& inliers /* positions in corr */, Mat4& best, RansacDebug* dbg = nullptr)
{
  inliers.clear();
  best = Mat4::identity();
  const int nc = (int)corr.size();
  const int max_iterations = 1000;  // CorrespondenceRejectorSampleConsensus default; the reference never overrides it
  if (nc == 0) return false;  // fence: the reference reads an uninitialised matrix here
  std::vector<int> indices(nc), indices_tgt(nc);
  for (int i = 0; i < nc; ++i) { indices[i] = corr[i].q; indices_tgt[i] = corr[i].m; }
  std::map<int, int> correspondences;  // computeOriginalIndexMapping
  for (int i = 0; i < nc; ++i) correspondences[indices[i]] = indices_tgt[i];
  // computeSampleDistanceThreshold
  float cov[3][3], cen[3];
  mean_and_cov(skp, indices, cov, cen);
  float evals[3];
  eigen33_values(cov, evals);
  double sample_dist_thresh = (double)((std::sqrt(evals[0]) + std::sqrt(evals[1])) + std::sqrt(evals[2])) / 3.0;
  sample_dist_thresh *= sample_dist_thresh;
  if (dbg) dbg->sample_dist_thresh = sample_dist_thresh;

  std::mt19937 rng(12345u);
  std::vector<int> shuffled(indices);
  const double thresh = inlier_threshold * inlier_threshold;

  int iterations = 0;
  int n_best = -INT_MAX;
  double k = 1.0;
  const double log_probability = std::log(1.0 - 0.99);
  const double one_over_indices = 1.0 / (double)nc;
  std::vector<int> model;
  Mat4 model_coefficients = Mat4::identity(), best_coefficients = Mat4::identity();
  std::vector<int> selection;
  while (iterations < k) {
    // getSamples
    selection.clear();
    if (nc < 3) {
      iterations = INT_MAX - 1;
    } else {
      selection.resize(3);
      bool good = false;
      for (unsigned iter = 0; iter < 1000; ++iter) {
        for (unsigned i = 0; i < 3; ++i) {
          const int rnd = (int)(rng() >> 1);
          std::swap(shuffled[i], shuffled[i + ((size_t)rnd % (size_t)(nc - i))]);
        }
        for (int i = 0; i < 3; ++i) selection[i] = shuffled[i];
        const P4 &a = skp[selection[0]], &b = skp[selection[1]], &c = skp[selection[2]];
        const float p10[3] = {b.x - a.x, b.y - a.y, b.z - a.z};
        const float p20[3] = {c.x - a.x, c.y - a.y, c.z - a.z};
        const float p21[3] = {c.x - b.x, c.y - b.y, c.z - b.z};
        if ((p10[0] * p10[0] + p10[1] * p10[1] + p10[2] * p10[2]) > sample_dist_thresh &&
            (p20[0] * p20[0] + p20[1] * p20[1] + p20[2] * p20[2]) > sample_dist_thresh &&
            (p21[0] * p21[0] + p21[1] * p21[1] + p21[2] * p21[2]) > sample_dist_thresh) {
          good = true;
          break;
        }
      }
      if (!good) selection.clear();
    }
    if (selection.empty()) break;
    // computeModelCoefficients: Umeyama in double on the 3 sampled pairs
    std::vector<double> s3(9), t3(9);
    for (int i = 0; i < 3; ++i) {
      const P4& sp = skp[selection[i]];
      const P4& tp = tkp[correspondences[selection[i]]];
      s3[i * 3 + 0] = sp.x; s3[i * 3 + 1] = sp.y; s3[i * 3 + 2] = sp.z;
      t3[i * 3 + 0] = tp.x; t3[i * 3 + 1] = tp.y; t3[i * 3 + 2] = tp.z;
    }
    double Rt[16];
    umeyama<double>(s3, t3, 3, Rt);
    for (int i = 0; i < 16; ++i) model_coefficients.m[i] = (float)Rt[i];
    // countWithinDistance
    int n_inliers = 0;
    for (int i = 0; i < nc; ++i) {
      const P4& sp = skp[indices[i]];
      const P4& tp = tkp[indices_tgt[i]];
      float px, py, pz;
      xform(model_coefficients, sp.x, sp.y, sp.z, px, py, pz);
      const float ex = px - tp.x, ey = py - tp.y, ez = pz - tp.z;
      if (((ex * ex + ey * ey) + ez * ez) < thresh) ++n_inliers;
    }
    if (n_inliers > n_best) {
      n_best = n_inliers;
      model = selection;
      best_coefficients = model_coefficients;
      const double w = (double)n_best * one_over_indices;
      double p_no_outliers = 1.0 - std::pow(w, 3.0);
      p_no_outliers = std::max(std::numeric_limits<double>::epsilon(), p_no_outliers);
      p_no_outliers = std::min(1.0 - std::numeric_limits<double>::epsilon(), p_no_outliers);
      k = log_probability / std::log(p_no_outliers);
    }
    ++iterations;
    if (iterations > max_iterations) break;
  }
  if (dbg) { dbg->iterations = iterations; dbg->best_count = n_best; dbg->best_model = best_coefficients; }
  if (model.empty()) return false;  // computeModel false -> identity -> reference returns zero
  // selectWithinDistance
  for (int i = 0; i < nc; ++i) {
    const P4& sp = skp[indices[i]];
    const P4& tp = tkp[indices_tgt[i]];
    float px, py, pz;
    xform(best_coefficients, sp.x, sp.y, sp.z, px, py, pz);
    const float ex = px - tp.x, ey = py - tp.y, ez = pz - tp.z;
    if (((ex * ex + ey * ey) + ez * ez) < thresh) inliers.push_back(i);
  }
  if (inliers.size() < 3) {
    inliers.clear();
    return false;
  }
  best = best_coefficients;
  return true;
}

// TransformationEstimationSVD<PointT, PointT, float>::estimateRigidTransformation over correspondences
static Mat4 svd_transform(const Cloud& skp, const Cloud& tkp, const std::vector<Corr>& corr, const std::vector<int>& inliers)
{
  std::vector<float> sv(inliers.size() * 3), tv(inliers.size() * 3);
  for (size_t i = 0; i < inliers.size(); ++i) {
    const P4& sp = skp[corr[inliers[i]].q];
    const P4& tp = tkp[corr[inliers[i]].m];
    sv[i * 3 + 0] = sp.x; sv[i * 3 + 1] = sp.y; sv[i * 3 + 2] = sp.z;
    tv[i * 3 + 0] = tp.x; tv[i * 3 + 1] = tp.y; tv[i * 3 + 2] = tp.z;
  }
  Mat4 result;
  umeyama<float>(sv, tv, inliers.size(), result.m);
  return result;
}

// estimateTransformFromCorrespondences (matching.cpp:110-140): RANSAC rejection, "identity means failure", SVD on the inliers
static Mat4 ransac_transform(const Cloud& skp, const Cloud& tkp, const std::vector<Corr>& corr, double inlier_threshold,
                             std::vector<int>& inliers /* positions in corr */, RansacDebug* dbg = nullptr)
{
  Mat4 best;
  if (!ransac_reject(skp, tkp, corr, inlier_threshold, inliers, best, dbg) || is_identity(best)) {  // matching.cpp:128-133
    inliers.clear();
    return Mat4::zero();
  }
  return svd_transform(skp, tkp, corr, inliers);
}

// ===========================================================================
// a12-SAC_IA  estimateTransformFromDescriptorsSets -> pcl::SampleConsensusInitialAlignment
// [REF src/matching.cpp:142-194: setMinSampleDistance(inlier_threshold), setMaxCorrespondenceDistance(max_corr),
//  setMaximumIterations(max_iterations); inputs are the KEYPOINT clouds and their descriptors]
// [PCL-recall pcl/registration/impl/ia_ransac.hpp: nr_samples_ = 3, k_correspondences_ = 10, TruncatedError on the
//  SQUARED nearest-neighbour distance, C rand() (never seeded -> glibc TYPE_3 generator from seed 1)]
// The rand() stream is process-global in the reference; here it is an explicit state that the caller threads
// through the pair loop in row-major pair order (a fresh process's first estimateMapsTransforms call).
// ===========================================================================
struct GlibcRand {
  int32_t r[34];
  int f, b;  // front / rear indices of the additive feedback generator
  GlibcRand() { seed(1); }
  void seed(uint32_t s)
  {
    // glibc srandom_r, TYPE_3 (degree 31, separation 3)
    if (s == 0) s = 1;
    r[0] = (int32_t)s;
    for (int i = 1; i < 31; ++i) {
      const long hi = r[i - 1] / 127773, lo = r[i - 1] % 127773;
      long word = 16807 * lo - 2836 * hi;
      if (word < 0) word += 2147483647;
      r[i] = (int32_t)word;
    }
    f = 3;
    b = 0;
    for (int i = 0; i < 310; ++i) next();
  }
  int next()
  {
    uint32_t* st = (uint32_t*)r;
    st[f] += st[b];
    const int result = (int)(st[f] >> 1);
    if (++f >= 31) f = 0;
    if (++b >= 31) b = 0;
    return result;
  }
  uint64_t calls = 0;
  int rand_() { ++calls; return next(); }
};

static inline int sac_random_index(GlibcRand& rng, int n)
{
  return (int)(n * (rng.rand_() / (2147483647 + 1.0)));  // getRandomIndex: n * (rand() / (RAND_MAX + 1.0))
}

static Mat4 sac_ia_transform(const Cloud& skp, const std::vector<float>& sdesc, const Cloud& tkp, const std::vector<float>& tdesc, int D,
                             double min_sample_distance_d, double max_corr_dist, int max_iterations, GlibcRand& rng,
                             std::vector<float>* errors_dbg = nullptr)
{
  const int nr_samples = 3, k_corr = 10;
  const int ns = (int)skp.size(), nt = (int)tkp.size();
  Mat4 final_t = Mat4::identity();  // final_transformation_ = guess (identity): what align() leaves when nothing better is found
  if (ns < nr_samples || nt == 0) return final_t;  // "No. of samples > cloud size": selectSamples bails out; fence for empty target
  float min_sample_distance = (float)min_sample_distance_d;
  const float thr = (float)max_corr_dist;  // TruncatedError(corr_dist_threshold_)
  // feature-space k-NN of every source feature in the target set (exact, sorted)
  const int kk = std::min(k_corr, nt);
  std::vector<int> fi;
  std::vector<float> fd;
  knn_bruteforce(sdesc.data(), ns, tdesc.data(), nt, D, kk, fi, fd);
  Grid tree;
  tree.build(tkp, std::max(thr, 0.05f));
  float lowest_error = 0.f;
  std::vector<std::pair<float, int>> nn;
  for (int it = 0; it < max_iterations; ++it) {
    // selectSamples
    std::vector<int> sample;
    int without = 0;
    const int max_without = 3 * ns;
    while ((int)sample.size() < nr_samples) {
      const int idx = sac_random_index(rng, ns);
      bool valid = true;
      for (size_t i = 0; i < sample.size(); ++i) {
        const P4 &a = skp[idx], &b = skp[sample[i]];
        const float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
        const float dist = std::sqrt((dx * dx + dy * dy) + dz * dz);  // euclideanDistance
        if (idx == sample[i] || dist < min_sample_distance) { valid = false; break; }
      }
      if (valid) { sample.push_back(idx); without = 0; }
      else ++without;
      if (without >= max_without) { min_sample_distance *= 0.5f; without = 0; }
    }
    // findSimilarFeatures: one of the k nearest target features at random
    int corr[3];
    for (int i = 0; i < nr_samples; ++i) {
      const int rc = sac_random_index(rng, k_corr);
      corr[i] = fi[(size_t)sample[i] * kk + std::min(rc, kk - 1)];
    }
    // TransformationEstimationSVD (float Umeyama) on the three pairs
    std::vector<float> sv(9), tv(9);
    for (int i = 0; i < 3; ++i) {
      sv[i * 3] = skp[sample[i]].x; sv[i * 3 + 1] = skp[sample[i]].y; sv[i * 3 + 2] = skp[sample[i]].z;
      tv[i * 3] = tkp[corr[i]].x; tv[i * 3 + 1] = tkp[corr[i]].y; tv[i * 3 + 2] = tkp[corr[i]].z;
    }
    Mat4 T;
    umeyama<float>(sv, tv, 3, T.m);
    // computeErrorMetric over the transformed source keypoints
    float error = 0.f;
    for (int i = 0; i < ns; ++i) {
      float x, y, z;
      xform(T, skp[i].x, skp[i].y, skp[i].z, x, y, z);
      tree.knn(x, y, z, 1, nn);
      const float e = nn[0].first;
      error += (e <= thr) ? (e / thr) : 1.0f;
    }
    if (errors_dbg) errors_dbg->push_back(error);
    if (it == 0 || error < lowest_error) {
      lowest_error = error;
      final_t = T;
    }
  }
  return final_t;
}

// ===========================================================================
// a11  estimateTransformICP -> pcl::IterativeClosestPoint  [REF src/matching.cpp:196-221]
// [PCL-recall pcl/registration/impl/icp.hpp, correspondence_estimation.hpp,
//  default_convergence_criteria.hpp, transformation_estimation_svd.hpp]
// ===========================================================================
struct IcpDebug {
  int iterations = 0;
  int converged = 0;
  std::vector<long long> sums;  // per iteration: n, Sp[3], Sq[3], Sqp[9], Sd
};

static Mat4 icp_refine(const Cloud& source, const Cloud& target, const Mat4& initial_guess, double max_corr_dist,
                       int max_iterations, double transformation_epsilon, IcpDebug* dbg = nullptr)
{
  // pcl::transformPointCloud(source, transformed, initial_guess)
  std::vector<float> pts(source.size() * 3);
  for (size_t i = 0; i < source.size(); ++i) xform(initial_guess, source[i].x, source[i].y, source[i].z, pts[i * 3], pts[i * 3 + 1], pts[i * 3 + 2]);
  Grid tree;
  tree.build(target, (float)std::max(max_corr_dist * 0.25, 1e-3));
  const double max_dist_sqr = max_corr_dist * max_corr_dist;
  Mat4 final_t = Mat4::identity();
  int nr_iterations = 0;
  bool converged = false;
  double prev_mse = DBL_MAX;
  const double rotation_threshold = 1.0 - transformation_epsilon;
  const double translation_threshold = transformation_epsilon;
  do {
    long long cnt = 0, Sp[3] = {0, 0, 0}, Sq[3] = {0, 0, 0}, Sqp[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, Sd = 0;
    // integer (fixed-point) sums: order-free, so the loop may run on all host threads
#pragma omp parallel for schedule(dynamic, 1024) reduction(+ : cnt, Sp[:3], Sq[:3], Sqp[:9], Sd)
    for (long long i = 0; i < (long long)source.size(); ++i) {
      int j;
      float dd;
      const float px = pts[i * 3], py = pts[i * 3 + 1], pz = pts[i * 3 + 2];
      if (!tree.nn1(px, py, pz, max_dist_sqr, j, dd)) continue;  // distance[0] > max_dist_sqr -> skipped
      const P4& q = target[j];
      const double p[3] = {px, py, pz}, qq[3] = {q.x, q.y, q.z};
      ++cnt;
      for (int a = 0; a < 3; ++a) {
        Sp[a] += em::to_fix(p[a], MM3D_FIX1_SCALE);
        Sq[a] += em::to_fix(qq[a], MM3D_FIX1_SCALE);
        for (int b = 0; b < 3; ++b) Sqp[a * 3 + b] += em::to_fix(qq[a] * p[b], MM3D_FIX2_SCALE);
      }
      Sd += em::to_fix((double)dd, MM3D_FIXD_SCALE);
    }
    if (dbg) {
      dbg->sums.push_back(cnt);
      for (int a = 0; a < 3; ++a) dbg->sums.push_back(Sp[a]);
      for (int a = 0; a < 3; ++a) dbg->sums.push_back(Sq[a]);
      for (int a = 0; a < 9; ++a) dbg->sums.push_back(Sqp[a]);
      dbg->sums.push_back(Sd);
    }
    if (cnt < 3) {  // min_number_correspondences_
      converged = false;
      break;
    }
    // TransformationEstimationSVD<float> (Umeyama) from the exact sums
    const double n = (double)cnt;
    double pm[3], qm[3];
    for (int a = 0; a < 3; ++a) {
      pm[a] = ((double)Sp[a] / MM3D_FIX1_SCALE) / n;
      qm[a] = ((double)Sq[a] / MM3D_FIX1_SCALE) / n;
    }
    float sigma[9], pmf[3], qmf[3];
    for (int a = 0; a < 3; ++a) {
      pmf[a] = (float)pm[a];
      qmf[a] = (float)qm[a];
      for (int b = 0; b < 3; ++b) sigma[a * 3 + b] = (float)(((double)Sqp[a * 3 + b] / MM3D_FIX2_SCALE) / n - qm[a] * pm[b]);
    }
    Mat4 T;
    em::umeyama_from_sigma<float>(sigma, pmf, qmf, T.m);
    // transformCloud(input_transformed, input_transformed, transformation_)
    for (size_t i = 0; i < source.size(); ++i) {
      float ox, oy, oz;
      xform(T, pts[i * 3], pts[i * 3 + 1], pts[i * 3 + 2], ox, oy, oz);
      pts[i * 3] = ox; pts[i * 3 + 1] = oy; pts[i * 3 + 2] = oz;
    }
    final_t = mul(T, final_t);
    ++nr_iterations;
    // DefaultConvergenceCriteria::hasConverged
    converged = false;
    if (nr_iterations >= max_iterations) {
      converged = true;
    } else {
      const double cos_angle = 0.5 * (T.m[0] + T.m[5] + T.m[10] - 1);
      const double translation_sqr = T.m[3] * T.m[3] + T.m[7] * T.m[7] + T.m[11] * T.m[11];
      if (cos_angle >= rotation_threshold && translation_sqr <= translation_threshold) {
        converged = true;
      } else {
        const double mse = ((double)Sd / MM3D_FIXD_SCALE) / n;
        if (std::fabs(mse - prev_mse) < 1e-12) converged = true;
        else prev_mse = mse;
      }
    }
  } while (!converged);
  if (dbg) { dbg->iterations = nr_iterations; dbg->converged = converged ? 1 : 0; }
  return mul(final_t, initial_guess);  // matching.cpp:220
}

// ===========================================================================
// a13  transformScore -> TransformationValidationEuclidean  [REF src/matching.cpp:259-268]
// [PCL-recall pcl/registration/impl/transformation_validation_euclidean.hpp]
// ===========================================================================
static double transform_score(const Cloud& source, const Cloud& target, const Mat4& t, double max_range)
{
  Grid tree;
  tree.build(target, (float)std::max(std::sqrt(std::max(max_range, 0.0)) * 0.25, 1e-3));
  long long Sd = 0, nr = 0;
#pragma omp parallel for schedule(dynamic, 1024) reduction(+ : Sd, nr)
  for (long long i = 0; i < (long long)source.size(); ++i) {
    float x, y, z;
    xform(t, source[i].x, source[i].y, source[i].z, x, y, z);
    int j;
    float dd;
    // "nn_dists[0] > max_range_" compares the SQUARED distance with the plain range (reference quirk)
    if (!tree.nn1(x, y, z, max_range, j, dd)) continue;
    Sd += em::to_fix((double)dd, MM3D_FIXD_SCALE);
    ++nr;
  }
  if (nr > 0) return ((double)Sd / MM3D_FIXD_SCALE) / (double)nr;
  return DBL_MAX;
}

// ===========================================================================
// a14/a15  pose graph  [REF src/graph.cpp:7-175, src/map_merging.cpp:137-186]
// (array-based restatement; tests compare it with the reference's own graph.cpp
// built unmodified into oracle/_ref/libgraph_ref.so)
// ===========================================================================
struct Estimate {
  size_t s, t;
  Mat4 transform;
  double confidence;
};

struct DSU {
  std::vector<size_t> parent, size, rank;
  explicit DSU(size_t n) : parent(n), size(n, 1), rank(n, 0)
  {
    for (size_t i = 0; i < n; ++i) parent[i] = i;
  }
  size_t find(size_t e)
  {
    size_t s = e;
    while (s != parent[s]) s = parent[s];
    while (e != parent[e]) { size_t nx = parent[e]; parent[e] = s; e = nx; }
    return s;
  }
  size_t merge(size_t a, size_t b)
  {
    if (rank[a] < rank[b]) { parent[a] = b; size[b] += size[a]; return b; }
    if (rank[b] < rank[a]) { parent[b] = a; size[a] += size[b]; return a; }
    parent[a] = b; rank[b]++; size[b] += size[a];
    return b;
  }
};

static size_t n_nodes(const std::vector<Estimate>& e)
{
  size_t n = 0;
  for (const auto& x : e) n = std::max({n, x.s + 1, x.t + 1});
  return n;
}

static std::vector<Estimate> largest_component(const std::vector<Estimate>& est, double thr)
{
  std::vector<Estimate> sub;
  const size_t n = n_nodes(est);
  if (n == 0) return sub;
  DSU comps(n);
  for (const auto& e : est) {
    if (e.confidence < thr) continue;
    const size_t a = comps.find(e.s), b = comps.find(e.t);
    if (a != b) comps.merge(a, b);
  }
  const size_t max_comp = (size_t)(std::max_element(comps.size.begin(), comps.size.end()) - comps.size.begin());
  for (const auto& e : est)
    if (comps.find(e.s) == max_comp) sub.push_back(e);
  return sub;
}

struct Edge {
  size_t from, to;
  double weight;
  bool operator>(const Edge& o) const { return weight > o.weight; }
};

static void bfs(const std::vector<std::vector<Edge>>& adj, size_t from, const std::function<void(const Edge&)>& body)
{
  std::vector<bool> was(adj.size(), false);
  std::queue<size_t> q;
  was[from] = true;
  q.push(from);
  while (!q.empty()) {
    const size_t v = q.front();
    q.pop();
    for (const Edge& e : adj[v])
      if (!was[e.to]) {
        body(e);
        was[e.to] = true;
        q.push(e.to);
      }
  }
}

static void max_spanning_tree(const std::vector<Estimate>& est, std::vector<std::vector<Edge>>& tree, std::vector<size_t>& centers)
{
  const size_t n = n_nodes(est);
  std::vector<Edge> edges;
  for (const auto& e : est) edges.push_back(Edge{e.s, e.t, e.confidence});
  DSU comps(n);
  tree.assign(n, {});
  std::vector<size_t> power(n, 0);
  std::sort(edges.begin(), edges.end(), std::greater<Edge>());  // same libstdc++ introsort as the reference
  for (const Edge& e : edges) {
    const size_t a = comps.find(e.from), b = comps.find(e.to);
    if (a != b) {
      comps.merge(a, b);
      tree[e.from].push_back(Edge{e.from, e.to, e.weight});
      tree[e.to].push_back(Edge{e.to, e.from, e.weight});
      power[e.from]++;
      power[e.to]++;
    }
  }
  std::vector<size_t> leafs;
  for (size_t i = 0; i < n; ++i)
    if (power[i] == 1) leafs.push_back(i);
  std::vector<size_t> max_d(n, 0), cur;
  for (size_t l : leafs) {
    cur.assign(n, 0);
    bfs(tree, l, [&](const Edge& e) { cur[e.to] = cur[e.from] + 1; });
    for (size_t j = 0; j < n; ++j) max_d[j] = std::max(max_d[j], cur[j]);
  }
  centers.clear();
  if (n == 0) return;
  size_t mm = max_d[0];
  for (size_t i = 1; i < n; ++i) mm = std::min(mm, max_d[i]);
  for (size_t i = 0; i < n; ++i)
    if (max_d[i] == mm) centers.push_back(i);
}

static Mat4 get_transform(const std::vector<Estimate>& est, size_t from, size_t to)
{
  for (const auto& e : est) {
    if (e.s == from && e.t == to) return inverse(e.transform);
    if (e.s == to && e.t == from) return e.transform;
  }
  return Mat4::zero();
}

static std::vector<Mat4> global_transforms(const std::vector<Estimate>& pairwise, double thr, size_t* ref_frame = nullptr)
{
  std::vector<Estimate> comp = largest_component(pairwise, thr);
  std::vector<std::vector<Edge>> tree;
  std::vector<size_t> centers;
  max_spanning_tree(comp, tree, centers);
  const size_t nodes = n_nodes(pairwise);
  std::vector<Mat4> g(nodes, Mat4::zero());
  if (centers.empty()) return g;  // fence: reference indexes centers[0] on an empty vector (UB)
  const size_t ref = centers[0];
  if (ref_frame) *ref_frame = ref;
  g[ref] = Mat4::identity();
  bfs(tree, ref, [&](const Edge& e) { g[e.to] = mul(g[e.from], get_transform(comp, e.from, e.to)); });
  return g;
}

// ===========================================================================
// a2  estimateMapsTransforms  [REF src/map_merging.cpp:188-275]
// ===========================================================================
struct Params {
  double resolution, descriptor_radius;
  int32_t outliers_min_neighbours;
  double normal_radius;
  int32_t keypoint_type;
  double keypoint_threshold;
  int32_t descriptor_type, estimation_method, refine_transform;
  double inlier_threshold, max_correspondence_distance;
  int32_t max_iterations;
  uint64_t matching_k;
  double transform_epsilon, confidence_threshold, output_resolution;
};

struct PairResult {
  int i, j;
  Mat4 t;
  double confidence;
  int n_corr, n_inliers;
};

struct StageTimes {
  double t[10];  // downsample, outliers, normals, keypoints, descriptors, matching, ransac, icp, score, graph
};

static double now_s()
{
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct MapFeatures {
  Cloud cloud;
  Normals normals;
  Cloud keypoints;
  std::vector<float> desc;
  int dim = 33;
};

static void map_features(const Cloud& in, const Params& p, MapFeatures& f, StageTimes* st)
{
  double t0 = now_s();
  f.cloud = voxel_grid(in, (float)p.resolution);
  double t1 = now_s();
  f.cloud = radius_outlier_removal(f.cloud, p.descriptor_radius, p.outliers_min_neighbours, nullptr, nullptr);
  double t2 = now_s();
  f.normals = surface_normals(f.cloud, p.normal_radius);
  double t3 = now_s();
  if (p.keypoint_type == 1)  // Keypoint::HARRIS: radius = normal_radius (map_merging.cpp:233)
    f.keypoints = harris_keypoints(f.cloud, f.normals, (float)p.keypoint_threshold, (float)p.normal_radius);
  else
    f.keypoints = sift_keypoints(f.cloud, (float)p.resolution, 3, 3, (float)p.keypoint_threshold, 0);
  double t4 = now_s();
  if (p.descriptor_type == 4) f.desc = shot_descriptors(f.cloud, f.normals, f.keypoints, p.descriptor_radius);
  else if (p.descriptor_type == 0) f.desc = pfh_descriptors(f.cloud, f.normals, f.keypoints, p.descriptor_radius);
  else if (p.descriptor_type == 1) f.desc = pfhrgb_descriptors(f.cloud, f.normals, f.keypoints, p.descriptor_radius);
  else if (p.descriptor_type == 3) f.desc = rsd_descriptors(f.cloud, f.normals, f.keypoints, p.descriptor_radius);
  else if (p.descriptor_type == 5) f.desc = sc3d_descriptors(f.cloud, f.normals, f.keypoints, p.descriptor_radius);
  else f.desc = fpfh_descriptors(f.cloud, f.normals, f.keypoints, p.descriptor_radius);
  f.dim = p.descriptor_type == 4 ? 1344 : (p.descriptor_type == 0 ? 125 : (p.descriptor_type == 1 ? 250 : (p.descriptor_type == 3 ? 2 : (p.descriptor_type == 5 ? 1980 : 33))));
  double t5 = now_s();
  if (st) {
    st->t[0] += t1 - t0; st->t[1] += t2 - t1; st->t[2] += t3 - t2; st->t[3] += t4 - t3; st->t[4] += t5 - t4;
  }
}

static GlibcRand g_pipeline_rand;

static PairResult register_pair(const MapFeatures& a, const MapFeatures& b, int i, int j, const Params& p, StageTimes* st)
{
  if (p.estimation_method == 1) {  // EstimationMethod::SAC_IA (matching.cpp:242-247)
    PairResult r;
    r.i = i; r.j = j;
    double t0 = now_s();
    Mat4 t = sac_ia_transform(a.keypoints, a.desc, b.keypoints, b.desc, a.dim, p.inlier_threshold, p.max_correspondence_distance,
                              p.max_iterations, g_pipeline_rand);
    double t2 = now_s();
    if (p.refine_transform) t = icp_refine(a.cloud, b.cloud, t, p.max_correspondence_distance, p.max_iterations, p.transform_epsilon);
    double t3 = now_s();
    const double score = transform_score(a.cloud, b.cloud, t, p.max_correspondence_distance);
    double t4 = now_s();
    r.t = t; r.confidence = 1. / score; r.n_corr = 0; r.n_inliers = 0;
    if (st) { st->t[6] += t2 - t0; st->t[7] += t3 - t2; st->t[8] += t4 - t3; }
    return r;
  }
  PairResult r;
  r.i = i; r.j = j;
  double t0 = now_s();
  std::vector<Corr> corr = find_correspondences(a.desc.data(), a.keypoints.size(), b.desc.data(), b.keypoints.size(), a.dim, p.matching_k);
  double t1 = now_s();
  std::vector<int> inl;
  Mat4 t = ransac_transform(a.keypoints, b.keypoints, corr, p.inlier_threshold, inl);
  double t2 = now_s();
  if (p.refine_transform) t = icp_refine(a.cloud, b.cloud, t, p.max_correspondence_distance, p.max_iterations, p.transform_epsilon);
  double t3 = now_s();
  const double score = transform_score(a.cloud, b.cloud, t, p.max_correspondence_distance);
  double t4 = now_s();
  r.t = t;
  r.confidence = 1. / score;
  r.n_corr = (int)corr.size();
  r.n_inliers = (int)inl.size();
  if (st) { st->t[5] += t1 - t0; st->t[6] += t2 - t1; st->t[7] += t3 - t2; st->t[8] += t4 - t3; }
  return r;
}

}  // namespace orc

// ===========================================================================
// C interface for ctypes (tests / bench only)
// ===========================================================================
using namespace orc;

static Cloud to_cloud(const float* pts, uint64_t n)
{
  Cloud c(n);
  if (n) memcpy(c.data(), pts, n * sizeof(P4));
  return c;
}
static float* dup_f(const void* src, size_t bytes)
{
  float* p = (float*)malloc(bytes ? bytes : 1);
  if (bytes) memcpy(p, src, bytes);
  return p;
}
static void to_colmajor(const Mat4& t, float* out)
{
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) out[c * 4 + r] = t.m[r * 4 + c];
}
static Mat4 from_colmajor(const float* in)
{
  Mat4 t;
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) t.m[r * 4 + c] = in[c * 4 + r];
  return t;
}

extern "C" {

void orc_free(void* p) { free(p); }

// host threads the per-point loops use (OpenMP); results do not depend on it.  n <= 0: all cores.  Returns the setting.
int orc_set_threads(int n)
{
#ifdef _OPENMP
  if (n <= 0) n = omp_get_num_procs();
  omp_set_num_threads(n);
  return n;
#else
  (void)n;
  return 1;
#endif
}

int orc_uses_libm()
{
#ifdef ORACLE_LIBM
  return 1;
#else
  return 0;
#endif
}

// info: min_b[3], div_b[3], passthrough
int orc_downsample(const float* pts, uint64_t n, double resolution, float** out, uint64_t* n_out, int32_t* info, uint32_t** keys)
{
  VoxelInfo vi;
  std::vector<uint32_t> k;
  Cloud o = voxel_grid(to_cloud(pts, n), (float)resolution, &vi, keys ? &k : nullptr);
  *out = dup_f(o.data(), o.size() * sizeof(P4));
  *n_out = o.size();
  if (info) {
    for (int a = 0; a < 3; ++a) { info[a] = vi.min_b[a]; info[3 + a] = vi.div_b[a]; }
    info[6] = vi.passthrough;
  }
  if (keys) *keys = (uint32_t*)dup_f(k.data(), k.size() * 4);
  return 0;
}

int orc_remove_outliers(const float* pts, uint64_t n, double radius, int min_nb, float** out, uint64_t* n_out, int32_t** kept, int32_t** counts)
{
  std::vector<int> kp, cnt;
  Cloud o = radius_outlier_removal(to_cloud(pts, n), radius, min_nb, &kp, &cnt);
  *out = dup_f(o.data(), o.size() * sizeof(P4));
  *n_out = o.size();
  if (kept) *kept = (int32_t*)dup_f(kp.data(), kp.size() * 4);
  if (counts) *counts = (int32_t*)dup_f(cnt.data(), cnt.size() * 4);
  return 0;
}

int orc_normals(const float* pts, uint64_t n, double radius, float** out)
{
  Normals nm = surface_normals(to_cloud(pts, n), radius);
  *out = dup_f(nm.data(), nm.size() * sizeof(N4));
  return 0;
}

int orc_sift(const float* pts, uint64_t n, double min_scale, int n_octaves, int n_scales, double min_contrast, int order_mode,
             float** kp, uint64_t* nk, float** dog0, uint64_t* n_dog0, float** kp_scales)
{
  SiftDebug dbg;
  std::vector<float> sc;
  Cloud o = sift_keypoints(to_cloud(pts, n), (float)min_scale, n_octaves, n_scales, (float)min_contrast, order_mode, &dbg, &sc);
  *kp = dup_f(o.data(), o.size() * sizeof(P4));
  *nk = o.size();
  if (dog0) { *dog0 = dup_f(dbg.dog.data(), dbg.dog.size() * 4); *n_dog0 = dbg.dog.size(); }
  if (kp_scales) *kp_scales = dup_f(sc.data(), sc.size() * 4);
  return 0;
}

int orc_harris(const float* pts, uint64_t n, const float* normals, double threshold, double radius, float** kp, uint64_t* nk, float** response,
               float** unrefined, uint64_t* n_unrefined)
{
  Normals nm(n);
  if (n) memcpy(nm.data(), normals, n * sizeof(N4));
  std::vector<float> resp;
  Cloud unref;
  Cloud o = harris_keypoints(to_cloud(pts, n), nm, (float)threshold, (float)radius, response ? &resp : nullptr, unrefined ? &unref : nullptr);
  *kp = dup_f(o.data(), o.size() * sizeof(P4));
  *nk = o.size();
  if (response) *response = dup_f(resp.data(), resp.size() * 4);
  if (unrefined) { *unrefined = dup_f(unref.data(), unref.size() * sizeof(P4)); *n_unrefined = unref.size(); }
  return 0;
}

int orc_fpfh(const float* pts, uint64_t n, const float* normals, const float* kp_in, uint64_t nk_in, double radius, float** kp_out,
             uint64_t* nk_out, float** desc, float** spfh)
{
  Cloud surf = to_cloud(pts, n);
  Normals nm(n);
  if (n) memcpy(nm.data(), normals, n * sizeof(N4));
  Cloud kp = to_cloud(kp_in, nk_in);
  std::vector<float> sp;
  std::vector<float> d = fpfh_descriptors(surf, nm, kp, radius, spfh ? &sp : nullptr);
  *kp_out = dup_f(kp.data(), kp.size() * sizeof(P4));
  *nk_out = kp.size();
  *desc = dup_f(d.data(), d.size() * 4);
  if (spfh) *spfh = dup_f(sp.data(), sp.size() * 4);
  return 0;
}

int orc_pfh(const float* pts, uint64_t n, const float* normals, const float* kp_in, uint64_t nk_in, double radius, float** kp_out,
            uint64_t* nk_out, float** desc)
{
  Cloud surf = to_cloud(pts, n);
  Normals nm(n);
  if (n) memcpy(nm.data(), normals, n * sizeof(N4));
  Cloud kp = to_cloud(kp_in, nk_in);
  std::vector<float> d = pfh_descriptors(surf, nm, kp, radius);
  *kp_out = dup_f(kp.data(), kp.size() * sizeof(P4));
  *nk_out = kp.size();
  *desc = dup_f(d.data(), d.size() * 4);
  return 0;
}

int orc_pfhrgb(const float* pts, uint64_t n, const float* normals, const float* kp_in, uint64_t nk_in, double radius, float** kp_out,
               uint64_t* nk_out, float** desc)
{
  Cloud surf = to_cloud(pts, n);
  Normals nm(n);
  if (n) memcpy(nm.data(), normals, n * sizeof(N4));
  Cloud kp = to_cloud(kp_in, nk_in);
  std::vector<float> d = pfhrgb_descriptors(surf, nm, kp, radius);
  *kp_out = dup_f(kp.data(), kp.size() * sizeof(P4));
  *nk_out = kp.size();
  *desc = dup_f(d.data(), d.size() * 4);
  return 0;
}

int orc_rsd(const float* pts, uint64_t n, const float* normals, const float* kp_in, uint64_t nk_in, double radius, float** kp_out,
            uint64_t* nk_out, float** desc)
{
  Cloud surf = to_cloud(pts, n);
  Normals nm(n);
  if (n) memcpy(nm.data(), normals, n * sizeof(N4));
  Cloud kp = to_cloud(kp_in, nk_in);
  std::vector<float> d = rsd_descriptors(surf, nm, kp, radius);
  *kp_out = dup_f(kp.data(), kp.size() * sizeof(P4));
  *nk_out = kp.size();
  *desc = dup_f(d.data(), d.size() * 4);
  return 0;
}

int orc_sc3d(const float* pts, uint64_t n, const float* normals, const float* kp_in, uint64_t nk_in, double radius, float** kp_out,
             uint64_t* nk_out, float** desc)
{
  Cloud surf = to_cloud(pts, n);
  Normals nm(n);
  if (n) memcpy(nm.data(), normals, n * sizeof(N4));
  Cloud kp = to_cloud(kp_in, nk_in);
  std::vector<float> d = sc3d_descriptors(surf, nm, kp, radius);
  *kp_out = dup_f(kp.data(), kp.size() * sizeof(P4));
  *nk_out = kp.size();
  *desc = dup_f(d.data(), d.size() * 4);
  return 0;
}

int orc_shot(const float* pts, uint64_t n, const float* normals, const float* kp_in, uint64_t nk_in, double radius, float** kp_out,
             uint64_t* nk_out, float** desc, float** rf)
{
  Cloud surf = to_cloud(pts, n);
  Normals nm(n);
  if (n) memcpy(nm.data(), normals, n * sizeof(N4));
  Cloud kp = to_cloud(kp_in, nk_in);
  std::vector<float> rfv;
  std::vector<float> d = shot_descriptors(surf, nm, kp, radius, rf ? &rfv : nullptr);
  *kp_out = dup_f(kp.data(), kp.size() * sizeof(P4));
  *nk_out = kp.size();
  *desc = dup_f(d.data(), d.size() * 4);
  if (rf) *rf = dup_f(rfv.data(), rfv.size() * 4);
  return 0;
}

int orc_match(const float* ds, uint64_t ns, const float* dt, uint64_t nt, int dim, uint64_t k, int32_t** pairs, float** dist, uint64_t* nc)
{
  std::vector<Corr> c = find_correspondences(ds, ns, dt, nt, dim, k);
  int32_t* p = (int32_t*)malloc(c.size() * 8 + 1);
  float* d = (float*)malloc(c.size() * 4 + 1);
  for (size_t i = 0; i < c.size(); ++i) { p[2 * i] = c[i].q; p[2 * i + 1] = c[i].m; d[i] = c[i].d; }
  *pairs = p; *dist = d; *nc = c.size();
  return 0;
}

int orc_knn(const float* a, uint64_t na, const float* b, uint64_t nb, int dim, int k, int32_t* idx, float* dist)
{
  std::vector<int> i;
  std::vector<float> d;
  knn_bruteforce(a, na, b, nb, dim, k, i, d);
  memcpy(idx, i.data(), i.size() * 4);
  memcpy(dist, d.data(), d.size() * 4);
  return 0;
}

// dbg: iterations, best_count ; dbg_d: sample_dist_thresh ; best_model colmajor
int orc_ransac(const float* kps, uint64_t ns, const float* kpt, uint64_t nt, const int32_t* pairs, const float* dist, uint64_t nc,
               double inlier_threshold, float* T, int32_t** inliers, uint64_t* n_inl, int32_t* dbg, double* dbg_d, float* best_model)
{
  Cloud s = to_cloud(kps, ns), t = to_cloud(kpt, nt);
  std::vector<Corr> c(nc);
  for (size_t i = 0; i < nc; ++i) { c[i].q = pairs[2 * i]; c[i].m = pairs[2 * i + 1]; c[i].d = dist ? dist[i] : 0.f; }
  std::vector<int> inl;
  RansacDebug d;
  d.best_model = Mat4::identity();
  Mat4 r = ransac_transform(s, t, c, inlier_threshold, inl, &d);
  to_colmajor(r, T);
  if (inliers) *inliers = (int32_t*)dup_f(inl.data(), inl.size() * 4);
  if (n_inl) *n_inl = inl.size();
  if (dbg) { dbg[0] = d.iterations; dbg[1] = d.best_count; }
  if (dbg_d) *dbg_d = d.sample_dist_thresh;
  if (best_model) to_colmajor(d.best_model, best_model);
  return 0;
}

// rand_calls (in/out): number of rand() calls consumed before / after this pair (0 = fresh process)
int orc_sac_ia(const float* kps, uint64_t ns, const float* ds, const float* kpt, uint64_t nt, const float* dt, int dim, double min_sample_distance,
               double max_corr_dist, int max_iterations, uint64_t* rand_calls, float* T, float** errors, uint64_t* n_errors)
{
  Cloud s = to_cloud(kps, ns), t = to_cloud(kpt, nt);
  std::vector<float> sd(ds, ds + ns * dim), td(dt, dt + nt * dim);
  GlibcRand rng;
  const uint64_t skip = rand_calls ? *rand_calls : 0;
  for (uint64_t i = 0; i < skip; ++i) rng.rand_();
  std::vector<float> err;
  Mat4 r = sac_ia_transform(s, sd, t, td, dim, min_sample_distance, max_corr_dist, max_iterations, rng, &err);
  to_colmajor(r, T);
  if (rand_calls) *rand_calls = rng.calls;
  if (errors) { *errors = dup_f(err.data(), err.size() * 4); *n_errors = err.size(); }
  return 0;
}

int orc_glibc_rand(int n, int32_t* out)
{
  GlibcRand rng;
  for (int i = 0; i < n; ++i) out[i] = rng.rand_();
  return 0;
}

int orc_icp(const float* src, uint64_t ns, const float* tgt, uint64_t nt, const float* T0, double max_dist, int max_it, double eps,
            float* T, int32_t* dbg, long long** sums, uint64_t* n_sums)
{
  IcpDebug d;
  Mat4 r = icp_refine(to_cloud(src, ns), to_cloud(tgt, nt), from_colmajor(T0), max_dist, max_it, eps, &d);
  to_colmajor(r, T);
  if (dbg) { dbg[0] = d.iterations; dbg[1] = d.converged; }
  if (sums) { *sums = (long long*)dup_f(d.sums.data(), d.sums.size() * 8); *n_sums = d.sums.size(); }
  return 0;
}

int orc_score(const float* src, uint64_t ns, const float* tgt, uint64_t nt, const float* T, double max_range, double* score)
{
  *score = transform_score(to_cloud(src, ns), to_cloud(tgt, nt), from_colmajor(T), max_range);
  return 0;
}

// est: n x (src, tgt) ; transforms n x 16 colmajor ; conf n.  out: nodes x 16 colmajor
int orc_global_transforms(int n, const int32_t* st, const float* transforms, const double* conf, double thr, float* out, int* n_out,
                          int* ref_frame)
{
  std::vector<Estimate> e(n);
  for (int i = 0; i < n; ++i) {
    e[i].s = st[2 * i]; e[i].t = st[2 * i + 1];
    e[i].transform = from_colmajor(transforms + 16 * i);
    e[i].confidence = conf[i];
  }
  size_t ref = 0;
  std::vector<Mat4> g = global_transforms(e, thr, &ref);
  for (size_t i = 0; i < g.size(); ++i) to_colmajor(g[i], out + 16 * i);
  *n_out = (int)g.size();
  if (ref_frame) *ref_frame = (int)ref;
  return 0;
}

// graph pieces for comparison with oracle/_ref: component membership flags, tree edges, centers
int orc_graph(int n, const int32_t* st, const double* conf, double thr, int32_t* in_component, int32_t* tree_edges, int* n_tree_edges,
              int32_t* centers, int* n_centers)
{
  std::vector<Estimate> e(n);
  for (int i = 0; i < n; ++i) { e[i].s = st[2 * i]; e[i].t = st[2 * i + 1]; e[i].transform = Mat4::identity(); e[i].confidence = conf[i]; }
  std::vector<Estimate> comp = largest_component(e, thr);
  // mark membership by order-preserving scan
  size_t c = 0;
  for (int i = 0; i < n; ++i) {
    in_component[i] = 0;
    if (c < comp.size() && comp[c].s == e[i].s && comp[c].t == e[i].t) { in_component[i] = 1; ++c; }
  }
  std::vector<std::vector<Edge>> tree;
  std::vector<size_t> cen;
  max_spanning_tree(comp, tree, cen);
  int ne = 0;
  for (size_t v = 0; v < tree.size(); ++v)
    for (const Edge& ed : tree[v]) { tree_edges[2 * ne] = (int)ed.from; tree_edges[2 * ne + 1] = (int)ed.to; ++ne; }
  *n_tree_edges = ne;
  for (size_t i = 0; i < cen.size(); ++i) centers[i] = (int)cen[i];
  *n_centers = (int)cen.size();
  return 0;
}

// Full path.  pair_out: P x (i, j, n_corr, n_inliers); pair_T: P x 16 colmajor; pair_conf: P.
// max_pairs < 0: all pairs; otherwise only the first max_pairs pairs are registered (bounded CPU baseline).
int orc_estimate_maps_transforms(int n_maps, const float* const* clouds, const uint64_t* n_points, const Params* p, float* out_transforms,
                                 int* n_out, double* stage_times, int32_t* pair_out, float* pair_T, double* pair_conf, int* n_pairs,
                                 int max_pairs)
{
  *n_out = 0;
  if (n_pairs) *n_pairs = 0;
  if (n_maps == 0) return 0;
  if (n_maps == 1) {
    to_colmajor(Mat4::identity(), out_transforms);
    *n_out = 1;
    return 0;
  }
  StageTimes st;
  memset(&st, 0, sizeof(st));
  g_pipeline_rand.seed(1);  // a fresh process: rand() has never been called
  std::vector<MapFeatures> f(n_maps);
  for (int i = 0; i < n_maps; ++i) map_features(to_cloud(clouds[i], n_points[i]), *p, f[i], &st);
  std::vector<Estimate> est;
  int np = 0;
  for (int i = 0; i < n_maps - 1; ++i)
    for (int j = i + 1; j < n_maps; ++j) {
      if (f[i].keypoints.size() > 0 && f[j].keypoints.size() > 0) {
        if (max_pairs >= 0 && np >= max_pairs) continue;
        PairResult r = register_pair(f[i], f[j], i, j, *p, &st);
        Estimate e;
        e.s = i; e.t = j; e.transform = r.t; e.confidence = r.confidence;
        est.push_back(e);
        if (pair_out) { pair_out[4 * np] = i; pair_out[4 * np + 1] = j; pair_out[4 * np + 2] = r.n_corr; pair_out[4 * np + 3] = r.n_inliers; }
        if (pair_T) to_colmajor(r.t, pair_T + 16 * np);
        if (pair_conf) pair_conf[np] = r.confidence;
        ++np;
      }
    }
  if (n_pairs) *n_pairs = np;
  double t0 = now_s();
  std::vector<Mat4> g = global_transforms(est, p->confidence_threshold);
  st.t[9] += now_s() - t0;
  for (size_t i = 0; i < g.size(); ++i) to_colmajor(g[i], out_transforms + 16 * i);
  *n_out = (int)g.size();
  if (stage_times) memcpy(stage_times, st.t, sizeof(st.t));
  return 0;
}

// a16 composeMaps [REF src/map_merging.cpp:277-305]; returns 1 = nullptr result, 2 = size mismatch (throws)
int orc_compose_maps(int n_maps, const float* const* clouds, const uint64_t* n_points, int n_transforms, const float* transforms,
                     double resolution, float** out, uint64_t* n_out)
{
  *out = nullptr;
  *n_out = 0;
  if (n_maps == 0) return 1;
  if (n_maps != n_transforms) return 2;
  Cloud result;
  for (int i = 0; i < n_maps; ++i) {
    Mat4 t = from_colmajor(transforms + 16 * i);
    bool zero = true;  // Eigen isZero(): all |a_ij| <= 1e-5
    for (int k = 0; k < 16; ++k)
      if (!(std::fabs(t.m[k]) <= 1e-5f)) zero = false;
    if (zero) continue;
    const P4* c = (const P4*)clouds[i];
    for (uint64_t k = 0; k < n_points[i]; ++k) {
      P4 o = c[k];
      xform(t, c[k].x, c[k].y, c[k].z, o.x, o.y, o.z);
      result.push_back(o);
    }
  }
  Cloud o = voxel_grid(result, (float)resolution);
  *out = dup_f(o.data(), o.size() * sizeof(P4));
  *n_out = o.size();
  return 0;
}

// exact_math probes
float orc_em_expf(float x) { return mm3d::em::expf_(x); }
float orc_em_atan2f(float y, float x) { return mm3d::em::atan2f_(y, x); }
long long orc_em_to_fix32_pos(float v) { return mm3d::em::to_fix32_pos(v); }
long long orc_em_to_fix(double v, double scale) { return mm3d::em::to_fix(v, scale); }
float orc_em_cosf(float x) { return mm3d::em::cosf_small_(x); }
float orc_em_sinf(float x) { return mm3d::em::sinf_small_(x); }
void orc_em_svd3d(const double* A, double* U, double* s, double* V) { mm3d::em::svd3<double>(A, U, s, V); }
void orc_em_svd3f(const float* A, float* U, float* s, float* V) { mm3d::em::svd3<float>(A, U, s, V); }
void orc_umeyama_d(const double* src, const double* dst, uint64_t n, double* Rt)
{
  std::vector<double> s(src, src + 3 * n), d(dst, dst + 3 * n);
  umeyama<double>(s, d, n, Rt);
}
void orc_umeyama_f(const float* src, const float* dst, uint64_t n, float* Rt)
{
  std::vector<float> s(src, src + 3 * n), d(dst, dst + 3 * n);
  umeyama<float>(s, d, n, Rt);
}

}  // extern "C"
