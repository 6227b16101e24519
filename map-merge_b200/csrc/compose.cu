// compose.cu — composeMaps (map_merge_3d/src/map_merging.cpp:277-305) sharded over ranks.
// Every rank transforms its own maps, the ranks agree on the global voxel geometry (all-reduced bounding box) and on
// balanced key-range splitters (all-reduced key histogram), each rank partitions its points by owning rank — stably, so
// that concatenating the received chunks in rank order reproduces the reference's (map, point) order inside every voxel —
// and voxel-grids the points it owns.  Ranges are ordered by key, so the per-rank outputs concatenate to the reference's
// output order.  RAW POINTS are exchanged, not partial sums: the centroid sums stay bit-identical to one big voxel grid.
#include <algorithm>
#include <cfloat>
#include <cmath>

#include "mm3d_internal.cuh"

namespace mm3d {

namespace {

struct KeyGeom {
  float inv_leaf;
  int min_b[3];
  int div_b[3];
  unsigned long long bucket_width;
};

__device__ __forceinline__ unsigned int global_key(const KeyGeom& g, const float4& p)
{
  const int i0 = (int)(floorf(p.x * g.inv_leaf) - (float)g.min_b[0]);
  const int i1 = (int)(floorf(p.y * g.inv_leaf) - (float)g.min_b[1]);
  const int i2 = (int)(floorf(p.z * g.inv_leaf) - (float)g.min_b[2]);
  return (unsigned int)(i0 + i1 * g.div_b[0] + i2 * g.div_b[0] * g.div_b[1]);
}

__global__ void __launch_bounds__(256) bbox1_kernel(const float4* __restrict__ pts, int n, float* __restrict__ out /*6, pre-initialised*/)
{
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 p = pts[i];
    if (!(isfinite(p.x) && isfinite(p.y) && isfinite(p.z))) continue;  // left out of the box, dropped by the voxel grid
    mn[0] = fminf(mn[0], p.x); mx[0] = fmaxf(mx[0], p.x);
    mn[1] = fminf(mn[1], p.y); mx[1] = fmaxf(mx[1], p.y);
    mn[2] = fminf(mn[2], p.z); mx[2] = fmaxf(mx[2], p.z);
  }
  for (int k = 0; k < 3; ++k) {
    for (int o = 16; o > 0; o >>= 1) {
      mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
      mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
    }
  }
  if ((threadIdx.x & 31) == 0) {
    // float atomics through the ordered-int trick
    for (int k = 0; k < 3; ++k) {
      int* lo = (int*)&out[k];
      int* hi = (int*)&out[3 + k];
      if (mn[k] >= 0.f) atomicMin(lo, __float_as_int(mn[k])); else atomicMax((unsigned int*)lo, __float_as_uint(mn[k]));
      if (mx[k] >= 0.f) atomicMax(hi, __float_as_int(mx[k])); else atomicMin((unsigned int*)hi, __float_as_uint(mx[k]));
    }
  }
}

__global__ void __launch_bounds__(256) key_hist_kernel(const float4* __restrict__ pts, int n, KeyGeom g, int n_buckets,
                                                      unsigned long long* __restrict__ hist)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long b = (unsigned long long)global_key(g, pts[i]) / g.bucket_width;
  atomicAdd(&hist[min((unsigned long long)(n_buckets - 1), b)], 1ull);
}

__global__ void __launch_bounds__(256) dest_kernel(const float4* __restrict__ pts, int n, KeyGeom g, int n_buckets, const int* __restrict__ splitters,
                                                  int n_ranks, uint32_t* __restrict__ dest, uint32_t* __restrict__ idx,
                                                  unsigned long long* __restrict__ counts)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned long long b = (unsigned long long)global_key(g, pts[i]) / g.bucket_width;
  if (b > (unsigned long long)(n_buckets - 1)) b = n_buckets - 1;
  int r = 0;
  while (r + 1 < n_ranks && (int)b >= splitters[r + 1]) ++r;  // splitters[r] <= b < splitters[r + 1]
  dest[i] = (uint32_t)r;
  idx[i] = (uint32_t)i;
  atomicAdd(&counts[r], 1ull);
}

__global__ void __launch_bounds__(256) gather4_kernel(const float4* __restrict__ src, const uint32_t* __restrict__ order, int n,
                                                     float4* __restrict__ dst)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[order[i]];
}

}  // namespace

KeyGeomHost compose_geometry(const float* bbox, double resolution, int n_buckets)
{
  KeyGeomHost h;
  memset(&h, 0, sizeof(h));
  const float leaf = (float)resolution;
  const float inv = 1.0f / leaf;
  h.inv_leaf = inv;
  h.passthrough = !(leaf > 0.0f);
  if (!h.passthrough) {
    const long long dx = (long long)((bbox[3] - bbox[0]) * inv) + 1, dy = (long long)((bbox[4] - bbox[1]) * inv) + 1,
                    dz = (long long)((bbox[5] - bbox[2]) * inv) + 1;
    if (dx * dy * dz > 2147483647LL) h.passthrough = 1;  // pcl::VoxelGrid: "Leaf size is too small" -> output = input
  }
  if (!h.passthrough) {
    long long cells = 1;
    for (int k = 0; k < 3; ++k) {
      h.min_b[k] = (int)std::floor(bbox[k] * inv);
      h.div_b[k] = (int)std::floor(bbox[3 + k] * inv) - h.min_b[k] + 1;
      cells *= h.div_b[k];
    }
    h.bucket_width = (unsigned long long)((cells + n_buckets - 1) / n_buckets);
    if (h.bucket_width == 0) h.bucket_width = 1;
  }
  return h;
}

void compose_bbox(Ctx& c, const DCloud& cloud, float* bbox_host)
{
  float init[6] = {FLT_MAX, FLT_MAX, FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX};
  DBuf<float> d(c, 6);
  d.upload(c, init, 6);
  if (cloud.n) {
    const int blocks = std::max(1, std::min((cloud.n + 255) / 256, 148 * 8));
    MM_LAUNCH(c, bbox1_kernel, blocks, 256, 0, cloud.pts.p, cloud.n, d.p);
  }
  d.download(c, bbox_host, 6);
  c.sync();
}

static KeyGeom to_dev(const KeyGeomHost& h)
{
  KeyGeom g;
  g.inv_leaf = h.inv_leaf;
  for (int k = 0; k < 3; ++k) { g.min_b[k] = h.min_b[k]; g.div_b[k] = h.div_b[k]; }
  g.bucket_width = h.bucket_width;
  return g;
}

void compose_histogram(Ctx& c, const DCloud& cloud, const KeyGeomHost& geom, int n_buckets, unsigned long long* hist_host)
{
  DBuf<unsigned long long> h(c, n_buckets);
  h.zero(c);
  if (cloud.n) MM_LAUNCH(c, key_hist_kernel, (cloud.n + 255) / 256, 256, 0, cloud.pts.p, cloud.n, to_dev(geom), n_buckets, h.p);
  h.download(c, hist_host, n_buckets);
  c.sync();
}

void compose_partition(Ctx& c, const DCloud& cloud, const KeyGeomHost& geom, int n_buckets, const std::vector<int>& splitters, int n_ranks,
                       unsigned long long* counts_host, float4* out_dev)
{
  for (int r = 0; r < n_ranks; ++r) counts_host[r] = 0;
  if (cloud.n == 0) return;
  DBuf<int> ds = to_device(c, splitters);
  DBuf<unsigned long long> dc(c, n_ranks);
  dc.zero(c);
  DBuf<uint32_t> dest(c, cloud.n), idx(c, cloud.n), dest2(c, cloud.n), idx2(c, cloud.n);
  MM_LAUNCH(c, dest_kernel, (cloud.n + 255) / 256, 256, 0, cloud.pts.p, cloud.n, to_dev(geom), n_buckets, ds.p, n_ranks, dest.p, idx.p, dc.p);
  int nbits = 1;
  while ((1 << nbits) < n_ranks) ++nbits;
  uint32_t *ks, *vs;
  radix_sort_pairs_batch(c, dest.p, idx.p, dest2.p, idx2.p, {Seg{0, cloud.n}}, nbits, &ks, &vs);  // stable: keeps (map, point) order
  MM_LAUNCH(c, gather4_kernel, (cloud.n + 255) / 256, 256, 0, cloud.pts.p, vs, cloud.n, out_dev);
  dc.download(c, counts_host, n_ranks);
  c.sync();
}

}  // namespace mm3d
