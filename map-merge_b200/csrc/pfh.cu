// pfh.cu — PFH descriptors (5 x 5 x 5 = 125 bins), the reference's DEFAULT descriptor_type, and PFHRGB (125 geometric +
// 125 colour-ratio bins).
//   <- pcl::PFHEstimation / pcl::PFHRGBEstimation via map_merge_3d/src/dispatch_descriptors.h:38-39, src/features.cpp:99-150
//   [PCL-recall pcl/features/impl/pfh.hpp computePointPFHSignature, pcl/features/impl/pfhrgb.hpp computePointPFHRGBSignature]
// One persistent CTA per keypoint at a time: thread 0 gathers the radius neighbourhood (ascending index) into shared
// memory, all threads split the n(n-1)/2 pairs, votes are integer atomics on a shared 125-bin histogram, and because
// every vote adds the same float (100 / #pairs) the float histogram is the vote count replayed as additions.
#include <algorithm>
#include <cmath>

#include "mm3d_internal.cuh"
#include "pair_features.cuh"

namespace mm3d {

namespace {

constexpr int PB = 256;
constexpr int PFH_SMEM_NB = 1024;   // neighbours staged in shared memory
constexpr int PFH_MAX_NB = 16384;   // hard cap (global scratch per CTA)

struct PfhJob {
  GridView g;
  const float4* normals;
  const float4* kp;
  int nk;
  float* desc_raw;  // nk x 125 (PFH) / nk x 250 (PFHRGB)
  uint32_t* valid;  // nk
};

// colour-ratio bin of PFHRGB by integer quotient c1 / c2 (0..255): ratio r = q, folded to -1/r when > 1, bin =
// floor(5 * ((r + 1.0) * 0.5)) clamped — evaluated on the host with the literal double formula
struct ColourBins {
  unsigned char b[256];
};

struct PfhWork {
  int job, kp;
};

template <bool RGB>
__global__ void __launch_bounds__(PB) pfh_kernel(const PfhJob* __restrict__ jobs, const PfhWork* __restrict__ work, int n_work, float r2, int rv,
                                                BinTable bins, ColourBins cbins, int* __restrict__ scratch, int* __restrict__ overflow)
{
  constexpr int DIM = RGB ? 250 : 125;
  __shared__ float4 spt[PFH_SMEM_NB];
  __shared__ float4 snm[PFH_SMEM_NB];
  __shared__ unsigned int hist[DIM];
  __shared__ float thr[3][12];
  __shared__ int s_n;
  if (threadIdx.x < 36) (&thr[0][0])[threadIdx.x] = (&bins.t[0][0])[threadIdx.x];
  int* my_scratch = scratch + (size_t)blockIdx.x * PFH_MAX_NB;
  for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
    const PfhJob& j = jobs[work[w].job];
    const int t = work[w].kp;
    __syncthreads();
    if (threadIdx.x < DIM) hist[threadIdx.x] = 0;
    if (threadIdx.x == 0) {
      const float4 c = j.kp[t];
      int n = 0;
      for_each_in_radius(j.g, true, c.x, c.y, c.z, r2, rv, [&](int s, const float4& p, float) {
        if (n < PFH_SMEM_NB) {
          spt[n] = p;
          snm[n] = j.normals[j.g.orig ? j.g.orig[s] : s];
        }
        if (n < PFH_MAX_NB) my_scratch[n] = s;
        ++n;
      });
      s_n = n;
      if (n > PFH_MAX_NB) atomicExch(overflow, 1);
    }
    __syncthreads();
    const int n = min(s_n, PFH_MAX_NB);
    const bool in_smem = n <= PFH_SMEM_NB;
    const long long total = (long long)n * (n - 1) / 2;
    for (long long e = threadIdx.x; e < total; e += PB) {
      // pair (i, j < i) of the triangular enumeration
      int i = (int)((1.0 + sqrt(1.0 + 8.0 * (double)e)) * 0.5);
      while ((long long)i * (i - 1) / 2 > e) --i;
      while ((long long)(i + 1) * i / 2 <= e) ++i;
      const int jj = (int)(e - (long long)i * (i - 1) / 2);
      float4 p1, n1, p2, n2;
      if (in_smem) {
        p1 = spt[i]; n1 = snm[i]; p2 = spt[jj]; n2 = snm[jj];
      } else {
        const int si = my_scratch[i], sj = my_scratch[jj];
        p1 = j.g.pts[si]; n1 = j.normals[j.g.orig ? j.g.orig[si] : si];
        p2 = j.g.pts[sj]; n2 = j.normals[j.g.orig ? j.g.orig[sj] : sj];
      }
      float f1, f2, f3;
      // "if (!compute(RGB)PairFeatures (...)) continue;": a degenerate pair does not vote
      const bool ok = RGB ? pair_features_noswap(p1, n1, p2, n2, &f1, &f2, &f3) : pair_features(p1, n1, p2, n2, &f1, &f2, &f3);
      if (!ok) continue;
      const int h1 = lookup_bin(thr[0], 5, f1, 5.0f * 0.15915494f, 3.14159274f);
      const int h2 = lookup_bin(thr[1], 5, f2, 2.5f, 1.0f);
      const int h3 = lookup_bin(thr[2], 5, f3, 2.5f, 1.0f);
      atomicAdd(&hist[h1 + 5 * h2 + 25 * h3], 1u);
      if (RGB) {
        const unsigned int c1 = __float_as_uint(p1.w), c2 = __float_as_uint(p2.w);  // PointXYZRGB::rgba bits
        int hc[3];
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          const unsigned int a1 = (c1 >> (16 - 8 * ch)) & 0xffu, a2 = (c2 >> (16 - 8 * ch)) & 0xffu;  // r, g, b
          hc[ch] = cbins.b[a2 != 0 ? a1 / a2 : 1u];  // integer division as in pcl::computeRGBPairFeatures; x / 0 -> 1.0f
        }
        atomicAdd(&hist[125 + hc[0] + 5 * hc[1] + 25 * hc[2]], 1u);
      }
    }
    __syncthreads();
    if (threadIdx.x < DIM) {
      const float hist_incr = 100.0f / (float)((unsigned long long)n * (unsigned long long)(n - 1) / 2ull);
      const unsigned int cnt = hist[threadIdx.x];
      float h = 0.f;
      for (unsigned int k = 0; k < cnt; ++k) h += hist_incr;
      j.desc_raw[(size_t)t * DIM + threadIdx.x] = h;
    }
    // PFH marks an empty neighbourhood with NaN (dropped by the caller); PFHRGB leaves zeros (kept)
    if (threadIdx.x == 0) j.valid[t] = ((RGB || s_n > 0) && s_n <= PFH_MAX_NB) ? 1u : 0u;
  }
}

struct PfhEmitJob {
  const float4* kp;
  const float* desc_raw;
  const uint32_t* flags;
  const uint32_t* pos;
  float4* kp_out;
  float* desc_out;
  int nk;
};
__global__ void __launch_bounds__(128) pfh_emit_kernel(const PfhEmitJob* __restrict__ jobs, int dim)
{
  const PfhEmitJob& j = jobs[blockIdx.y];
  const int kp = blockIdx.x;
  if (kp >= j.nk || !j.flags[kp]) return;
  const uint32_t o = j.pos[kp];
  for (int t = threadIdx.x; t < dim; t += blockDim.x) j.desc_out[(size_t)o * dim + t] = j.desc_raw[(size_t)kp * dim + t];
  if (threadIdx.x == 0) j.kp_out[o] = j.kp[kp];
}

}  // namespace

void pfh_batch(Ctx& c, const std::vector<CloudView>& clouds, const std::vector<DIndex>& idx, const std::vector<const float4*>& normals,
               std::vector<DCloud>& keypoints, double radius, std::vector<DBuf<float>>& desc, bool rgb)
{
  const int M = (int)clouds.size();
  const int dim = rgb ? 250 : 125;
  desc.clear();
  desc.resize(M);
  if (M == 0) return;
  std::vector<int> nks(M);
  int mxk = 0, totalk = 0;
  std::vector<Seg> segk(M);
  std::vector<PfhWork> work;
  for (int m = 0; m < M; ++m) {
    nks[m] = keypoints[m].n;
    segk[m].off = totalk;
    segk[m].n = nks[m];
    totalk += nks[m];
    mxk = std::max(mxk, nks[m]);
    for (int k = 0; k < nks[m]; ++k) work.push_back(PfhWork{m, k});
  }
  if (totalk == 0) {
    for (int m = 0; m < M; ++m) { keypoints[m].n = 0; keypoints[m].pts.release(); }
    return;
  }
  DBuf<uint32_t> flags(c, totalk), pos(c, totalk);
  std::vector<DBuf<float>> raw(M);
  std::vector<PfhJob> jobs(M);
  for (int m = 0; m < M; ++m) {
    raw[m].alloc(c, (size_t)nks[m] * dim);
    jobs[m] = PfhJob{idx[m].v, normals[m], keypoints[m].pts.p, nks[m], raw[m].p, flags.p + segk[m].off};
  }
  DBuf<PfhJob> dj = to_device(c, jobs);
  DBuf<PfhWork> dw = to_device(c, work);
  const int grid = std::min(totalk, 148 * 4);  // persistent: one CTA per keypoint at a time, 4 CTAs per SM
  DBuf<int> scratch(c, (size_t)grid * PFH_MAX_NB);
  DBuf<int> overflow(c, 1);
  overflow.zero(c);
  static const BinTable bins = make_bin_table(5);
  static const ColourBins cbins = []() {
    ColourBins cb;
    for (int q = 0; q < 256; ++q) {
      float r = (float)q;
      if (r > 1.0f) r = -1.0f / r;
      cb.b[q] = (unsigned char)pf_bin_f23(r, 5);
    }
    return cb;
  }();
  const float r2 = (float)(radius * radius);
  const int rv = (int)std::ceil(radius / (double)idx[0].v.leaf) + 1;
  { double b = 0; for (int m = 0; m < M; ++m) b += 32.0 * clouds[m].n + (16.0 + 4.0 * dim) * nks[m]; MM_BYTES(c, b); }
  if (rgb) MM_LAUNCH(c, pfh_kernel<true>, grid, PB, 0, dj.p, dw.p, totalk, r2, rv, bins, cbins, scratch.p, overflow.p);
  else MM_LAUNCH(c, pfh_kernel<false>, grid, PB, 0, dj.p, dw.p, totalk, r2, rv, bins, cbins, scratch.p, overflow.p);
  int h_over = 0;
  overflow.download(c, &h_over, 1);
  std::vector<int> totals;
  scan_flags_batch(c, flags.p, pos.p, segk, totals);  // synchronises
  if (h_over) throw std::runtime_error("PFH / PFHRGB: a keypoint has more than 16384 neighbours inside the feature radius");
  std::vector<DCloud> kept(M);
  std::vector<PfhEmitJob> ej(M);
  for (int m = 0; m < M; ++m) {
    kept[m].n = totals[m];
    kept[m].pts.alloc(c, totals[m]);
    desc[m].alloc(c, (size_t)totals[m] * dim);
    ej[m] = PfhEmitJob{keypoints[m].pts.p, raw[m].p, flags.p + segk[m].off, pos.p + segk[m].off, kept[m].pts.p, desc[m].p, nks[m]};
  }
  DBuf<PfhEmitJob> dej = to_device(c, ej);
  MM_LAUNCH(c, pfh_emit_kernel, dim3(mxk, M), 128, 0, dej.p, dim);
  for (int m = 0; m < M; ++m) keypoints[m] = std::move(kept[m]);
}

}  // namespace mm3d
