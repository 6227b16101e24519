// sc3d.cu — 3D shape context descriptors (12 azimuth x 11 elevation x 15 radius = 1980 bins)
//   <- pcl::ShapeContext3DEstimation<PointXYZRGB, Normal, ShapeContext1980> via map_merge_3d/src/dispatch_descriptors.h:47-48,
//      src/features.cpp:99-150   [PCL-recall pcl/features/impl/3dsc.hpp initCompute + computePoint, 3dsc.h constructor]
// Kernels: local point density of every surface point (thread per point), a has-neighbours flag per keypoint (its rank
// among the keypoints that have neighbours selects its three draws of the estimator's mt19937 stream, generated on the
// host), and the descriptor kernel: thread per keypoint, neighbours in ascending index, float adds into the keypoint's
// private 1980-float row.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <random>

#include "mm3d_internal.cuh"

namespace mm3d {

namespace {

constexpr int CB = 128;
constexpr int SC_D = 1980;
constexpr int AZ_BINS = 12, EL_BINS = 11, RAD_BINS = 15;

struct Sc3dTables {
  float radii[RAD_BINS + 1], theta_div[EL_BINS + 1], phi_div[AZ_BINS + 1], volume_lut[SC_D];
};

// ShapeContext3DEstimation::initCompute, min_radius_ = 0.1 (host libm, as in PCL)
void make_tables(double search_radius, Sc3dTables& t)
{
  const double min_radius = 0.1;
  const float azimuth_interval = 360.0f / (float)AZ_BINS, elevation_interval = 180.0f / (float)EL_BINS;
  for (int j = 0; j <= RAD_BINS; ++j)
    t.radii[j] = (float)(exp(log(min_radius) + (((float)j / (float)RAD_BINS) * log(search_radius / min_radius))));
  for (int k = 0; k <= EL_BINS; ++k) t.theta_div[k] = (float)k * elevation_interval;
  for (int l = 0; l <= AZ_BINS; ++l) t.phi_div[l] = (float)l * azimuth_interval;
  const float integr_phi = (t.phi_div[1] * 0.017453293f) - (t.phi_div[0] * 0.017453293f);
  const float e = 1.0f / 3.0f;
  for (int j = 0; j < RAD_BINS; ++j) {
    const float integr_r = (t.radii[j + 1] * t.radii[j + 1] * t.radii[j + 1] / 3.0f) - (t.radii[j] * t.radii[j] * t.radii[j] / 3.0f);
    for (int k = 0; k < EL_BINS; ++k) {
      const float integr_theta = cosf(t.theta_div[k] * 0.017453293f) - cosf(t.theta_div[k + 1] * 0.017453293f);
      const float V = integr_phi * integr_theta * integr_r;
      for (int l = 0; l < AZ_BINS; ++l) t.volume_lut[(l * EL_BINS * RAD_BINS) + k * RAD_BINS + j] = 1.0f / powf(V, e);
    }
  }
}

struct Sc3dJob {
  GridView g;
  const float4* normals;
  const float4* kp;
  int nk;
  int* density;     // per surface slot
  float* desc_raw;  // nk x 1980
  uint32_t* has_nb; // nk: the keypoint has at least one neighbour (= it draws random numbers and its row is finite)
  const uint32_t* rank;  // exclusive scan of has_nb
};

__global__ void __launch_bounds__(256) sc3d_density_kernel(const Sc3dJob* __restrict__ jobs, float r2, int rv)
{
  const Sc3dJob& j = jobs[blockIdx.y];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < j.g.n;
  const float4 p = live ? j.g.pts[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  int n = 0;
  for_each_in_radius(j.g, live, p.x, p.y, p.z, r2, rv, [&](int, const float4&, float) { ++n; });
  if (live) j.density[i] = n;
}

__global__ void __launch_bounds__(CB) sc3d_flag_kernel(const Sc3dJob* __restrict__ jobs, float r2, int rv)
{
  const Sc3dJob& j = jobs[blockIdx.y];
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = k < j.nk;
  const float4 c = live ? j.kp[k] : make_float4(0.f, 0.f, 0.f, 0.f);
  const bool finite = isfinite(c.x) && isfinite(c.y) && isfinite(c.z);
  int n = 0;
  for_each_in_radius(j.g, live && finite, c.x, c.y, c.z, r2, rv, [&](int, const float4&, float) { ++n; });
  if (live) j.has_nb[k] = n > 0 ? 1u : 0u;
}

__device__ __forceinline__ bool is_zero(float v) { return fabsf(v - 0.0f) < FLT_MIN; }  // pcl::utils::equal (v, 0.0f)

__device__ __forceinline__ void normalize3(float* v)  // Eigen normalize(): untouched when the squared norm is 0
{
  const float z = (v[0] * v[0] + v[1] * v[1]) + v[2] * v[2];
  if (z > 0.0f) {
    const float nrm = sqrtf(z);
    v[0] /= nrm; v[1] /= nrm; v[2] /= nrm;
  }
}

__global__ void __launch_bounds__(CB) sc3d_kernel(const Sc3dJob* __restrict__ jobs, float r2, int rv, const Sc3dTables* __restrict__ tables,
                                                 const float* __restrict__ draws)
{
  __shared__ float radii[RAD_BINS + 1], theta_div[EL_BINS + 1], phi_div[AZ_BINS + 1];
  if (threadIdx.x <= RAD_BINS) radii[threadIdx.x] = tables->radii[threadIdx.x];
  if (threadIdx.x <= EL_BINS) theta_div[threadIdx.x] = tables->theta_div[threadIdx.x];
  if (threadIdx.x <= AZ_BINS) phi_div[threadIdx.x] = tables->phi_div[threadIdx.x];
  __syncthreads();
  const Sc3dJob& j = jobs[blockIdx.y];
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= j.nk || !j.has_nb[k]) return;
  const float4 o = j.kp[k];
  // nearest surface point -> normal
  float bd = FLT_MAX;
  int bslot = -1;
  for_each_in_radius(j.g, true, o.x, o.y, o.z, r2, rv, [&](int s, const float4&, float d2) {
    if (d2 < bd) { bd = d2; bslot = s; }  // ascending index: the first minimum is the lowest index
  });
  const float4 n4 = j.normals[j.g.orig ? j.g.orig[bslot] : bslot];
  const float normal[3] = {n4.x, n4.y, n4.z};
  const float* dr = draws + (size_t)j.rank[k] * 3;
  float x_axis[3] = {dr[0], dr[1], dr[2]};
  if (!is_zero(normal[2])) x_axis[2] = -(normal[0] * x_axis[0] + normal[1] * x_axis[1]) / normal[2];
  else if (!is_zero(normal[1])) x_axis[1] = -(normal[0] * x_axis[0] + normal[2] * x_axis[2]) / normal[1];
  else if (!is_zero(normal[0])) x_axis[0] = -(normal[1] * x_axis[1] + normal[2] * x_axis[2]) / normal[0];
  normalize3(x_axis);
  float* desc = j.desc_raw + (size_t)k * SC_D;
  for (int t = 0; t < SC_D; ++t) desc[t] = 0.0f;
  for_each_in_radius(j.g, true, o.x, o.y, o.z, r2, rv, [&](int s, const float4& q, float d2) {
    if (is_zero(d2)) return;
    const float r = sqrtf(d2);
    // pcl::geometry::project (neighbour, origin, normal, proj); proj -= origin
    const float po[3] = {q.x - o.x, q.y - o.y, q.z - o.z};
    const float lambda = (normal[0] * po[0] + normal[1] * po[1]) + normal[2] * po[2];
    float proj[3] = {q.x - lambda * normal[0], q.y - lambda * normal[1], q.z - lambda * normal[2]};
    proj[0] -= o.x; proj[1] -= o.y; proj[2] -= o.z;
    normalize3(proj);
    const float cr[3] = {x_axis[1] * proj[2] - x_axis[2] * proj[1], x_axis[2] * proj[0] - x_axis[0] * proj[2],
                         x_axis[0] * proj[1] - x_axis[1] * proj[0]};
    const float cr_norm = sqrtf((cr[0] * cr[0] + cr[1] * cr[1]) + cr[2] * cr[2]);
    float phi = em::atan2f_(cr_norm, (x_axis[0] * proj[0] + x_axis[1] * proj[1]) + x_axis[2] * proj[2]) * 57.29578f;  // pcl::rad2deg
    phi = ((cr[0] * normal[0] + cr[1] * normal[1]) + cr[2] * normal[2]) < 0.f ? (360.0f - phi) : phi;
    float no[3] = {po[0], po[1], po[2]};
    normalize3(no);
    float theta = (normal[0] * no[0] + normal[1] * no[1]) + normal[2] * no[2];
    theta = (float)em::acos_d_((double)fminf(1.0f, fmaxf(-1.0f, theta))) * 57.29578f;
    int bj = 0, bk = 0, bl = 0;
    for (int rad = 1; rad < RAD_BINS + 1; ++rad)
      if (r <= radii[rad]) { bj = rad - 1; break; }
    for (int ang = 1; ang < EL_BINS + 1; ++ang)
      if (theta <= theta_div[ang]) { bk = ang - 1; break; }
    for (int ang = 1; ang < AZ_BINS + 1; ++ang)
      if (phi <= phi_div[ang]) { bl = ang - 1; break; }
    const int dens = j.density[s];
    if (dens == 0) return;
    const int bin = (bl * EL_BINS * RAD_BINS) + (bk * RAD_BINS) + bj;
    const float w = (1.0f / (float)dens) * tables->volume_lut[bin];
    desc[bin] += w;
  });
}

// the row is valid when every value is finite (DefaultPointRepresentation<ShapeContext1980>::isValid)
__global__ void __launch_bounds__(128) sc3d_valid_kernel(const Sc3dJob* __restrict__ jobs, uint32_t* __restrict__ const* valid)
{
  const Sc3dJob& j = jobs[blockIdx.y];
  const int k = blockIdx.x;
  if (k >= j.nk) return;
  __shared__ int bad;
  if (threadIdx.x == 0) bad = j.has_nb[k] ? 0 : 1;
  __syncthreads();
  if (j.has_nb[k])
    for (int t = threadIdx.x; t < SC_D; t += blockDim.x)
      if (!isfinite(j.desc_raw[(size_t)k * SC_D + t])) bad = 1;
  __syncthreads();
  if (threadIdx.x == 0) valid[blockIdx.y][k] = bad ? 0u : 1u;
}

struct Sc3dEmitJob {
  const float4* kp;
  const float* desc_raw;
  const uint32_t* flags;
  const uint32_t* pos;
  float4* kp_out;
  float* desc_out;
  int nk;
};
__global__ void __launch_bounds__(256) sc3d_emit_kernel(const Sc3dEmitJob* __restrict__ jobs)
{
  const Sc3dEmitJob& j = jobs[blockIdx.y];
  const int kp = blockIdx.x;
  if (kp >= j.nk || !j.flags[kp]) return;
  const uint32_t o = j.pos[kp];
  for (int t = threadIdx.x; t < SC_D; t += blockDim.x) j.desc_out[(size_t)o * SC_D + t] = j.desc_raw[(size_t)kp * SC_D + t];
  if (threadIdx.x == 0) j.kp_out[o] = j.kp[kp];
}

}  // namespace

void sc3d_batch(Ctx& c, const std::vector<CloudView>& clouds, const std::vector<DIndex>& idx, const std::vector<const float4*>& normals,
                std::vector<DCloud>& keypoints, double radius, std::vector<DBuf<float>>& desc)
{
  const int M = (int)clouds.size();
  desc.clear();
  desc.resize(M);
  if (M == 0) return;
  std::vector<int> nks(M);
  int mxk = 0, totalk = 0, mxn = 0;
  std::vector<Seg> segk(M);
  for (int m = 0; m < M; ++m) {
    nks[m] = keypoints[m].n;
    segk[m].off = totalk;
    segk[m].n = nks[m];
    totalk += nks[m];
    mxk = std::max(mxk, nks[m]);
    mxn = std::max(mxn, clouds[m].n);
  }
  // initCompute fails when search_radius_ < min_radius_ (0.1): PCL returns an empty cloud, every keypoint is dropped
  if (totalk == 0 || radius < 0.1) {
    for (int m = 0; m < M; ++m) { keypoints[m].n = 0; keypoints[m].pts.release(); }
    return;
  }
  Sc3dTables ht;
  make_tables(radius, ht);
  DBuf<Sc3dTables> dt(c, 1);
  dt.upload(c, &ht, 1);
  // one estimator per computeLocalDescriptors call: boost::uniform_01<boost::mt19937>, seed 12345u, three draws per keypoint
  std::vector<float> draws((size_t)mxk * 3);
  {
    std::mt19937 rng(12345u);
    for (float& d : draws) d = (float)((double)rng() * (1.0 / 4294967296.0));
  }
  DBuf<float> dd = to_device(c, draws);
  DBuf<uint32_t> has_nb(c, totalk), rank(c, totalk), flags(c, totalk), pos(c, totalk);
  std::vector<DBuf<float>> raw(M);
  std::vector<DBuf<int>> dens(M);
  std::vector<Sc3dJob> jobs(M);
  std::vector<uint32_t*> vptr(M);
  for (int m = 0; m < M; ++m) {
    raw[m].alloc(c, (size_t)nks[m] * SC_D);
    dens[m].alloc(c, clouds[m].n);
    jobs[m] = Sc3dJob{idx[m].v, normals[m], keypoints[m].pts.p, nks[m], dens[m].p, raw[m].p, has_nb.p + segk[m].off, rank.p + segk[m].off};
    vptr[m] = flags.p + segk[m].off;
  }
  DBuf<Sc3dJob> dj = to_device(c, jobs);
  DBuf<uint32_t*> dv = to_device(c, vptr);
  const float r2 = (float)(radius * radius);
  const int rv = (int)std::ceil(radius / (double)idx[0].v.leaf) + 1;
  const double density_radius = 0.2;  // point_density_radius_
  const float dr2 = (float)(density_radius * density_radius);
  const int drv = (int)std::ceil(density_radius / (double)idx[0].v.leaf) + 1;
  MM_LAUNCH(c, sc3d_density_kernel, dim3((mxn + 255) / 256, M), 256, 0, dj.p, dr2, drv);
  MM_LAUNCH(c, sc3d_flag_kernel, dim3((mxk + CB - 1) / CB, M), CB, 0, dj.p, r2, rv);
  std::vector<int> with_nb;
  scan_flags_batch(c, has_nb.p, rank.p, segk, with_nb);
  { double b = 0; for (int m = 0; m < M; ++m) b += 36.0 * clouds[m].n + (16.0 + 4.0 * SC_D) * nks[m]; MM_BYTES(c, b); }
  MM_LAUNCH(c, sc3d_kernel, dim3((mxk + CB - 1) / CB, M), CB, 0, dj.p, r2, rv, dt.p, dd.p);
  MM_LAUNCH(c, sc3d_valid_kernel, dim3(mxk, M), 128, 0, dj.p, dv.p);
  std::vector<int> totals;
  scan_flags_batch(c, flags.p, pos.p, segk, totals);
  std::vector<DCloud> kept(M);
  std::vector<Sc3dEmitJob> ej(M);
  for (int m = 0; m < M; ++m) {
    kept[m].n = totals[m];
    kept[m].pts.alloc(c, totals[m]);
    desc[m].alloc(c, (size_t)totals[m] * SC_D);
    ej[m] = Sc3dEmitJob{keypoints[m].pts.p, raw[m].p, flags.p + segk[m].off, pos.p + segk[m].off, kept[m].pts.p, desc[m].p, nks[m]};
  }
  DBuf<Sc3dEmitJob> dej = to_device(c, ej);
  MM_LAUNCH(c, sc3d_emit_kernel, dim3(mxk, M), 256, 0, dej.p);
  for (int m = 0; m < M; ++m) keypoints[m] = std::move(kept[m]);
}

}  // namespace mm3d
