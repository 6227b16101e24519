// icp.cu — K11 point-to-point ICP refine and K12 Euclidean transform score,
// all pairs per launch (blockIdx.y = pair).
//   <- map_merge_3d/src/matching.cpp:196-221 (pcl::IterativeClosestPoint, TransformationEstimationSVD,
//      DefaultConvergenceCriteria) and :259-268 (pcl::registration::TransformationValidationEuclidean)
// One iteration = one kernel that does the nearest-neighbour search, the distance
// gate and the {n, sum p, sum q, sum q p^T, sum d^2} reduction for every active
// pair, plus a one-warp-per-pair kernel that solves Umeyama (3x3 SVD), updates the
// transform and tests convergence on the device.  Sums are fixed-point int64, so
// the reduction is order-independent and matches the CPU checker bit for bit.
#include <algorithm>
#include <cmath>

#include "mm3d_internal.cuh"

namespace mm3d {

namespace {

constexpr int IB = 128;
constexpr int NSUM = 17;  // n, Sp[3], Sq[3], Sqp[9], Sd

struct IcpState {
  float final_t[16];  // accumulated ICP transform (row-major)
  float step[16];     // transformation_ of the last iteration
  double prev_mse;
  int iterations;
  int active;
  int converged;
  int pad;
};

struct IcpJob {
  GridView tgt;      // target index
  const float4* src; // source cloud
  int ns;
  float4* work;      // transformed source (input_transformed)
  long long* sums;   // NSUM
  IcpState* st;
  int* ticket;       // blocks of this pair that finished the current iteration
  float t0[16];      // initial guess, row-major
  long long* sums_log;  // optional max_log x NSUM
  int* nn_slot;      // per source point: the target slot matched in the previous iteration (-1 = none)
};

__global__ void __launch_bounds__(256) icp_init_kernel(const IcpJob* __restrict__ jobs)
{
  const IcpJob& j = jobs[blockIdx.y];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < j.ns; i += gridDim.x * blockDim.x) {
    const float4 p = j.src[i];
    float4 o;
    em::transform_point(j.t0, p.x, p.y, p.z, &o.x, &o.y, &o.z);  // pcl::transformPointCloud(source, initial_guess)
    o.w = p.w;
    j.work[i] = o;
    j.nn_slot[i] = -1;
  }
}

__device__ __forceinline__ void block_accumulate(long long* vals, int nvals, long long* gsums)
{
  __shared__ long long sh[IB / 32][NSUM];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int k = 0; k < nvals; ++k) {
    const long long v = warp_sum_ll(vals[k]);
    if (lane == 0) sh[w][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < nvals) {
    long long s = 0;
    for (int ww = 0; ww < IB / 32; ++ww) s += sh[ww][threadIdx.x];
    if (s != 0) atomicAdd((unsigned long long*)&gsums[threadIdx.x], (unsigned long long)s);
  }
}

__device__ void mat4_mul(const float* a, const float* b, float* r)
{
  for (int i = 0; i < 4; ++i)
    for (int c = 0; c < 4; ++c) {
      float acc = a[i * 4 + 0] * b[0 * 4 + c];
      acc += a[i * 4 + 1] * b[1 * 4 + c];
      acc += a[i * 4 + 2] * b[2 * 4 + c];
      acc += a[i * 4 + 3] * b[3 * 4 + c];
      r[i * 4 + c] = acc;
    }
}

// Umeyama from the finished reductions, transform update and pcl::registration::DefaultConvergenceCriteria for one pair;
// runs on one thread of the LAST block of that pair to finish (so an ICP iteration is a single kernel)
__device__ void icp_solve(const IcpJob& j, int max_iterations, double rotation_threshold, double translation_threshold, int max_log,
                          int* __restrict__ n_active)
{
  IcpState& st = *j.st;
  long long s[NSUM];
  for (int k = 0; k < NSUM; ++k) {
    s[k] = (long long)atomicExch((unsigned long long*)&j.sums[k], 0ull);  // read the other blocks' atomics, reset for the next iteration
  }
  if (j.sums_log && st.iterations < max_log)
    for (int k = 0; k < NSUM; ++k) j.sums_log[(size_t)st.iterations * NSUM + k] = s[k];
  const long long cnt = s[0];
  if (cnt < 3) {  // "Not enough correspondences found": stop, not converged
    st.active = 0;
    st.converged = 0;
    atomicSub(n_active, 1);
    return;
  }
  const double n = (double)cnt;
  double pm[3], qm[3];
  for (int a = 0; a < 3; ++a) {
    pm[a] = ((double)s[1 + a] / MM3D_FIX1_SCALE) / n;
    qm[a] = ((double)s[4 + a] / MM3D_FIX1_SCALE) / n;
  }
  float sigma[9], pmf[3], qmf[3];
  for (int a = 0; a < 3; ++a) {
    pmf[a] = (float)pm[a];
    qmf[a] = (float)qm[a];
    for (int b = 0; b < 3; ++b) sigma[a * 3 + b] = (float)(((double)s[7 + a * 3 + b] / MM3D_FIX2_SCALE) / n - qm[a] * pm[b]);
  }
  float T[16];
  em::umeyama_from_sigma<float>(sigma, pmf, qmf, T);
  float nf[16];
  mat4_mul(T, st.final_t, nf);
  for (int k = 0; k < 16; ++k) {
    st.final_t[k] = nf[k];
    st.step[k] = T[k];
  }
  st.iterations += 1;
  // pcl::registration::DefaultConvergenceCriteria::hasConverged
  bool conv = false;
  if (st.iterations >= max_iterations) {
    conv = true;
  } else {
    const double cos_angle = 0.5 * (double)(T[0] + T[5] + T[10] - 1);
    const double translation_sqr = (double)(T[3] * T[3] + T[7] * T[7] + T[11] * T[11]);
    if (cos_angle >= rotation_threshold && translation_sqr <= translation_threshold) {
      conv = true;
    } else {
      const double mse = ((double)s[16] / MM3D_FIXD_SCALE) / n;
      if (fabs(mse - st.prev_mse) < 1e-12) conv = true;
      else st.prev_mse = mse;
    }
  }
  if (conv) {
    st.active = 0;
    st.converged = 1;
    atomicSub(n_active, 1);
  }
}

// the step transform of the previous iteration is applied on the fly
// (transformCloud(input_transformed, input_transformed, transformation_))
__global__ void __launch_bounds__(IB) icp_iteration_kernel(const IcpJob* __restrict__ jobs, double max_dist_sqr, int rv, int apply_step,
                                                           int max_iterations, double rotation_threshold, double translation_threshold,
                                                           int max_log, int* __restrict__ n_active)
{
  const IcpJob& j = jobs[blockIdx.y];
  if (!j.st->active) return;
  if (blockIdx.x * IB >= j.ns) return;
  long long v[NSUM];
#pragma unroll
  for (int k = 0; k < NSUM; ++k) v[k] = 0;
  const int i = blockIdx.x * IB + threadIdx.x;
  if (i < j.ns) {
    float4 p = j.work[i];
    if (apply_step) {
      float4 o;
      em::transform_point(j.st->step, p.x, p.y, p.z, &o.x, &o.y, &o.z);
      o.w = p.w;
      p = o;
      j.work[i] = p;
    }
    int idx;
    float d2;
    float4 q;
    int slot;
    const bool hit = nearest_bounded(j.tgt, p.x, p.y, p.z, max_dist_sqr, rv, &idx, &d2, &q, j.nn_slot[i], &slot);
    j.nn_slot[i] = slot;  // the transform moves little between iterations: last iteration's match seeds the next search
    if (hit) {
      const double pp[3] = {(double)p.x, (double)p.y, (double)p.z};
      const double qq[3] = {(double)q.x, (double)q.y, (double)q.z};
      v[0] = 1;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        v[1 + a] = em::to_fix(pp[a], MM3D_FIX1_SCALE);
        v[4 + a] = em::to_fix(qq[a], MM3D_FIX1_SCALE);
#pragma unroll
        for (int b = 0; b < 3; ++b) v[7 + a * 3 + b] = em::to_fix(qq[a] * pp[b], MM3D_FIX2_SCALE);
      }
      v[16] = em::to_fix((double)d2, MM3D_FIXD_SCALE);
    }
  }
  block_accumulate(v, NSUM, j.sums);
  // last block of this pair: solve + convergence test
  __shared__ int s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const int nblocks = (j.ns + IB - 1) / IB;
    const int ticket = atomicAdd(j.ticket, 1);
    s_last = (ticket == nblocks - 1) ? 1 : 0;
    if (s_last) {
      *j.ticket = 0;
      __threadfence();
      icp_solve(j, max_iterations, rotation_threshold, translation_threshold, max_log, n_active);
    }
  }
}

// ---------------------------------------------------------------- score
struct ScoreJob {
  GridView tgt;
  const float4* src;
  int ns;
  float t[16];
  long long* sums;  // Sd, nr
  const int* nn_guess;  // optional: the slots matched in the last ICP iteration
};

__global__ void __launch_bounds__(IB) score_kernel(const ScoreJob* __restrict__ jobs, double max_range, int rv)
{
  const ScoreJob& j = jobs[blockIdx.y];
  if (blockIdx.x * IB >= j.ns) return;
  long long v[2] = {0, 0};
  const int i = blockIdx.x * IB + threadIdx.x;
  if (i < j.ns) {
    const float4 p = j.src[i];
    float x, y, z;
    em::transform_point(j.t, p.x, p.y, p.z, &x, &y, &z);
    int idx;
    float d2;
    float4 q;
    // the reference compares the squared distance with the plain range
    if (nearest_bounded(j.tgt, x, y, z, max_range, rv, &idx, &d2, &q, j.nn_guess ? j.nn_guess[i] : -1)) {
      v[0] = em::to_fix((double)d2, MM3D_FIXD_SCALE);
      v[1] = 1;
    }
  }
  block_accumulate(v, 2, j.sums);
}

}  // namespace

void icp_batch(Ctx& c, const std::vector<CloudView>& clouds, const std::vector<DIndex>& idx, const std::vector<PairJob>& jobs,
               const std::vector<const float*>& T0, double max_dist, int max_it, double eps, std::vector<IcpOut>& out,
               std::vector<std::vector<long long>>* sums_dbg, IcpNeighbours* nn_keep)
{
  const int P = (int)jobs.size();
  out.assign(P, IcpOut());
  if (nn_keep) nn_keep->offset.assign(P, -1);
  if (sums_dbg) { sums_dbg->clear(); sums_dbg->resize(P); }
  if (P == 0) return;
  const int max_log = sums_dbg ? 64 : 0;
  // A zero initial guess (failed RANSAC) stays zero through ICP: final * 0 == 0 (matching.cpp:220).
  std::vector<int> act;
  for (int p = 0; p < P; ++p) {
    bool zero = true;
    for (int k = 0; k < 16; ++k)
      if (T0[p][k] != 0.0f) zero = false;
    if (zero) {
      for (int k = 0; k < 16; ++k) out[p].T[k] = 0.0f;
      out[p].iterations = 0;
      out[p].converged = 0;
    } else {
      act.push_back(p);
    }
  }
  const int A = (int)act.size();
  if (A == 0) return;
  size_t tot = 0;
  int mx = 0;
  for (int a = 0; a < A; ++a) {
    tot += (size_t)clouds[jobs[act[a]].a].n;
    mx = std::max(mx, clouds[jobs[act[a]].a].n);
  }
  DBuf<float4> work(c, tot + 1);
  DBuf<int> nn_local;
  DBuf<int>& nn = nn_keep ? nn_keep->slots : nn_local;
  nn.alloc(c, tot + 1);
  DBuf<long long> sums(c, (size_t)A * NSUM);
  sums.zero(c);
  DBuf<int> tickets(c, A);
  tickets.zero(c);
  DBuf<long long> slog;
  if (max_log) {
    slog.alloc(c, (size_t)A * max_log * NSUM);
    slog.zero(c);
  }
  std::vector<IcpState> hst(A);
  for (int a = 0; a < A; ++a) {
    memset(&hst[a], 0, sizeof(IcpState));
    for (int k = 0; k < 16; ++k) hst[a].final_t[k] = hst[a].step[k] = (k % 5 == 0) ? 1.f : 0.f;
    hst[a].prev_mse = 1.7976931348623157e308;
    // an empty source launches no block, so nothing would ever retire the pair: it starts out finished and not converged
    // ("Not enough correspondences found", the transform stays the initial guess)
    hst[a].active = clouds[jobs[act[a]].a].n > 0 ? 1 : 0;
  }
  DBuf<IcpState> dst = to_device(c, hst);
  std::vector<IcpJob> ij(A);
  size_t off = 0;
  for (int a = 0; a < A; ++a) {
    const PairJob& pj = jobs[act[a]];
    ij[a].tgt = idx[pj.b].v;
    ij[a].src = clouds[pj.a].pts;
    ij[a].ns = clouds[pj.a].n;
    ij[a].work = work.p + off;
    ij[a].nn_slot = nn.p + off;
    if (nn_keep) nn_keep->offset[act[a]] = (long long)off;
    ij[a].sums = sums.p + (size_t)a * NSUM;
    ij[a].st = dst.p + a;
    ij[a].ticket = tickets.p + a;
    for (int k = 0; k < 16; ++k) ij[a].t0[k] = T0[act[a]][k];
    ij[a].sums_log = max_log ? slog.p + (size_t)a * max_log * NSUM : nullptr;
    off += (size_t)ij[a].ns;
  }
  DBuf<IcpJob> dij = to_device(c, ij);
  DBuf<int> dn_active(c, 1);
  int n_active = 0;
  for (int a = 0; a < A; ++a) n_active += hst[a].active;
  dn_active.upload(c, &n_active, 1);
  const int iblocks = std::max(1, std::min((mx + 255) / 256, 148 * 8));
  MM_LAUNCH(c, icp_init_kernel, dim3(iblocks, A), 256, 0, dij.p);
  const double max_dist_sqr = max_dist * max_dist;
  const float leaf = idx[jobs[act[0]].b].v.leaf;
  const int rv = (int)std::ceil(max_dist / (double)leaf) + 1;
  const dim3 grid(std::max(1, (mx + IB - 1) / IB), A);
  int it = 0;
  while (n_active > 0 && it < std::max(max_it, 1)) {
    { double b = 0; for (int a = 0; a < A; ++a) b += 16.0 * ((double)ij[a].ns * (it > 0 ? 2 : 1) + ij[a].tgt.n); MM_BYTES(c, b * ((double)n_active / A)); }
    MM_LAUNCH(c, icp_iteration_kernel, grid, IB, 0, dij.p, max_dist_sqr, rv, it > 0 ? 1 : 0, max_it, 1.0 - eps, eps, max_log, dn_active.p);
    dn_active.download(c, &n_active, 1);
    c.sync();
    ++it;
  }
  dst.download(c, hst.data(), A);
  std::vector<long long> hlog;
  if (max_log) {
    hlog.resize((size_t)A * max_log * NSUM);
    slog.download(c, hlog.data(), hlog.size());
  }
  c.sync();
  for (int a = 0; a < A; ++a) {
    IcpOut& o = out[act[a]];
    // icp.getFinalTransformation() * initial_guess
    const float* f = hst[a].final_t;
    const float* g = T0[act[a]];
    for (int i = 0; i < 4; ++i)
      for (int cc = 0; cc < 4; ++cc) {
        float acc = f[i * 4 + 0] * g[0 * 4 + cc];
        acc += f[i * 4 + 1] * g[1 * 4 + cc];
        acc += f[i * 4 + 2] * g[2 * 4 + cc];
        acc += f[i * 4 + 3] * g[3 * 4 + cc];
        o.T[i * 4 + cc] = acc;
      }
    o.iterations = hst[a].iterations;
    o.converged = hst[a].converged;
    if (sums_dbg) {
      const int nl = std::min(hst[a].iterations + (hst[a].converged ? 0 : 1), max_log);
      (*sums_dbg)[act[a]].assign(hlog.begin() + (size_t)a * max_log * NSUM, hlog.begin() + ((size_t)a * max_log + nl) * NSUM);
    }
  }
}

void score_batch(Ctx& c, const std::vector<CloudView>& clouds, const std::vector<DIndex>& idx, const std::vector<PairJob>& jobs,
                 const std::vector<const float*>& T, double max_range, std::vector<double>& scores, const IcpNeighbours* nn_guess)
{
  const int P = (int)jobs.size();
  scores.assign(P, 1.7976931348623157e308);
  if (P == 0) return;
  DBuf<long long> sums(c, (size_t)P * 2);
  sums.zero(c);
  std::vector<ScoreJob> sj(P);
  int mx = 0;
  for (int p = 0; p < P; ++p) {
    sj[p].tgt = idx[jobs[p].b].v;
    sj[p].src = clouds[jobs[p].a].pts;
    sj[p].ns = clouds[jobs[p].a].n;
    for (int k = 0; k < 16; ++k) sj[p].t[k] = T[p][k];
    sj[p].sums = sums.p + (size_t)p * 2;
    sj[p].nn_guess = (nn_guess && p < (int)nn_guess->offset.size() && nn_guess->offset[p] >= 0) ? nn_guess->slots.p + nn_guess->offset[p] : nullptr;
    mx = std::max(mx, sj[p].ns);
  }
  if (mx == 0) return;
  DBuf<ScoreJob> dsj = to_device(c, sj);
  const float leaf = idx[jobs[0].b].v.leaf;
  const int rv = (int)std::ceil(std::sqrt(std::max(max_range, 0.0)) / (double)leaf) + 1;
  { double b = 0; for (int p = 0; p < P; ++p) b += 16.0 * ((double)sj[p].ns + sj[p].tgt.n); MM_BYTES(c, b); }
  MM_LAUNCH(c, score_kernel, dim3((mx + IB - 1) / IB, P), IB, 0, dsj.p, max_range, rv);
  std::vector<long long> h((size_t)P * 2);
  sums.download(c, h.data(), h.size());
  c.sync();
  for (int p = 0; p < P; ++p)
    if (h[2 * p + 1] > 0) scores[p] = ((double)h[2 * p] / MM3D_FIXD_SCALE) / (double)h[2 * p + 1];
}

}  // namespace mm3d
