// primitives.cu — batched (per-map segmented) radix sort and flag scan.
//
// Both primitives take a list of segments of one concatenated array and run
// every segment in the same launches (blockIdx.y = segment).  The radix sort is
// a stable LSD sort, 8 bits per pass, three kernels per pass (tile histogram,
// per-segment scan, stable scatter).  Stability is what makes the voxel
// centroid summation order canonical (ascending original point index).
#include "mm3d_internal.cuh"

namespace mm3d {

namespace {

constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;  // 4096
constexpr int RS_WARPS = RS_THREADS / 32;

__global__ void __launch_bounds__(RS_THREADS) rs_hist_kernel(const uint32_t* __restrict__ keys, const Seg* __restrict__ segs, int shift,
                                                            uint32_t* __restrict__ hist, int tiles_max)
{
  const Seg sg = segs[blockIdx.y];
  const int tile = blockIdx.x;
  const int base = tile * RS_TILE;
  if (base >= sg.n) return;
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t* k = keys + sg.off;
#pragma unroll
  for (int it = 0; it < RS_ITEMS; ++it) {
    const int i = base + it * RS_THREADS + threadIdx.x;
    if (i < sg.n) atomicAdd(&h[(k[i] >> shift) & 255u], 1u);
  }
  __syncthreads();
  hist[((size_t)blockIdx.y * 256 + threadIdx.x) * tiles_max + tile] = h[threadIdx.x];
}

// exclusive block scan helper (BLOCK threads), returns exclusive prefix, *total = block sum
template <int BLOCK>
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* warp_sums /*[BLOCK/32]*/, uint32_t* total)
{
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_sums[w] = inc;
  __syncthreads();
  if (w == 0) {
    uint32_t s = (lane < BLOCK / 32) ? warp_sums[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += t;
    }
    if (lane < BLOCK / 32) warp_sums[lane] = s;  // inclusive over warps
  }
  __syncthreads();
  const uint32_t warp_off = (w == 0) ? 0 : warp_sums[w - 1];
  *total = warp_sums[BLOCK / 32 - 1];
  const uint32_t r = warp_off + inc - v;
  __syncthreads();
  return r;
}

// Exclusive scan of the tile histogram in (digit-major, tile-minor) order, two levels so that ONE long segment (composeMaps:
// 40 M points = 9 766 tiles) is not scanned by a single block: one block per (segment, digit) scans that digit's tile counts
// in place and leaves the digit's total; a second, tiny kernel turns the 256 totals of a segment into digit bases, which the
// scatter adds.
__global__ void __launch_bounds__(1024) rs_scan_kernel(uint32_t* __restrict__ hist, const Seg* __restrict__ segs, int tiles_max,
                                                      uint32_t* __restrict__ digit_total)
{
  __shared__ uint32_t ws[32];
  const Seg sg = segs[blockIdx.y];
  const int ntiles = (sg.n + RS_TILE - 1) / RS_TILE;
  uint32_t* h = hist + ((size_t)blockIdx.y * 256 + blockIdx.x) * tiles_max;
  uint32_t carry = 0;
  for (int base = 0; base < ntiles; base += 1024) {
    const int t = base + threadIdx.x;
    const uint32_t v = t < ntiles ? h[t] : 0;
    uint32_t tot;
    const uint32_t ex = block_exclusive_scan<1024>(v, ws, &tot);
    if (t < ntiles) h[t] = carry + ex;
    carry += tot;
  }
  if (threadIdx.x == 0) digit_total[blockIdx.y * 256 + blockIdx.x] = carry;
}

__global__ void __launch_bounds__(256) rs_digit_scan_kernel(uint32_t* __restrict__ digit_total)
{
  __shared__ uint32_t ws[8];
  uint32_t tot;
  const uint32_t v = digit_total[blockIdx.x * 256 + threadIdx.x];
  digit_total[blockIdx.x * 256 + threadIdx.x] = block_exclusive_scan<256>(v, ws, &tot);
}

__global__ void __launch_bounds__(RS_THREADS) rs_scatter_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                                               uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                                                               const Seg* __restrict__ segs, int shift, const uint32_t* __restrict__ hist,
                                                               int tiles_max, const uint32_t* __restrict__ digit_base)
{
  const Seg sg = segs[blockIdx.y];
  const int tile = blockIdx.x;
  const int tbase = tile * RS_TILE;
  if (tbase >= sg.n) return;
  __shared__ uint32_t whist[RS_WARPS][256];
  __shared__ uint32_t gbase[256];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&whist[0][0])[i] = 0;
  __syncthreads();
  const uint32_t* kin = keys_in + sg.off;
  const uint32_t* vin = vals_in + sg.off;
  uint32_t key[RS_ITEMS];
  uint32_t rank[RS_ITEMS];
  const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
  for (int it = 0; it < RS_ITEMS; ++it) {
    const int i = tbase + w * (RS_ITEMS * 32) + it * 32 + lane;
    const bool valid = i < sg.n;
    key[it] = valid ? kin[i] : 0xffffffffu;
    const uint32_t d = valid ? ((key[it] >> shift) & 255u) : 256u;
    const uint32_t mask = __match_any_sync(0xffffffffu, d);
    const uint32_t prev = valid ? whist[w][d] : 0u;
    __syncwarp();
    if (valid && lane == (__ffs(mask) - 1)) whist[w][d] = prev + __popc(mask);
    __syncwarp();
    rank[it] = prev + __popc(mask & lt);
  }
  __syncthreads();
  {
    const int d = threadIdx.x;  // RS_THREADS == 256
    uint32_t run = 0;
#pragma unroll
    for (int ww = 0; ww < RS_WARPS; ++ww) {
      const uint32_t t = whist[ww][d];
      whist[ww][d] = run;
      run += t;
    }
    gbase[d] = digit_base[blockIdx.y * 256 + d] + hist[((size_t)blockIdx.y * 256 + d) * tiles_max + tile];
  }
  __syncthreads();
  uint32_t* kout = keys_out + sg.off;
  uint32_t* vout = vals_out + sg.off;
#pragma unroll
  for (int it = 0; it < RS_ITEMS; ++it) {
    const int i = tbase + w * (RS_ITEMS * 32) + it * 32 + lane;
    if (i < sg.n) {
      const uint32_t d = (key[it] >> shift) & 255u;
      const uint32_t pos = gbase[d] + whist[w][d] + rank[it];
      kout[pos] = key[it];
      vout[pos] = vin[i];
    }
  }
}

// ---- flag scan ------------------------------------------------------------
constexpr int SC_THREADS = 256;
constexpr int SC_ITEMS = 8;
constexpr int SC_TILE = SC_THREADS * SC_ITEMS;  // 2048

__global__ void __launch_bounds__(SC_THREADS) sc_tile_sum_kernel(const uint32_t* __restrict__ flags, const Seg* __restrict__ segs,
                                                                uint32_t* __restrict__ tile_sums, int tiles_max)
{
  const Seg sg = segs[blockIdx.y];
  const int base = blockIdx.x * SC_TILE;
  if (base >= sg.n) return;
  __shared__ uint32_t ws[SC_THREADS / 32];
  const uint32_t* f = flags + sg.off;
  uint32_t s = 0;
#pragma unroll
  for (int it = 0; it < SC_ITEMS; ++it) {
    const int i = base + threadIdx.x * SC_ITEMS + it;
    if (i < sg.n) s += f[i];
  }
  uint32_t tot;
  block_exclusive_scan<SC_THREADS>(s, ws, &tot);
  if (threadIdx.x == 0) tile_sums[(size_t)blockIdx.y * tiles_max + blockIdx.x] = tot;
}

__global__ void __launch_bounds__(1024) sc_tile_scan_kernel(uint32_t* __restrict__ tile_sums, const Seg* __restrict__ segs, int tiles_max,
                                                           int* __restrict__ totals)
{
  __shared__ uint32_t ws[32];
  const Seg sg = segs[blockIdx.x];
  const int ntiles = (sg.n + SC_TILE - 1) / SC_TILE;
  uint32_t carry = 0;
  for (int base = 0; base < ntiles; base += 1024) {
    const int e = base + threadIdx.x;
    const uint32_t v = (e < ntiles) ? tile_sums[(size_t)blockIdx.x * tiles_max + e] : 0;
    uint32_t tot;
    const uint32_t ex = block_exclusive_scan<1024>(v, ws, &tot);
    if (e < ntiles) tile_sums[(size_t)blockIdx.x * tiles_max + e] = carry + ex;
    carry += tot;
  }
  if (threadIdx.x == 0) totals[blockIdx.x] = (int)carry;
}

__global__ void __launch_bounds__(SC_THREADS) sc_apply_kernel(const uint32_t* __restrict__ flags, uint32_t* __restrict__ pos,
                                                             const Seg* __restrict__ segs, const uint32_t* __restrict__ tile_sums,
                                                             int tiles_max)
{
  const Seg sg = segs[blockIdx.y];
  const int base = blockIdx.x * SC_TILE;
  if (base >= sg.n) return;
  __shared__ uint32_t ws[SC_THREADS / 32];
  const uint32_t* f = flags + sg.off;
  uint32_t v[SC_ITEMS];
  uint32_t s = 0;
#pragma unroll
  for (int it = 0; it < SC_ITEMS; ++it) {
    const int i = base + threadIdx.x * SC_ITEMS + it;
    v[it] = (i < sg.n) ? f[i] : 0;
    s += v[it];
  }
  uint32_t tot;
  uint32_t ex = block_exclusive_scan<SC_THREADS>(s, ws, &tot) + tile_sums[(size_t)blockIdx.y * tiles_max + blockIdx.x];
  uint32_t* p = pos + sg.off;
#pragma unroll
  for (int it = 0; it < SC_ITEMS; ++it) {
    const int i = base + threadIdx.x * SC_ITEMS + it;
    if (i < sg.n) p[i] = ex;
    ex += v[it];
  }
}

}  // namespace

void radix_sort_pairs_batch(Ctx& c, uint32_t* keys, uint32_t* vals, uint32_t* keys_tmp, uint32_t* vals_tmp, const std::vector<Seg>& segs,
                            int nbits, uint32_t** keys_sorted, uint32_t** vals_sorted)
{
  *keys_sorted = keys;
  *vals_sorted = vals;
  int max_n = 0;
  for (const Seg& s : segs) max_n = std::max(max_n, s.n);
  if (max_n == 0 || nbits <= 0) return;
  const int tiles_max = (max_n + RS_TILE - 1) / RS_TILE;
  DBuf<Seg> dsegs = to_device(c, segs);
  DBuf<uint32_t> hist(c, (size_t)segs.size() * 256 * tiles_max), digit(c, segs.size() * 256);
  const dim3 grid(tiles_max, (unsigned)segs.size());
  uint32_t *kin = keys, *vin = vals, *kout = keys_tmp, *vout = vals_tmp;
  for (int shift = 0; shift < nbits; shift += 8) {
    double tot_n = 0;
    for (const Seg& s : segs) tot_n += s.n;
    MM_BYTES(c, 4.0 * tot_n);
    MM_LAUNCH(c, rs_hist_kernel, grid, RS_THREADS, 0, kin, dsegs.p, shift, hist.p, tiles_max);
    MM_LAUNCH(c, rs_scan_kernel, dim3(256, (unsigned)segs.size()), 1024, 0, hist.p, dsegs.p, tiles_max, digit.p);
    MM_LAUNCH(c, rs_digit_scan_kernel, (unsigned)segs.size(), 256, 0, digit.p);
    MM_BYTES(c, 16.0 * tot_n);
    MM_LAUNCH(c, rs_scatter_kernel, grid, RS_THREADS, 0, kin, vin, kout, vout, dsegs.p, shift, hist.p, tiles_max, digit.p);
    std::swap(kin, kout);
    std::swap(vin, vout);
  }
  *keys_sorted = kin;
  *vals_sorted = vin;
}

void scan_flags_batch(Ctx& c, const uint32_t* flags, uint32_t* pos, const std::vector<Seg>& segs, std::vector<int>& totals)
{
  totals.assign(segs.size(), 0);
  int max_n = 0;
  for (const Seg& s : segs) max_n = std::max(max_n, s.n);
  if (max_n == 0) return;
  const int tiles_max = (max_n + SC_TILE - 1) / SC_TILE;
  DBuf<Seg> dsegs = to_device(c, segs);
  DBuf<uint32_t> tile_sums(c, (size_t)segs.size() * tiles_max);
  DBuf<int> dtot(c, segs.size());
  const dim3 grid(tiles_max, (unsigned)segs.size());
  MM_LAUNCH(c, sc_tile_sum_kernel, grid, SC_THREADS, 0, flags, dsegs.p, tile_sums.p, tiles_max);
  MM_LAUNCH(c, sc_tile_scan_kernel, (unsigned)segs.size(), 1024, 0, tile_sums.p, dsegs.p, tiles_max, dtot.p);
  MM_LAUNCH(c, sc_apply_kernel, grid, SC_THREADS, 0, flags, pos, dsegs.p, tile_sums.p, tiles_max);
  dtot.download(c, totals.data(), segs.size());
  c.sync();
}

}  // namespace mm3d
