// knn_tc.cu — descriptor k-NN on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), exact after re-rank.
//
// findFeatureCorrespondences (map_merge_3d/src/matching.cpp:31-93) needs, for every descriptor of one map, its k nearest
// descriptors of another map under flann::L2_Simple — the one dense contraction on the path.  Squared distances are
// ||a||^2 + ||b||^2 - 2 a.b; the a.b part is a GEMM:
//   * operands are split x = x_hi + x_lo (x_hi = the 10 mantissa bits TF32 keeps, x_lo the exact remainder) and stored per
//     16 dimensions as one 128-byte row [hi(16) | lo(16)]; per 8-dimension slice three kind::tf32 MMAs accumulate
//     a_hi.b_hi + a_lo.b_hi + a_hi.b_lo, i.e. the dot product to ~2^-20 relative instead of TF32's 2^-10;
//   * descriptors are centred on their common mean first (distances are translation invariant, the norms — and with them
//     the absolute error of the expansion — shrink), -2 is folded into A and ||b||^2 rides along as three virtual
//     dimensions, so the accumulator IS the filter value ||b||^2 - 2 a.b;
//   * TMA (cp.async.bulk.tensor, 128-byte swizzle) brings two A tiles (256 query rows) in once — they stay in shared memory
//     when D <= 61 — and streams 128-column k-blocks of B through an mbarrier ring; one elected thread issues tcgen05.mma
//     (M = 128, N = 128, K = 8) into two double-buffered TMEM accumulators, so every B tile read from L2 feeds two MMAs;
//   * eight epilogue warps read the accumulators with tcgen05.ld (one TMEM lane = one query row per thread).  The GEMM runs
//     twice: pass 0 picks, per row, k distinct columns with small filter values, whose EXACT distances bound the row's k-th
//     nearest distance; pass 1 marks every column whose rigorous lower bound is within that bound as one bit in memory;
//   * knn_eval_kernel evaluates the marked columns with the EXACT sequential FP32 distance on the original descriptors in
//     ascending column order — indices and distances are bit-identical to the brute-force scan (and to the CPU checker).
#include <cuda.h>

#include <algorithm>
#include <cstring>

#include "mm3d_internal.cuh"

namespace mm3d {

namespace {

constexpr int TM = 128, TN = 128, TK = 32;  // rows of one MMA operand tile, B rows, floats per k-block (one 128-byte swizzle row = 16 dimensions, hi | lo)
constexpr int A_TILES = 2;                  // A tiles per CTA: 256 query rows share every B tile that comes in from L2
constexpr int CTA_ROWS = A_TILES * TM;
constexpr int KB_BYTES = TM * TK * 4;       // one k-block of one operand tile: 16 KB
constexpr int A_RES_MAX_KB = 4;             // A stays resident in shared memory when it has at most this many k-blocks (D <= 61)
constexpr int KMAXTC = 16;
constexpr int ACC_BUFS = 2;                 // TMEM accumulator buffers of A_TILES x 128 columns (2 x 256 = all 512 columns)
constexpr int EPI_WARPS = 8;                // two per scheduler: warps w and w + 4 share a TMEM lane quarter, one A tile each
constexpr int TC_THREADS = 64 + EPI_WARPS * 32;  // warp 0: TMA producer, 1: MMA issuer + TMEM owner, 2-9: epilogue
constexpr int APAD = 36;                    // floats per row of the padded copy of the original descriptors (D = 33 path)
constexpr int EV_WARPS = 24;                // evaluation kernel: warps per block (one block per SM), one query row at a time each
constexpr int EV_LIST = 1024 + 32;          // evaluation kernel: candidates of 32 mask words + the leftovers of the previous group
constexpr int EV_WORDS = 288;               // evaluation kernel: mask words of a row fetched at once (9 per lane)
constexpr int EV_BLOCKS_X = 16;             // evaluation kernel: blocks per job, each a contiguous range of query rows
__host__ __device__ constexpr size_t ev_smem(int q) { return (size_t)EV_WARPS * (EV_LIST * 2 + EV_WORDS * 4 + (size_t)32 * q * 16); }
// shared memory: [A resident: A_TILES x (3 k-blocks for D = 33, else 4)] [stages: B (+ A_TILES x A when A is streamed)] [barriers]
__host__ __device__ constexpr int tc_a_kb(int dreg, bool a_res) { return a_res ? (dreg > 0 ? 3 : A_RES_MAX_KB) : 0; }
__host__ __device__ constexpr int tc_stages(int dreg, bool a_res) { return a_res ? (dreg > 0 ? 6 : 5) : 3; }
__host__ __device__ constexpr size_t tc_smem(int dreg, bool a_res)
{
  return (size_t)A_TILES * tc_a_kb(dreg, a_res) * KB_BYTES + (size_t)tc_stages(dreg, a_res) * (a_res ? 1 : 1 + A_TILES) * KB_BYTES + 1024 + 256;
}

struct TcJob {
  int a_map, b_map;  // tensor maps: A-form of a_map, B-form of b_map
  int na, nb;
  const float* normA;  // ||a - mean||^2
  const float* origA;  // na x D original descriptors
  const float* origB;
  const float4* padA;  // rows padded to a multiple of 4 floats (D = 33 only), else null
  const float4* padB;
  int k;
  int* idx;     // na x k
  float* dist;  // na x k
  float* audit; // optional na x nb: the raw accumulator (tests: error-bound audit)
  int* pick;    // na x KCAP: per row k distinct columns with small filter values — written by pass 0, read by knn_thr_kernel
  float* thr;   // na: per row, upper bound (accumulator space) of the k-th nearest distance — written by knn_thr_kernel, read by pass 1
  uint32_t* masks;  // candidate bits of pass 1: [row][32-column chunk], 4 chunks per B tile
};

// ---- PTX wrappers ------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity), "r"(20000u)  // suspend-time hint (ns): the waiting thread sleeps in the barrier unit instead of spinning in the issue slots
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1)
{
  asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
               "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar)
{
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
{
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major operand tile, 128-byte swizzle: rows at a 128-byte pitch, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t umma_desc_k_sw128(const void* tile)
{
  uint64_t d = 0;
  d |= (uint64_t)((smem_u32(tile) >> 4) & 0x3fff);  // start address
  d |= (uint64_t)1 << 16;                           // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                 // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                           // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                           // SWIZZLE_128B
  return d;
}
// Asynchronous TMEM -> register load of 32 columns of this thread's lane; the registers are valid only after tmem_ld_wait(r),
// which takes them as in/out operands so that the compiler cannot read them early.
__device__ __forceinline__ void tmem_ld_32x32_issue(uint32_t taddr, uint32_t* r)
{
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait(uint32_t* r)
{
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}

// named barrier for the epilogue warps only (the producer / MMA warps never join it)
template <int THREADS>
__device__ __forceinline__ void named_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory"); }

__device__ __forceinline__ float fmin3(float a, float b, float c)
{
  float r;
  asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));  // one FMNMX3 on sm_100
  return r;
}

// Error model of the filter (validated on the device by tests/test_full_size_gpu.py::test_tensor_core_knn_error_bound, which
// reads the raw accumulators back through mm3d_knn_tc_audit):
//   acc(i, j) = (1 - ES) ||b'_j||^2 - 2 a'_i . b'_j     as the tensor core delivers it (a', b' = descriptors minus the common mean)
//   v(i, j)   = acc(i, j) + (1 - ES) ||a'_i||^2
//   v(i, j) <= d(i, j) <= v(i, j) + 2 ES (||a'_i||^2 + ||b'_j||^2)      d = the sequential FP32 distance of the exact scan
// The true deviation of the dot product has three parts: the dropped lo.lo products and the TF32 truncation of the lo parts
// (|x_lo| <= 2^-11 |x|, truncated to 11 bits: <= (2^-22 + 2 * 2^-21) |a_k||b_k| per dimension, i.e. <= 1.2e-6 ||a'|| ||b'||), the
// FP32 accumulation inside the tensor core (15-18 chained MMAs of 8 products; each partial sum is bounded by ||b'||^2 + 2 ||a'|| ||b'||),
// and the FP32 rounding of the norms and of d itself (<= 35 * 2^-24 relative).  With 2 ||a'|| ||b'|| <= ||a'||^2 + ||b'||^2 the sum
// of the three stays below EG (||a'||^2 + ||b'||^2); ES = EG + margin, so v is a lower bound of d.  Only the LOWER bound is used
// by the filter: the k-th nearest distance is bounded from above by EXACT distances (below).
constexpr float TC_ERR_STORE = 3.1e-5f;  // ES

// Insertion of (d, j) into a list sorted by (distance, arrival): candidates arrive in ascending column order, so a new
// element goes behind every element with an equal distance (strict <), and from its slot on the displaced elements shift
// down UNCONDITIONALLY — comparing them again would let a displaced element lose against an equal neighbour that sat
// behind it (two identical descriptors: the higher column would overtake the lower one).  Result: the brute-force scan's
// (distance, index) order.
template <int KCAP>
__device__ __forceinline__ void topk_insert(float (&bd)[KCAP], int (&bi)[KCAP], float cd, int ci)
{
  bool shifting = false;
#pragma unroll
  for (int u = 0; u < KCAP; ++u) {
    if (shifting || cd < bd[u]) {
      const float td = bd[u];
      const int ti = bi[u];
      bd[u] = cd;
      bi[u] = ci;
      cd = td;
      ci = ti;
      shifting = true;
    }
  }
}

// The exact sequential FP32 distance of the brute-force scan (knn_small_kernel): acc += (a_t - b_t)^2 for t = 0 .. D-1, no
// contraction.  DREG = descriptor length when rows are read as float4 from the padded copies, 0 = scalar reads.
template <int DREG>
__device__ __forceinline__ float exact_dist(const TcJob& job, int D, int row, int col)
{
  float acc = 0.f;
  if constexpr (DREG > 0) {
    constexpr int Q = (DREG + 3) / 4;
    const float4* bp = job.padB + (size_t)col * Q;
    const float4* ap = job.padA + (size_t)row * Q;
    float a_[Q * 4], b_[Q * 4];
#pragma unroll
    for (int t = 0; t < Q; ++t) {
      const float4 v = __ldg(&bp[t]);
      b_[4 * t] = v.x; b_[4 * t + 1] = v.y; b_[4 * t + 2] = v.z; b_[4 * t + 3] = v.w;
      const float4 w = __ldg(&ap[t]);
      a_[4 * t] = w.x; a_[4 * t + 1] = w.y; a_[4 * t + 2] = w.z; a_[4 * t + 3] = w.w;
    }
#pragma unroll
    for (int t = 0; t < DREG; ++t) {
      const float diff = a_[t] - b_[t];
      acc += diff * diff;
    }
  } else {
    const float* a = job.origA + (size_t)row * D;
    const float* b = job.origB + (size_t)col * D;
    for (int t = 0; t < D; ++t) {
      const float diff = __ldg(&a[t]) - __ldg(&b[t]);
      acc += diff * diff;
    }
  }
  return acc;
}

// KCAP = capacity of the top lists (>= k); DREG = descriptor length when rows are read as float4 from the padded copies,
// 0 = scalar reads; A_RES = the A tiles stay in shared memory for the whole CTA (kblocks <= A_RES_MAX_KB); AUDIT = also dump
// the accumulators (tests).  A CTA owns 256 query rows — two 128-row A tiles, so every B tile that comes in from L2 feeds
// two MMAs (with 128 rows per CTA both passes were bound by the L2 -> shared-memory stream of B: 5.5-6.7 TB/s, tensor pipe
// 41-61 % busy, ncu round 2) — and the same GEMM runs twice with two minimal epilogues, one query row per thread:
//   PASS 0 (picks): per 32-column chunk the smallest accumulator, its column carried in the five low mantissa bits (an
//     error of <= 32 ulp on a value that is only used to PICK columns); per row the k chunks with the smallest minima.
//     They are k distinct columns, so the largest of their EXACT distances (knn_thr_kernel) bounds the row's k-th nearest
//     distance from above, with no error term at all: thr = that - (1 - ES)||a'||^2.
//   PASS 1 (candidates): a column can be among the k nearest only if its lower bound acc <= thr.  One bit per (row,
//     column) goes to global memory (zeroed beforehand, only non-zero words are stored); knn_eval_kernel turns the bits
//     into the exact result.
// Why not evaluate inside the epilogue (round 2's first versions): candidates are very unevenly spread over the rows
// (clustered descriptors: a thousand columns within the margin of some rows, five for most), and one CTA per SM with
// 10-16 warps cannot hide the latency of the exact evaluations — that pass ran at 17 % tensor-pipe utilisation.
template <int KCAP, int DREG, bool A_RES, bool AUDIT, int PASS>
__global__ void __launch_bounds__(TC_THREADS, 1) knn_tc_kernel(const TcJob* __restrict__ jobs, int job0, const CUtensorMap* __restrict__ mapsA,
                                                              const CUtensorMap* __restrict__ mapsB, int kblocks, int last_kb_half)
{
  constexpr int STAGES = tc_stages(DREG, A_RES);
  constexpr int STAGE_BYTES = (A_RES ? 1 : 1 + A_TILES) * KB_BYTES;
  constexpr int A_RES_BYTES = A_TILES * tc_a_kb(DREG, A_RES) * KB_BYTES;
  constexpr int ACC_COLS = A_TILES * TN;  // TMEM columns of one accumulator buffer
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by OFFSET, so that every pointer below stays a shared-memory pointer for the compiler
  uint8_t* a_res = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // [A tile][k-block]
  uint8_t* tiles = a_res + A_RES_BYTES;
  uint64_t* full_bar = (uint64_t*)(tiles + (size_t)STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + ACC_BUFS;
  uint64_t* a_full = tmem_empty + ACC_BUFS;
  uint32_t* tmem_slot = (uint32_t*)(a_full + 1);

  const TcJob& job = jobs[job0 + blockIdx.y];
  const int na = job.na, nb = job.nb;
  const int m0 = blockIdx.x * CTA_ROWS;
  if (m0 >= na) return;  // block-uniform
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (nb + TN - 1) / TN;
  const int a_map = job.a_map, b_map = job.b_map;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < ACC_BUFS; ++b) {
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_empty[b], EPI_WARPS);  // one arrival per epilogue warp
    }
    mbar_init(a_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)(ACC_BUFS * ACC_COLS)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====  (rows past the end of a set are zero-filled by the TMA unit)
    if (lane == 0) {
      if (A_RES) {
        mbar_arrive_expect_tx(a_full, (uint32_t)(A_TILES * kblocks) * KB_BYTES);
        for (int at = 0; at < A_TILES; ++at)
          for (int kb = 0; kb < kblocks; ++kb)
            tma_load_2d(&mapsA[a_map], a_full, a_res + (size_t)(at * tc_a_kb(DREG, A_RES) + kb) * KB_BYTES, kb * TK, m0 + at * TM);
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int nt = 0; nt < n_tiles; ++nt)
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
          uint8_t* st = tiles + (size_t)stage * STAGE_BYTES;
          tma_load_2d(&mapsB[b_map], &full_bar[stage], st, kb * TK, nt * TN);
          if (!A_RES)
            for (int at = 0; at < A_TILES; ++at) tma_load_2d(&mapsA[a_map], &full_bar[stage], st + (size_t)(1 + at) * KB_BYTES, kb * TK, m0 + at * TM);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      // kind::tf32, FP32 accumulate, both operands K-major, M = 128, N = 128
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
      if (A_RES) {
        mbar_wait(a_full, 0);
        tc_fence_after();
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int nt = 0; nt < n_tiles; ++nt) {
        const int buf = nt % ACC_BUFS;
        const uint32_t acc_phase = (uint32_t)(nt / ACC_BUFS) & 1u;
        mbar_wait(&tmem_empty[buf], acc_phase ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          uint8_t* st = tiles + (size_t)stage * STAGE_BYTES;
          const uint64_t db = umma_desc_k_sw128(st);
          // a k-block row = [hi(8) hi(8) | lo(8) lo(8)] of 16 dimensions; 8 TF32 = 32 bytes = +2 in the (address >> 4) field.
          // hi.hi + lo.hi + hi.lo per 8-dimension slice: the dot product to ~2^-20 relative.  When the last k-block holds
          // at most 8 dimensions (D = 33: 33 + 3 norm dimensions = 36 = 2 k-blocks + 4) its second slice is all zeros and
          // is skipped: 15 MMAs per A tile and B tile instead of 18.
          constexpr int PA[6] = {0, 2, 0, 1, 3, 1};
          constexpr int PB[6] = {0, 0, 2, 1, 1, 3};
          const int np = (last_kb_half && kb == kblocks - 1) ? 3 : 6;  // the first three products are the first slice's
#pragma unroll
          for (int at = 0; at < A_TILES; ++at) {
            const uint64_t da = umma_desc_k_sw128(A_RES ? a_res + (size_t)(at * tc_a_kb(DREG, A_RES) + kb) * KB_BYTES : st + (size_t)(1 + at) * KB_BYTES);
            const uint32_t tmem_d = tmem_base + (uint32_t)(buf * ACC_COLS + at * TN);
#pragma unroll
            for (int p = 0; p < 6; ++p)
              if (p < np) tc_mma_tf32(tmem_d, da + (uint64_t)(PA[p] * 2), db + (uint64_t)(PB[p] * 2), idesc, (kb | p) != 0 ? 1u : 0u);
          }
          tc_commit(&empty_bar[stage]);  // the smem slot is free once these MMAs have read it
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit(&tmem_full[buf]);  // both accumulators ready for the epilogue
      }
    }
  } else {
    // ===== epilogue: one query row per thread; warps w and w + 4 share a TMEM lane quarter and take one A tile each =====
    const int ew = warp - 2;
    const int q = warp & 3;   // TMEM lane quarter this warp may access
    const int at = ew >> 2;   // which A tile
    const int row = m0 + at * TM + q * 32 + lane;
    const bool live = row < na;
    const float INF = __int_as_float(0x7f800000);
    const uint32_t tmem_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(at * TN);
    // PASS 0 state: the KCAP smallest chunk minima (column in the low 5 bits) and their chunk ids
    float t5[KCAP];
    int c5[KCAP];
#pragma unroll
    for (int i = 0; i < KCAP; ++i) {
      t5[i] = INF;
      c5[i] = -1;
    }
    uint32_t keep = 0xffffffe0u;
    asm volatile("" : "+r"(keep));  // a register, so that (r & keep) | column is ONE LOP3 (two immediates would be two)
    // PASS 1 state
    float thr = -INF;
    uint32_t* mrow = nullptr;
    float* audit_row = nullptr;
    if constexpr (PASS == 1) {
      thr = live ? job.thr[row] : -INF;
      mrow = job.masks + (size_t)row * (size_t)(n_tiles * 4);
      if (AUDIT) audit_row = job.audit + (size_t)row * nb;
    }
    // one chunk: r = the thread's 32 accumulators of columns chunk_id * 32 .. (destroyed); returns the candidate bits (pass 1)
    auto process = [&](uint32_t (&r)[32], int chunk_id) -> uint32_t {
      const int jbase = chunk_id * 32;
      const int left = nb - jbase;  // columns past nb are zero rows of the B form: keep them out
      uint32_t msk = 0;
      if constexpr (PASS == 0) {
        if (left > 0) {
#pragma unroll
          for (int c = 0; c < 32; ++c) r[c] = (r[c] & keep) | (uint32_t)c;
          if (left < 32) {
#pragma unroll
            for (int c = 0; c < 32; ++c)
              if (c >= left) r[c] = 0x7f800000u;
          }
          // minimum as a tree (depth 4 of three-input minima): the chunk is a dependent chain otherwise
          float l1[11];
#pragma unroll
          for (int c = 0; c < 10; ++c) l1[c] = fmin3(__uint_as_float(r[3 * c]), __uint_as_float(r[3 * c + 1]), __uint_as_float(r[3 * c + 2]));
          l1[10] = fminf(__uint_as_float(r[30]), __uint_as_float(r[31]));
          const float l2a = fmin3(l1[0], l1[1], l1[2]), l2b = fmin3(l1[3], l1[4], l1[5]), l2c = fmin3(l1[6], l1[7], l1[8]);
          const float l2d = fminf(l1[9], l1[10]);
          const float m = fminf(fmin3(l2a, l2b, l2c), l2d);
          if (m < t5[KCAP - 1]) {
            float cd = m;
            int ci = chunk_id;
#pragma unroll
            for (int u = 0; u < KCAP; ++u) {
              if (cd < t5[u]) {
                const float td = t5[u];
                const int ti = c5[u];
                t5[u] = cd;
                c5[u] = ci;
                cd = td;
                ci = ti;
              }
            }
          }
        }
      } else {
        if (AUDIT) {
          if (live) {
#pragma unroll
            for (int c = 0; c < 32; ++c)
              if (c < left) audit_row[jbase + c] = __uint_as_float(r[c]);
          }
        }
        if (left > 0) {
          if (left < 32) {
#pragma unroll
            for (int c = 0; c < 32; ++c)
              if (c >= left) r[c] = 0x7f800000u;
          }
          // smallest accumulator of the chunk; the common case — nothing passes — ends here
          float l1[11];
#pragma unroll
          for (int c = 0; c < 10; ++c) l1[c] = fmin3(__uint_as_float(r[3 * c]), __uint_as_float(r[3 * c + 1]), __uint_as_float(r[3 * c + 2]));
          l1[10] = fminf(__uint_as_float(r[30]), __uint_as_float(r[31]));
          const float l2a = fmin3(l1[0], l1[1], l1[2]), l2b = fmin3(l1[3], l1[4], l1[5]), l2c = fmin3(l1[6], l1[7], l1[8]);
          const float l2d = fminf(l1[9], l1[10]);
          const float m = fminf(fmin3(l2a, l2b, l2c), l2d);
          if (m <= thr) {
            uint32_t m0_ = 0, m1_ = 0, m2_ = 0, m3_ = 0;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              m0_ |= (__uint_as_float(r[c]) <= thr) ? (1u << c) : 0u;
              m1_ |= (__uint_as_float(r[8 + c]) <= thr) ? (1u << (8 + c)) : 0u;
              m2_ |= (__uint_as_float(r[16 + c]) <= thr) ? (1u << (16 + c)) : 0u;
              m3_ |= (__uint_as_float(r[24 + c]) <= thr) ? (1u << (24 + c)) : 0u;
            }
            msk = (m0_ | m1_) | (m2_ | m3_);
            if (left < 32) msk &= (1u << left) - 1u;  // columns past nb (masked to +inf above) would still pass an infinite bound
          }
        }
      }
      return msk;
    };
    // Four chunks (the 128 columns of this warp's accumulator) per tile, two register arrays in turn: while one chunk is being
    // processed the TMEM load of the next is in flight, and nothing is copied between the arrays.
    uint32_t r0[32], r1[32];
    mbar_wait(&tmem_full[0], 0);
    tc_fence_after();
    tmem_ld_32x32_issue(tmem_row, r0);
    tmem_ld_wait(r0);
    for (int nt = 0; nt < n_tiles; ++nt) {
      const int buf = nt % ACC_BUFS;
      const uint32_t trow = tmem_row + (uint32_t)(buf * ACC_COLS);
      tmem_ld_32x32_issue(trow + 32, r1);
      const uint32_t k0 = process(r0, nt * 4);
      tmem_ld_wait(r1);
      tmem_ld_32x32_issue(trow + 64, r0);
      const uint32_t k1 = process(r1, nt * 4 + 1);
      tmem_ld_wait(r0);
      tmem_ld_32x32_issue(trow + 96, r1);
      const uint32_t k2 = process(r0, nt * 4 + 2);
      tmem_ld_wait(r1);  // every read of this tile's accumulator is done
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[buf]);
      if (nt + 1 < n_tiles) {
        const int buf1 = (nt + 1) % ACC_BUFS;
        mbar_wait(&tmem_full[buf1], (uint32_t)((nt + 1) / ACC_BUFS) & 1u);
        tc_fence_after();
        tmem_ld_32x32_issue(tmem_row + (uint32_t)(buf1 * ACC_COLS), r0);
      }
      const uint32_t k3 = process(r1, nt * 4 + 3);
      if constexpr (PASS == 1) {
        if ((k0 | k1) | (k2 | k3)) *reinterpret_cast<uint4*>(mrow + nt * 4) = make_uint4(k0, k1, k2, k3);  // the buffer is zeroed beforehand
      }
      if (nt + 1 < n_tiles) tmem_ld_wait(r0);
    }
    if constexpr (PASS == 0) {
      if (live) {
        int* pk = job.pick + (size_t)row * KCAP;
#pragma unroll
        for (int t = 0; t < KCAP; ++t) pk[t] = c5[t] >= 0 ? c5[t] * 32 + (int)(__float_as_uint(t5[t]) & 31u) : -1;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(ACC_BUFS * ACC_COLS)) : "memory");
  }
}

// Upper bound of every row's k-th nearest distance from the EXACT distances of the k distinct columns pass 0 picked; in
// accumulator space: a column j is needed only if d_j <= dmax, and acc_j + (1 - ES)||a'||^2 <= d_j.
template <int KCAP, int DREG>
__global__ void __launch_bounds__(128) knn_thr_kernel(const TcJob* __restrict__ jobs, int job0, int D)
{
  const TcJob& job = jobs[job0 + blockIdx.y];
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= job.na) return;
  const float INF = __int_as_float(0x7f800000);
  const int k = min(job.k, KCAP);
  const int* pk = job.pick + (size_t)row * KCAP;
  float dmax = 0.f;
  bool all = true;
  for (int t = 0; t < k; ++t) {
    const int col = pk[t];
    if (col < 0) {  // fewer than k columns in B: everything is a candidate
      all = false;
      break;
    }
    dmax = fmaxf(dmax, exact_dist<DREG>(job, D, row, col));
  }
  const float na_low = (1.0f - TC_ERR_STORE) * job.normA[row];
  job.thr[row] = all ? (dmax - na_low) + 1e-6f * (1.0f + fabsf(dmax) + na_low) : INF;  // + rounding pad of the two FP32 operations
}

// Exact evaluation of the candidates pass 1 marked.  One warp per query row at a time: the row's mask words are fetched
// (lane = 32-column chunk), the set bits expanded — column index only, ascending — into the warp's list in shared memory,
// and every 32 candidates are evaluated by the warp together, lane c = candidate c: the exact sequential FP32 distance on
// the original descriptors, the few that beat the row's current k-th distance inserted in list order = ascending column
// order with strict <, which is the brute-force scan's (distance, index) order bit for bit.  The query row stays in
// registers; the 32 B rows of a step are fetched by the warp TOGETHER — consecutive lanes read consecutive 16-byte pieces,
// so one load instruction touches ~4 rows = ~5 cache lines instead of 32 — and every lane picks its own row up from shared
// memory.  A block is 24 warps = the whole SM, working on 24 CONSECUTIVE query rows at any time: consecutive keypoints have
// similar descriptors, hence nearly the same candidate columns (clustered descriptors: a thousand columns inside the error
// margin of the k-th distance for ~15 % of the rows, five for the median row), and all warps sweep them in ascending column
// order — the B rows one warp pulls in from L2 are L1 hits for the others (the gather ran at the L2 sector rate before).
template <int KCAP, int DREG>
__global__ void __launch_bounds__(EV_WARPS * 32, 1) knn_eval_kernel(const TcJob* __restrict__ jobs, int job0, int D,
                                                                   unsigned long long* __restrict__ stats)
{
  constexpr int Q = DREG > 0 ? (DREG + 3) / 4 : 1;  // float4 pieces per padded row
  extern __shared__ float4 ev_smem_raw[];
  float4* stage_all = ev_smem_raw;                                                  // [EV_WARPS][32 * Q]: the B rows of one step
  uint32_t* words_all = (uint32_t*)(stage_all + (size_t)EV_WARPS * 32 * Q);         // [EV_WARPS][EV_WORDS]
  unsigned short* list_all = (unsigned short*)(words_all + (size_t)EV_WARPS * EV_WORDS);  // [EV_WARPS][EV_LIST]
  const TcJob& job = jobs[job0 + blockIdx.y];
  const int na = job.na, nb = job.nb, k = job.k;
  const uint32_t* masks = job.masks;
  const float4* padA = job.padA;
  const float4* padB = job.padB;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_chunks = ((nb + TN - 1) / TN) * 4;
  const float INF = __int_as_float(0x7f800000);
  const unsigned FULL = 0xffffffffu;
  unsigned short* lst = list_all + (size_t)warp * EV_LIST;
  uint32_t* wbuf = words_all + (size_t)warp * EV_WORDS;
  float4* stage = stage_all + (size_t)warp * 32 * Q;
  // cooperative fetch geometry: piece f = i * 32 + lane of the 32 x Q pieces of a step belongs to candidate f / Q
  int gcand[Q], gpiece[Q];
#pragma unroll
  for (int i = 0; i < Q; ++i) {
    const int f = i * 32 + lane;
    gcand[i] = f / Q;
    gpiece[i] = f - gcand[i] * Q;
  }
  unsigned evals = 0, flushes = 0, rows_done = 0;
  const int rows_per_block = (na + (int)gridDim.x - 1) / (int)gridDim.x;
  const int row_end = min(na, ((int)blockIdx.x + 1) * rows_per_block);
  for (int row = (int)blockIdx.x * rows_per_block + warp; row < row_end; row += EV_WARPS) {
    float a_[Q * 4];
    if constexpr (DREG > 0) {
#pragma unroll
      for (int t = 0; t < Q; ++t) {
        const float4 w = __ldg(&padA[(size_t)row * Q + t]);
        a_[4 * t] = w.x; a_[4 * t + 1] = w.y; a_[4 * t + 2] = w.z; a_[4 * t + 3] = w.w;
      }
    }
    float bd[KCAP];
    int bi[KCAP];
#pragma unroll
    for (int t = 0; t < KCAP; ++t) {
      bd[t] = INF;
      bi[t] = -1;
    }
    // n candidates at lst[base ..]: every lane keeps the same copy of the row's list, so nothing is broadcast afterwards
    auto step = [&](int base, int n) {
      const int jcol = lane < n ? (int)lst[base + lane] : 0;
      float d = INF;
      if constexpr (DREG > 0) {
        constexpr int H = (Q + 1) / 2;  // two rounds: fewer registers in flight
#pragma unroll
        for (int h0 = 0; h0 < Q; h0 += H) {
          float4 v[H];
#pragma unroll
          for (int i = 0; i < H; ++i) {
            v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (h0 + i < Q && gcand[h0 + i] < n) v[i] = __ldg(&padB[(size_t)lst[base + gcand[h0 + i]] * Q + gpiece[h0 + i]]);
          }
#pragma unroll
          for (int i = 0; i < H; ++i)
            if (h0 + i < Q) stage[(h0 + i) * 32 + lane] = v[i];
        }
        __syncwarp();
        if (lane < n) {
          float acc = 0.f;
#pragma unroll
          for (int t = 0; t < Q; ++t) {
            const float4 bv = stage[lane * Q + t];
            const float b_[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (4 * t + u < DREG) {
                const float diff = a_[4 * t + u] - b_[u];
                acc += diff * diff;
              }
          }
          d = acc;
        }
      } else {
        if (lane < n) d = exact_dist<DREG>(job, D, row, jcol);
      }
      unsigned better = __ballot_sync(FULL, d < bd[KCAP - 1]);
      while (better) {
        const int b = __ffs(better) - 1;
        const float dv = __shfl_sync(FULL, d, b);
        const int jv = __shfl_sync(FULL, jcol, b);
        topk_insert<KCAP>(bd, bi, dv, jv);
        better &= __ballot_sync(FULL, d < bd[KCAP - 1]) & ~((2u << b) - 1u);
      }
      evals += (unsigned)n;
      ++flushes;
      __syncwarp();  // the stage is free again
    };
    int pending = 0;  // candidates waiting at lst[0 .. pending)
    const uint32_t* mp = masks + (size_t)row * n_chunks;
    for (int g0 = 0; g0 < n_chunks; g0 += EV_WORDS) {
      // all mask words of this stretch at once (independent loads), then through shared memory group by group
      {
        uint32_t wv[EV_WORDS / 32];
#pragma unroll
        for (int u = 0; u < EV_WORDS / 32; ++u) {
          const int ch = g0 + u * 32 + lane;
          wv[u] = ch < n_chunks ? __ldcs(&mp[ch]) : 0u;
        }
#pragma unroll
        for (int u = 0; u < EV_WORDS / 32; ++u) wbuf[u * 32 + lane] = wv[u];
      }
      for (int g = 0; g < EV_WORDS && g0 + g < n_chunks; g += 32) {
        uint32_t w = wbuf[g + lane];  // written by this lane: no synchronisation needed
        if (!__any_sync(FULL, w != 0u)) continue;
        const int cnt = __popc(w);
        int pre = cnt;  // inclusive prefix sum over the lanes
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(FULL, pre, o);
          if (lane >= o) pre += t;
        }
        const int total = __shfl_sync(FULL, pre, 31);
        int pos = pending + pre - cnt;
        const int jb = (g0 + g + lane) * 32;
        while (w) {
          const int c = __ffs(w) - 1;
          w &= w - 1;
          lst[pos++] = (unsigned short)(jb + c);
        }
        __syncwarp();
        pending += total;
        int done = 0;
        while (pending - done >= 32) {
          step(done, 32);
          done += 32;
        }
        if (done > 0) {  // the leftovers move to the front
          const int rest = pending - done;
          const unsigned short v = lane < rest ? lst[done + lane] : (unsigned short)0;
          __syncwarp();
          if (lane < rest) lst[lane] = v;
          __syncwarp();
          pending = rest;
        }
      }
    }
    if (pending > 0) step(0, pending);
    if (lane < k) {
      float dv = 0.f;
      int jv = -1;
#pragma unroll
      for (int u = 0; u < KCAP; ++u)
        if (u == lane && bi[u] >= 0) {
          dv = bd[u];
          jv = bi[u];
        }
      job.idx[(size_t)row * k + lane] = jv;
      job.dist[(size_t)row * k + lane] = dv;
    }
    ++rows_done;
  }
  if (stats && lane == 0 && rows_done) {
    atomicAdd(&stats[0], (unsigned long long)rows_done);
    atomicAdd(&stats[1], (unsigned long long)flushes);
    atomicAdd(&stats[2], (unsigned long long)evals);
  }
}

// Per-dimension mean of all descriptors (two deterministic levels): distances are translation invariant, and centring
// shrinks ||a||^2 + ||b||^2 — hence the absolute error of the norm expansion — by the clustering of the descriptors.
struct PrepJob {
  const float* src;  // n x D
  float* formA;      // n x Kp
  float* formB;
  float* norm;
  float* pad;  // n x PADW (D = 33) or null
  int n;
};
constexpr int MEAN_CHUNKS = 32;
__global__ void __launch_bounds__(128) knn_tc_mean_partial_kernel(const PrepJob* __restrict__ jobs, int D, double* __restrict__ partial)
{
  const PrepJob& j = jobs[blockIdx.y];
  const int per = (j.n + MEAN_CHUNKS - 1) / MEAN_CHUNKS;
  const int r0 = blockIdx.x * per, r1 = min(j.n, r0 + per);
  for (int t = threadIdx.x; t < D; t += blockDim.x) {
    double s = 0.0;
    for (int r = r0; r < r1; ++r) s += (double)j.src[(size_t)r * D + t];
    partial[((size_t)blockIdx.y * MEAN_CHUNKS + blockIdx.x) * D + t] = s;
  }
}
__global__ void __launch_bounds__(128) knn_tc_mean_kernel(const double* __restrict__ partial, int n_partials, int D, double inv_rows,
                                                          float* __restrict__ mean)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= D) return;
  double s = 0.0;
  for (int p = 0; p < n_partials; ++p) s += partial[(size_t)p * D + t];
  mean[t] = (float)(s * inv_rows);
}

__device__ __forceinline__ float tf32_hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xffffe000u); }  // the bits TF32 keeps

// Rows of Kp floats in k-blocks of 32 = 16 dimensions: [hi(16) | lo(16)].  A form: -2 (a - mean) split hi/lo;
// B form: (b - mean) split hi/lo.  Three virtual dimensions after D carry the norm: A = 1, B = ||b'||^2 in three TF32 pieces.
__global__ void __launch_bounds__(128) knn_tc_prep_kernel(const PrepJob* __restrict__ jobs, const float* __restrict__ mean, int D, int Kp,
                                                          int padw)
{
  const PrepJob& j = jobs[blockIdx.y];
  const int row = blockIdx.x;
  if (row >= j.n) return;
  const float* a = j.src + (size_t)row * D;
  float* fa = j.formA + (size_t)row * Kp;
  float* fb = j.formB + (size_t)row * Kp;
  __shared__ float s_norm;
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int t = 0; t < D; ++t) {
      const float v = a[t] - mean[t];
      s += v * v;
    }
    j.norm[row] = s;
    s_norm = s;
  }
  __syncthreads();
  const float nrm = (1.0f - TC_ERR_STORE) * s_norm;  // the B form carries the lower-bound norm (see knn_tc_kernel)
  const float nh = tf32_hi(nrm), r1 = nrm - nh, nm = tf32_hi(r1), nl = r1 - nm;
  for (int t = threadIdx.x; t < Kp; t += blockDim.x) {
    const int kb = t >> 5, w = t & 31, is_lo = w >> 4, dim = kb * 16 + (w & 15);
    float va = 0.f, vb = 0.f;
    if (dim < D) {
      const float v = a[dim] - mean[dim];
      const float hi = tf32_hi(v);
      const float part = is_lo ? v - hi /* exact */ : hi;
      va = -2.0f * part;
      vb = part;
    } else if (dim < D + 3 && !is_lo) {
      va = 1.0f;
      vb = dim == D ? nh : (dim == D + 1 ? nm : nl);
    }
    fa[t] = va;
    fb[t] = vb;
  }
  if (j.pad)
    for (int t = threadIdx.x; t < padw; t += blockDim.x) j.pad[(size_t)row * padw + t] = t < D ? a[t] : 0.f;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled()
{
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    MM_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    if (!p || q != cudaDriverEntryPointSuccess) throw CudaError("CUDA error: cuTensorMapEncodeTiled is not available");
    fn = (EncodeTiledFn)p;
  }
  return fn;
}

CUtensorMap make_map(float* base, int rows, int Kp)
{
  CUtensorMap m;
  memset(&m, 0, sizeof(m));
  const cuuint64_t dims[2] = {(cuuint64_t)Kp, (cuuint64_t)std::max(rows, 1)};
  const cuuint64_t strides[1] = {(cuuint64_t)Kp * 4};
  const cuuint32_t box[2] = {(cuuint32_t)TK, (cuuint32_t)TM};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = encode_tiled()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw CudaError("CUDA error: cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
  return m;
}

}  // namespace

// Same contract as the brute-force kernels: idx / dist hold, per row of A, the k nearest rows of B sorted by (distance, index).
// audit (tests only, one problem): receives the raw accumulators (na x nb), the centred norms of both sides and ES.
void knn_tc_batch(Ctx& c, const std::vector<const float*>& desc, const std::vector<int>& n_rows, int D, const std::vector<KnnProblem>& probs,
                  KnnAudit* audit)
{
  if (probs.empty()) return;
  const int M = (int)desc.size();
  const int kblocks = (D + 3 + 15) / 16;
  const int last_kb_half = ((D + 3) - 16 * (kblocks - 1)) <= 8 ? 1 : 0;
  const int Kp = kblocks * TK;
  const bool reg_path = D == 33;
  const int padw = reg_path ? APAD : 0;
  std::vector<char> used(M, 0);
  for (const KnnProblem& p : probs) { used[p.a] = 1; used[p.b] = 1; }
  std::vector<DBuf<float>> formA(M), formB(M), norm(M), pad(M);
  std::vector<PrepJob> pj;
  int mxn = 0;
  for (int m = 0; m < M; ++m) {
    if (!used[m] || n_rows[m] == 0) continue;
    formA[m].alloc(c, (size_t)n_rows[m] * Kp);
    formB[m].alloc(c, (size_t)n_rows[m] * Kp);
    norm[m].alloc(c, n_rows[m]);
    if (reg_path) pad[m].alloc(c, (size_t)n_rows[m] * padw);
    pj.push_back(PrepJob{desc[m], formA[m].p, formB[m].p, norm[m].p, reg_path ? pad[m].p : nullptr, n_rows[m]});
    mxn = std::max(mxn, n_rows[m]);
  }
  if (pj.empty()) return;
  DBuf<PrepJob> dpj = to_device(c, pj);
  size_t rows_all = 0;
  for (const PrepJob& j : pj) rows_all += (size_t)j.n;
  DBuf<double> partial(c, pj.size() * MEAN_CHUNKS * (size_t)D);
  DBuf<float> mean(c, D);
  MM_LAUNCH(c, knn_tc_mean_partial_kernel, dim3(MEAN_CHUNKS, (unsigned)pj.size()), 128, 0, dpj.p, D, partial.p);
  MM_LAUNCH(c, knn_tc_mean_kernel, (D + 127) / 128, 128, 0, partial.p, (int)pj.size() * MEAN_CHUNKS, D, 1.0 / (double)rows_all, mean.p);
  MM_LAUNCH(c, knn_tc_prep_kernel, dim3(mxn, (unsigned)pj.size()), 128, 0, dpj.p, mean.p, D, Kp, padw);
  std::vector<CUtensorMap> hA(M), hB(M);
  for (int m = 0; m < M; ++m) {
    if (!used[m] || n_rows[m] == 0) {
      memset(&hA[m], 0, sizeof(CUtensorMap));
      memset(&hB[m], 0, sizeof(CUtensorMap));
      continue;
    }
    hA[m] = make_map(formA[m].p, n_rows[m], Kp);
    hB[m] = make_map(formB[m].p, n_rows[m], Kp);
  }
  DBuf<CUtensorMap> dA = to_device(c, hA), dB = to_device(c, hB);
  // Jobs go through the kernels in batches whose candidate bits fit a fixed buffer (1 bit per distance-matrix entry:
  // 9.9 MB for 8 800 x 8 800 descriptors, 9.8 GB for the 992 directed problems of config 3 at once).
  constexpr size_t MASK_WORDS_MAX = (size_t)3 << 27;  // 1.5 GiB
  struct Batch {
    int first, count, max_na;
    size_t words;
    double bytes;
  };
  std::vector<Batch> batches;
  std::vector<TcJob> tj;
  int kmax = 0;
  for (const KnnProblem& p : probs) kmax = std::max(kmax, p.k);
  const int kcap = kmax <= 5 ? 5 : (reg_path && kmax <= 10 ? 10 : KMAXTC);  // the KCAP of the kernel variant chosen below
  DBuf<float> audit_acc;
  size_t rows_total = 0;
  for (const KnnProblem& p : probs)
    if (p.na > 0 && n_rows[p.b] > 0) rows_total += (size_t)p.na;
  DBuf<float> thr(c, rows_total);
  DBuf<int> pick(c, rows_total * (size_t)kcap);
  size_t row_off = 0, words_max = 0;
  std::vector<size_t> mask_off;
  for (const KnnProblem& p : probs) {
    if (p.na == 0 || n_rows[p.b] == 0) continue;
    TcJob j;
    j.thr = thr.p + row_off;
    j.pick = pick.p + row_off * (size_t)kcap;
    row_off += (size_t)p.na;
    j.a_map = p.a;
    j.b_map = p.b;
    j.na = p.na;
    j.nb = n_rows[p.b];
    j.normA = norm[p.a].p;
    j.origA = desc[p.a];
    j.origB = desc[p.b];
    j.padA = reg_path ? (const float4*)pad[p.a].p : nullptr;
    j.padB = reg_path ? (const float4*)pad[p.b].p : nullptr;
    j.k = p.k;
    j.idx = p.idx;
    j.dist = p.dist;
    j.audit = nullptr;
    j.masks = nullptr;
    const size_t words = (size_t)j.na * (size_t)((j.nb + TN - 1) / TN * 4);  // a multiple of 4: every job's bits start 16-byte aligned
    if (batches.empty() || batches.back().words + words > MASK_WORDS_MAX || batches.back().count >= 65535)  // gridDim.y limit
      batches.push_back(Batch{(int)tj.size(), 0, 0, 0, 0.0});
    Batch& b = batches.back();
    mask_off.push_back(b.words);
    b.words += words;
    b.count += 1;
    b.max_na = std::max(b.max_na, j.na);
    b.bytes += 4.0 * D * ((double)j.na + j.nb) + 8.0 * j.k * j.na;
    words_max = std::max(words_max, b.words);
    tj.push_back(j);
  }
  if (tj.empty()) return;
  DBuf<uint32_t> masks(c, words_max);
  for (size_t t = 0; t < tj.size(); ++t) tj[t].masks = masks.p + mask_off[t];
  const bool do_audit = audit && tj.size() == 1 && reg_path && kmax <= 5;
  if (do_audit) {
    audit_acc.alloc(c, (size_t)tj[0].na * tj[0].nb);
    tj[0].audit = audit_acc.p;
  }
  DBuf<TcJob> dtj = to_device(c, tj);
  if (!c.knn_stats) {
    MM_CUDA(cudaMalloc((void**)&c.knn_stats, 3 * sizeof(unsigned long long)));
    MM_CUDA(cudaMemsetAsync(c.knn_stats, 0, 3 * sizeof(unsigned long long), c.stream));
  }
#define MM_TC(KCAP, DREG, RES, AUD)                                                                                                    \
  do {                                                                                                                                 \
    const size_t smem = tc_smem(DREG, RES);                                                                                            \
    /* per device and cheap: set on every call (a process may drive several GPUs) */                                                  \
    MM_CUDA(cudaFuncSetAttribute(knn_tc_kernel<KCAP, DREG, RES, false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
    MM_CUDA(cudaFuncSetAttribute(knn_tc_kernel<KCAP, DREG, RES, AUD, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
    const size_t esmem = ev_smem(DREG > 0 ? (DREG + 3) / 4 : 1);                                                                       \
    MM_CUDA(cudaFuncSetAttribute(knn_eval_kernel<KCAP, DREG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)esmem));               \
    for (const Batch& b : batches) {                                                                                                   \
      const dim3 grid((b.max_na + CTA_ROWS - 1) / CTA_ROWS, (unsigned)b.count);                                                        \
      MM_CUDA(cudaMemsetAsync(masks.p, 0, b.words * sizeof(uint32_t), c.stream));                                                      \
      MM_BYTES(c, b.bytes);                                                                                                            \
      MM_LAUNCH(c, (knn_tc_kernel<KCAP, DREG, RES, false, 0>), grid, TC_THREADS, smem, dtj.p, b.first, dA.p, dB.p, kblocks,            \
                last_kb_half);                                                                                                         \
      MM_LAUNCH(c, (knn_thr_kernel<KCAP, DREG>), dim3((b.max_na + 127) / 128, (unsigned)b.count), 128, 0, dtj.p, b.first, D);          \
      MM_BYTES(c, b.bytes);                                                                                                            \
      MM_LAUNCH(c, (knn_tc_kernel<KCAP, DREG, RES, AUD, 1>), grid, TC_THREADS, smem, dtj.p, b.first, dA.p, dB.p, kblocks,              \
                last_kb_half);                                                                                                         \
      MM_BYTES(c, b.bytes);                                                                                                            \
      MM_LAUNCH(c, (knn_eval_kernel<KCAP, DREG>), dim3(EV_BLOCKS_X, (unsigned)b.count), EV_WARPS * 32, esmem, dtj.p, b.first, D,       \
                c.knn_stats);                                                                                                          \
    }                                                                                                                                  \
  } while (0)
  const bool res = kblocks <= A_RES_MAX_KB;
  if (do_audit) MM_TC(5, 33, true, true);
  else if (reg_path && kmax <= 5) MM_TC(5, 33, true, false);
  else if (reg_path && kmax <= 10) MM_TC(10, 33, true, false);
  else if (reg_path) MM_TC(KMAXTC, 33, true, false);
  else if (res && kmax <= 5) MM_TC(5, 0, true, false);
  else if (res) MM_TC(KMAXTC, 0, true, false);
  else if (kmax <= 5) MM_TC(5, 0, false, false);
  else MM_TC(KMAXTC, 0, false, false);
#undef MM_TC
  if (do_audit) {
    audit->acc.resize((size_t)tj[0].na * tj[0].nb);
    audit->norm_a.resize(tj[0].na);
    audit->norm_b.resize(tj[0].nb);
    audit_acc.download(c, audit->acc.data(), audit->acc.size());
    norm[probs[0].a].download(c, audit->norm_a.data(), audit->norm_a.size());
    norm[probs[0].b].download(c, audit->norm_b.data(), audit->norm_b.size());
    audit->err_store = TC_ERR_STORE;
    c.sync();
  }
}

}  // namespace mm3d
