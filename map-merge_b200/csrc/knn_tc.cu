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
//   * TMA (cp.async.bulk.tensor, 128-byte swizzle) brings the A tile in once (it stays in shared memory when D <= 61) and
//     streams 128-column k-blocks of B through an mbarrier ring; one elected thread issues tcgen05.mma (M = 128, N = 128,
//     K = 8) into a double-buffered TMEM accumulator;
//   * four epilogue warps read the accumulator with tcgen05.ld (one TMEM lane = one query row per thread) and keep, per
//     row, a short shared-memory list of the columns that can still be among the k nearest under a rigorous error bound;
//   * the survivors are evaluated with the EXACT sequential FP32 distance on the original descriptors in ascending column
//     order — indices and distances are bit-identical to the brute-force scan (and to the CPU checker).
#include <cuda.h>

#include <algorithm>
#include <cstring>

#include "mm3d_internal.cuh"

namespace mm3d {

namespace {

constexpr int TM = 128, TN = 128, TK = 32;  // A rows, B rows, floats per k-block (one 128-byte swizzle row = 16 dimensions, hi | lo)
constexpr int KB_BYTES = TM * TK * 4;       // one k-block of one operand: 16 KB
constexpr int A_RES_MAX_KB = 4;             // A stays resident in shared memory when it has at most this many k-blocks (D <= 61)
constexpr int LIST_CAP = 64;                // candidate list entries per query row
constexpr int KMAXTC = 16;
constexpr int TC_THREADS = 192;  // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-5: epilogue
constexpr int LIST_BYTES = TM * LIST_CAP * 8;
// A resident: A (<= 64 KB) + 6 B stages; otherwise 4 stages of A + B
constexpr int STAGES_RES = 6, STAGES_STREAM = 4;
constexpr size_t TC_SMEM_RES = (size_t)A_RES_MAX_KB * KB_BYTES + (size_t)STAGES_RES * KB_BYTES + LIST_BYTES + 1024 + 256;
constexpr size_t TC_SMEM_STREAM = (size_t)STAGES_STREAM * 2 * KB_BYTES + LIST_BYTES + 1024 + 256;

struct TcJob {
  int a_map, b_map;  // tensor maps: A-form of a_map, B-form of b_map
  int na, nb;
  const float* normA;  // ||a - mean||^2
  const float* origA;  // na x D original descriptors
  const float* origB;
  const float4* padA;  // rows padded to a multiple of 4 floats (D = 33 only), else null
  const float4* padB;
  const float* normB;  // ||b - mean||^2
  int k;
  int* dense_rows;            // rows of this job that are handed to the exact scan (D = 33 path), capacity na
  unsigned int* dense_count;  // their number
  int* idx;     // na x k
  float* dist;  // na x k
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
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
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

// exact flann::L2_Simple distance on the original descriptors (sequential FP32, the order the CPU scan uses)
template <int DREG>
__device__ __forceinline__ float exact_dist(const float* __restrict__ a_reg, const float* __restrict__ a, const TcJob& job, int jcol, int D)
{
  float acc = 0.f;
  if (DREG > 0) {
    constexpr int Q = DREG > 0 ? (DREG + 3) / 4 : 1;
    const float4* bp = job.padB + (size_t)jcol * Q;
    float b[Q * 4];
#pragma unroll
    for (int t = 0; t < Q; ++t) {
      const float4 v = __ldg(&bp[t]);
      b[4 * t] = v.x; b[4 * t + 1] = v.y; b[4 * t + 2] = v.z; b[4 * t + 3] = v.w;
    }
#pragma unroll
    for (int t = 0; t < DREG; ++t) {
      const float diff = a_reg[t] - b[t];
      acc += diff * diff;
    }
  } else {
    const float* b = job.origB + (size_t)jcol * D;
    for (int t = 0; t < D; ++t) {
      const float diff = a[t] - __ldg(&b[t]);
      acc += diff * diff;
    }
  }
  return acc;
}

// One-sided bound on |approximate - exact| squared distance, relative to ||a'||^2 + ||b'||^2 of the centred descriptors:
// dropped a_lo.b_lo and TF32 truncation of the lo parts (3 * 2^-20), FP32 accumulation inside the tensor core, FP32 norms,
// centring; derived in DESIGN.md §3 (K9) and padded.
constexpr float TC_ERR = 3.0e-5f;
// The B form carries (1 - TC_ERR_STORE) ||b'||^2, so the accumulator is a LOWER bound of the exact distance minus
// (1 - TC_ERR_STORE) ||a'||^2; TC_ERR_STORE exceeds TC_ERR by more than the rounding of that product.
constexpr float TC_ERR_STORE = 3.1e-5f;

// KCAP = capacity of the register top lists (>= k; the first k are written out); DREG = descriptor length when the query
// row is cached in registers (and rows are read as float4 from the padded copies), 0 = scalar reads from global memory;
// A_RES = the A tile stays in shared memory for the whole CTA (kblocks <= A_RES_MAX_KB).
template <int KCAP, int DREG, bool A_RES>
__global__ void __launch_bounds__(TC_THREADS, 1) knn_tc_kernel(const TcJob* __restrict__ jobs, const CUtensorMap* __restrict__ mapsA,
                                                              const CUtensorMap* __restrict__ mapsB, int kblocks, int D,
                                                              unsigned long long* __restrict__ stats)
{
  constexpr int STAGES = A_RES ? STAGES_RES : STAGES_STREAM;
  constexpr int STAGE_BYTES = A_RES ? KB_BYTES : 2 * KB_BYTES;
  constexpr int A_RES_BYTES = A_RES ? A_RES_MAX_KB * KB_BYTES : 0;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* a_res = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* tiles = a_res + A_RES_BYTES;
  float* list_v = (float*)(tiles + (size_t)STAGES * STAGE_BYTES);  // [LIST_CAP][TM]
  int* list_j = (int*)(list_v + TM * LIST_CAP);
  uint64_t* full_bar = (uint64_t*)(list_j + TM * LIST_CAP);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* a_full = tmem_empty + 2;
  uint32_t* tmem_slot = (uint32_t*)(a_full + 1);

  const TcJob job = jobs[blockIdx.y];
  const int m0 = blockIdx.x * TM;
  if (m0 >= job.na) return;  // block-uniform
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (job.nb + TN - 1) / TN;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_empty[b], 4);  // one arrival per epilogue warp
    }
    mbar_init(a_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      if (A_RES) {
        mbar_arrive_expect_tx(a_full, (uint32_t)kblocks * KB_BYTES);
        for (int kb = 0; kb < kblocks; ++kb) tma_load_2d(&mapsA[job.a_map], a_full, a_res + (size_t)kb * KB_BYTES, kb * TK, m0);
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int nt = 0; nt < n_tiles; ++nt)
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
          uint8_t* st = tiles + (size_t)stage * STAGE_BYTES;
          tma_load_2d(&mapsB[job.b_map], &full_bar[stage], st, kb * TK, nt * TN);
          if (!A_RES) tma_load_2d(&mapsA[job.a_map], &full_bar[stage], st + KB_BYTES, kb * TK, m0);
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
        const int buf = nt & 1;
        const uint32_t acc_phase = (uint32_t)(nt >> 1) & 1u;
        mbar_wait(&tmem_empty[buf], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(buf * TN);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          uint8_t* st = tiles + (size_t)stage * STAGE_BYTES;
          const uint64_t db = umma_desc_k_sw128(st);
          const uint64_t da = umma_desc_k_sw128(A_RES ? a_res + (size_t)kb * KB_BYTES : st + KB_BYTES);
          // a k-block row = [hi(8) hi(8) | lo(8) lo(8)] of 16 dimensions; 8 TF32 = 32 bytes = +2 in the (address >> 4) field.
          // hi.hi + lo.hi + hi.lo per 8-dimension slice: the dot product to ~2^-20 relative.
          constexpr int PA[6] = {0, 1, 2, 3, 0, 1};
          constexpr int PB[6] = {0, 1, 0, 1, 2, 3};
#pragma unroll
          for (int p = 0; p < 6; ++p)
            tc_mma_tf32(tmem_d, da + (uint64_t)(PA[p] * 2), db + (uint64_t)(PB[p] * 2), idesc, (kb | p) != 0 ? 1u : 0u);
          tc_commit(&empty_bar[stage]);  // the smem slot is free once these MMAs have read it
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit(&tmem_full[buf]);  // accumulator ready for the epilogue
      }
    }
  } else {
    // ===== epilogue: one query row per thread =====
    // The accumulator holds acc = (1 - e) ||b'||^2 - 2 a'.b' (the scaled norm rides along as three extra dimensions), a
    // lower bound of (exact distance - (1 - e) ||a'||^2); acc + 2 e (||a'||^2 + ||b'||^2) is an upper bound.  Per row:
    //   * t5[]: the KCAP smallest UPPER bounds seen so far; a column can be among the exact nearest only if its lower
    //     bound acc <= t5[KCAP-1], so only those columns are appended to the row's list (the error term is per column:
    //     a far column with a large norm does not loosen the bound for the near ones);
    //   * the list is compacted against the (shrinking) bound when it runs low on room; if that does not help — ties:
    //     clustered or duplicated descriptors — the row goes to the exact scan (D = 33) or the listed columns are
    //     evaluated EXACTLY there and then ("early flush"), after which the exact k-th distance bounds acc directly;
    //   * after the last tile the survivors are evaluated exactly, in ascending column order with strict <, which is
    //     the brute-force scan's (distance, index) order bit for bit.
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int lrow = q * 32 + lane;
    const int row = m0 + lrow;
    const bool live = row < job.na;
    const float na = live ? job.normA[row] : 0.f;
    const float na_low = (1.0f - TC_ERR_STORE) * na;  // acc + na_low <= exact distance
    const float INF = __int_as_float(0x7f800000);
    float t5[KCAP], bd[KCAP];
    int bi[KCAP];
#pragma unroll
    for (int i = 0; i < KCAP; ++i) { t5[i] = INF; bd[i] = INF; bi[i] = -1; }
    const float* a = job.origA + (size_t)(live ? row : 0) * D;
    float areg[DREG > 0 ? DREG : 1];
    if (DREG > 0) {
      constexpr int Q = DREG > 0 ? (DREG + 3) / 4 : 1;
      const float4* ap = job.padA + (size_t)(live ? row : 0) * Q;
      float tmp[Q * 4];
#pragma unroll
      for (int t = 0; t < Q; ++t) {
        const float4 v = ap[t];
        tmp[4 * t] = v.x; tmp[4 * t + 1] = v.y; tmp[4 * t + 2] = v.z; tmp[4 * t + 3] = v.w;
      }
#pragma unroll
      for (int t = 0; t < DREG; ++t) areg[t] = tmp[t];
    }
    float thr = live ? INF : -INF;   // acc-space filter bound, only ever shrinks
    float thr_exact = INF;
    bool dense = false;  // D = 33 path: the row is tie-heavy and goes to the exact scan instead
    int n = 0;
    unsigned evals = 0, early = 0;
    float* lv = list_v + lrow;
    int* lj = list_j + lrow;

    const uint32_t tmem_row = tmem_base + ((uint32_t)(q * 32) << 16);
    const int n_chunks = n_tiles * (TN / 32);
    uint32_t r[32], rn[32];
    // chunk = 32 columns of one tile; the next chunk's TMEM load is in flight while the current one is filtered
    mbar_wait(&tmem_full[0], 0);
    tc_fence_after();
    tmem_ld_32x32_issue(tmem_row, r);
    tmem_ld_wait(r);
    for (int ch = 0; ch < n_chunks; ++ch) {
      const int nt = ch >> 2, c0 = (ch & 3) * 32;
      const bool tile_end = (ch & 3) == 3;
      if (ch + 1 < n_chunks) {
        const int nt1 = (ch + 1) >> 2, buf1 = nt1 & 1;
        if (tile_end) {
          mbar_wait(&tmem_full[buf1], (uint32_t)(nt1 >> 1) & 1u);
          tc_fence_after();
        }
        tmem_ld_32x32_issue(tmem_row + (uint32_t)(buf1 * TN + ((ch + 1) & 3) * 32), rn);
      }
      // ---- filter: one compare per column, four independent mask chains
      uint32_t m0_ = 0, m1_ = 0, m2_ = 0, m3_ = 0;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        m0_ |= (__uint_as_float(r[c]) <= thr) ? (1u << c) : 0u;
        m1_ |= (__uint_as_float(r[8 + c]) <= thr) ? (1u << (8 + c)) : 0u;
        m2_ |= (__uint_as_float(r[16 + c]) <= thr) ? (1u << (16 + c)) : 0u;
        m3_ |= (__uint_as_float(r[24 + c]) <= thr) ? (1u << (24 + c)) : 0u;
      }
      uint32_t mask = (m0_ | m1_) | (m2_ | m3_);
      const int jbase = nt * TN + c0;
      const int left = job.nb - jbase;  // columns past nb are zero rows of the B form
      if (left < 32) mask &= left <= 0 ? 0u : ((1u << left) - 1u);
      // (marking a row dense as soon as one chunk passes >= 8 columns was measured: no faster on config 2, and 25 % false
      //  positives on small sets, where the bound is still loose after 128 columns)
      while (mask) {
        const int c = __ffs(mask) - 1;
        mask &= mask - 1;
        // r[c] without dynamic register indexing: a 5-level select tree
        float s16[16], s8[8], s4[4], s2[2];
#pragma unroll
        for (int i = 0; i < 16; ++i) s16[i] = __uint_as_float((c & 1) ? r[2 * i + 1] : r[2 * i]);
#pragma unroll
        for (int i = 0; i < 8; ++i) s8[i] = (c & 2) ? s16[2 * i + 1] : s16[2 * i];
#pragma unroll
        for (int i = 0; i < 4; ++i) s4[i] = (c & 4) ? s8[2 * i + 1] : s8[2 * i];
#pragma unroll
        for (int i = 0; i < 2; ++i) s2[i] = (c & 8) ? s4[2 * i + 1] : s4[2 * i];
        const float v = (c & 16) ? s2[1] : s2[0];
        if (v <= thr) {  // thr may have shrunk since the mask was built
          lv[n * TM] = v;
          lj[n * TM] = jbase + c;
          ++n;
          // upper bound of this column in acc space
          const float up = v + (2.0f * TC_ERR_STORE) * (na + __ldg(&job.normB[jbase + c])) + 1e-6f;
          if (up < t5[KCAP - 1]) {
            float cv = up;
#pragma unroll
            for (int t = 0; t < KCAP; ++t) {
              const float lo_ = fminf(t5[t], cv);
              cv = fmaxf(t5[t], cv);
              t5[t] = lo_;
            }
            thr = fminf(t5[KCAP - 1], thr_exact);
          }
        }
      }
      const bool last = ch == n_chunks - 1;
      if (n > LIST_CAP - 32 || (last && live)) {
        int w = 0;
        for (int t = 0; t < n; ++t) {
          const float v = lv[t * TM];
          if (v <= thr) {
            lv[w * TM] = v;
            lj[w * TM] = lj[t * TM];
            ++w;
          }
        }
        n = w;
        if (DREG > 0 && n > LIST_CAP - 32 && !last) {
          // more than LIST_CAP - 32 columns tie with the k-th nearest inside the error margin (clustered descriptors):
          // evaluating them here would stall the pipeline on a few diverged lanes — the exact scan kernel takes the row
          dense = true;
          ++early;
          n = 0;
          thr = -INF;
        } else if (n > LIST_CAP - 32 || last) {
          if (!last) ++early;
          for (int t = 0; t < n; ++t) {
            const int jcol = lj[t * TM];
            const float d = exact_dist<DREG>(areg, a, job, jcol, D);
            ++evals;
            if (d < bd[KCAP - 1]) {
              float cd = d;
              int ci = jcol;
#pragma unroll
              for (int u = 0; u < KCAP; ++u) {  // strict <: an equal distance stays behind the lower column already there
                if (cd < bd[u]) {
                  const float td = bd[u];
                  const int ti = bi[u];
                  bd[u] = cd;
                  bi[u] = ci;
                  cd = td;
                  ci = ti;
                }
              }
            }
          }
          n = 0;
          thr_exact = (bd[KCAP - 1] - na_low) + 1e-6f * (1.0f + fabsf(bd[KCAP - 1]) + na);  // acc <= exact k-th - na_low (+ rounding pad)
          thr = fminf(thr, thr_exact);
        }
      }
      tmem_ld_wait(rn);  // rn has landed (and, at a tile end, every read of this tile's accumulator is done)
      if (tile_end) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[nt & 1]);
      }
#pragma unroll
      for (int c = 0; c < 32; ++c) r[c] = rn[c];
    }
    if (live && dense) {
      job.dense_rows[atomicAdd(job.dense_count, 1u)] = row;
    } else if (live) {
      const int k = job.k;
#pragma unroll
      for (int t = 0; t < KCAP; ++t)
        if (t < k) {
          job.idx[(size_t)row * k + t] = bi[t];
          job.dist[(size_t)row * k + t] = bi[t] >= 0 ? bd[t] : 0.f;
        }
    }
    if (stats) {
      const unsigned rows = __reduce_add_sync(0xffffffffu, live ? 1u : 0u);
      const unsigned ef = __reduce_add_sync(0xffffffffu, live ? early : 0u);
      const unsigned ev = __reduce_add_sync(0xffffffffu, live ? evals : 0u);
      if (lane == 0) {
        atomicAdd(&stats[0], (unsigned long long)rows);
        atomicAdd(&stats[1], (unsigned long long)ef);
        atomicAdd(&stats[2], (unsigned long long)ev);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
  }
}

// Exact FP32 scan of the rows the tensor-core pass handed over (same arithmetic and tie order as knn_small_kernel in
// matching.cu): one thread per listed row, B streams through shared memory dimension-major.
template <int D, int K>
__global__ void __launch_bounds__(128) knn_dense_rows_kernel(const TcJob* __restrict__ jobs)
{
  constexpr int TB = 64;
  __shared__ __align__(16) float sb[D * TB];
  const TcJob j = jobs[blockIdx.y];
  const int n_dense = (int)*j.dense_count;
  if ((int)(blockIdx.x * blockDim.x) >= n_dense) return;
  const int li = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = li < n_dense;
  const int row = live ? j.dense_rows[li] : 0;
  float a[D];
#pragma unroll
  for (int t = 0; t < D; ++t) a[t] = live ? j.origA[(size_t)row * D + t] : 0.f;
  float bd[K];
  int bi[K];
#pragma unroll
  for (int t = 0; t < K; ++t) {
    bd[t] = __int_as_float(0x7f800000);
    bi[t] = -1;
  }
  for (int base = 0; base < j.nb; base += TB) {
    const int tb = min(TB, j.nb - base);
    __syncthreads();
    for (int e = threadIdx.x; e < TB * D; e += blockDim.x) {
      const int r = e / D, t = e - r * D;
      sb[t * TB + r] = (r < tb) ? j.origB[(size_t)(base + r) * D + t] : 0.f;
    }
    __syncthreads();
    if (!live) continue;
    for (int r0 = 0; r0 < tb; r0 += 4) {
      float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
#pragma unroll
      for (int t = 0; t < D; ++t) {
        const float4 b = *reinterpret_cast<const float4*>(&sb[t * TB + r0]);
        const float d0 = a[t] - b.x, d1 = a[t] - b.y, d2 = a[t] - b.z, d3 = a[t] - b.w;
        acc0 += d0 * d0;
        acc1 += d1 * d1;
        acc2 += d2 * d2;
        acc3 += d3 * d3;
      }
      const float accs[4] = {acc0, acc1, acc2, acc3};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float cd = accs[u];
        if (r0 + u < tb && cd < bd[K - 1]) {
          int ci = base + r0 + u;
#pragma unroll
          for (int t = 0; t < K; ++t) {  // strict <: ties keep the lower column
            if (cd < bd[t]) {
              const float td = bd[t];
              const int ti = bi[t];
              bd[t] = cd;
              bi[t] = ci;
              cd = td;
              ci = ti;
            }
          }
        }
      }
    }
  }
  if (live) {
    const int k = j.k;
#pragma unroll
    for (int t = 0; t < K; ++t)
      if (t < k) {
        j.idx[(size_t)row * k + t] = bi[t];
        j.dist[(size_t)row * k + t] = bi[t] >= 0 ? bd[t] : 0.f;
      }
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
void knn_tc_batch(Ctx& c, const std::vector<const float*>& desc, const std::vector<int>& n_rows, int D, const std::vector<KnnProblem>& probs)
{
  if (probs.empty()) return;
  const int M = (int)desc.size();
  const int kblocks = (D + 3 + 15) / 16;
  const int Kp = kblocks * TK;
  const bool reg_path = D == 33;
  const int padw = reg_path ? 36 : 0;
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
  std::vector<TcJob> tj;
  int max_na = 0, kmax = 0;
  double bytes = 0;
  size_t rows_total = 0, n_jobs = 0;
  for (const KnnProblem& p : probs)
    if (p.na > 0 && n_rows[p.b] > 0) { rows_total += (size_t)p.na; ++n_jobs; }
  DBuf<int> dense_rows(c, reg_path ? rows_total : 0);
  DBuf<unsigned int> dense_count(c, n_jobs);
  dense_count.zero(c);
  size_t row_off = 0;
  for (const KnnProblem& p : probs) {
    if (p.na == 0 || n_rows[p.b] == 0) continue;
    TcJob j;
    j.dense_rows = reg_path ? dense_rows.p + row_off : nullptr;
    j.dense_count = dense_count.p + tj.size();
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
    j.normB = norm[p.b].p;
    j.k = p.k;
    j.idx = p.idx;
    j.dist = p.dist;
    tj.push_back(j);
    max_na = std::max(max_na, p.na);
    kmax = std::max(kmax, p.k);
    bytes += 4.0 * D * ((double)j.na + j.nb) + 8.0 * j.k * j.na;
  }
  if (tj.empty()) return;
  DBuf<TcJob> dtj = to_device(c, tj);
  MM_BYTES(c, bytes);
  if (!c.knn_stats) {
    MM_CUDA(cudaMalloc((void**)&c.knn_stats, 3 * sizeof(unsigned long long)));
    MM_CUDA(cudaMemsetAsync(c.knn_stats, 0, 3 * sizeof(unsigned long long), c.stream));
  }
  const dim3 grid((max_na + TM - 1) / TM, (unsigned)tj.size());
#define MM_TC(KCAP, DREG, RES)                                                                                                   \
  do {                                                                                                                           \
    const size_t smem = RES ? TC_SMEM_RES : TC_SMEM_STREAM;                                                                      \
    /* per device and cheap: set on every call (a process may drive several GPUs) */                                            \
    MM_CUDA(cudaFuncSetAttribute(knn_tc_kernel<KCAP, DREG, RES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));       \
    MM_LAUNCH(c, (knn_tc_kernel<KCAP, DREG, RES>), grid, TC_THREADS, smem, dtj.p, dA.p, dB.p, kblocks, D, c.knn_stats);          \
  } while (0)
  const bool res = kblocks <= A_RES_MAX_KB;
  const dim3 dgrid((max_na + 127) / 128, (unsigned)tj.size());  // blocks past a job's dense count exit at once
  if (reg_path && kmax <= 5) {
    MM_TC(5, 33, true);
    MM_LAUNCH(c, (knn_dense_rows_kernel<33, 5>), dgrid, 128, 0, dtj.p);
  } else if (reg_path && kmax <= 10) {
    MM_TC(10, 33, true);
    MM_LAUNCH(c, (knn_dense_rows_kernel<33, 10>), dgrid, 128, 0, dtj.p);
  } else if (reg_path) {
    MM_TC(KMAXTC, 33, true);
    MM_LAUNCH(c, (knn_dense_rows_kernel<33, KMAXTC>), dgrid, 128, 0, dtj.p);
  }
  else if (res && kmax <= 5) MM_TC(5, 0, true);
  else if (res) MM_TC(KMAXTC, 0, true);
  else if (kmax <= 1) MM_TC(1, 0, false);
  else if (kmax <= 5) MM_TC(5, 0, false);
  else if (kmax <= 10) MM_TC(10, 0, false);
  else MM_TC(KMAXTC, 0, false);
#undef MM_TC
}

}  // namespace mm3d
