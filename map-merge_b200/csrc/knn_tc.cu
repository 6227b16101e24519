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
constexpr int LIST_CAP = 64;                // pending candidates per (query row, column half), 16-bit column indices (nb <= 65535)
constexpr int LIST_STRIDE = LIST_CAP + 2;   // 33 words per list: appends by 32 rows (2-way) and reads of one row's 32 entries (none) stay cheap
constexpr int KMAXTC = 16;
constexpr int ACC_BUFS = 4;                 // TMEM accumulator buffers (4 x 128 columns = all 512): the epilogue warps may lag the
                                            // MMA by three tiles, so one warp's burst of exact evaluations no longer stalls the rest
constexpr int EPI_WARPS = 8;                // two per scheduler: warps w and w + 4 share a TMEM lane quarter and split every tile's columns
constexpr int EVAL_WARPS = 6;                // candidate pass only: warps that evaluate pending candidates exactly, any row of the CTA
constexpr int TC_THREADS = 64 + (EPI_WARPS + EVAL_WARPS) * 32;  // warp 0: TMA producer, 1: MMA issuer + TMEM owner, 2-9: epilogue, 10-15: evaluators
constexpr int N_SLOTS = TM * 2;              // one pending list per (query row, column half)
constexpr int LIST_BYTES = TM * 2 * LIST_STRIDE * 2;
constexpr int PARK_BYTES = N_SLOTS * KMAXTC * 8;  // per slot: the k best (distance, index) so far (candidate pass) / parked bounds (threshold pass)
constexpr int SLOT_STATE_BYTES = N_SLOTS * 16 + TM * 4 + 64;  // tail, head, exact bound, busy flag per slot; (1 - ES)||a'||^2 per row; counters
constexpr int APAD = 36;                    // floats per row of the shared-memory copy of the original query rows (D = 33 path)
constexpr int AORIG_BYTES = TM * APAD * 4;
// shared memory: [A resident (3 k-blocks for D = 33, else 4)] [B (or A + B) stages] [pending lists] [parking] [query rows] [barriers]
__host__ __device__ constexpr int tc_a_kb(int dreg, bool a_res) { return a_res ? (dreg > 0 ? 3 : A_RES_MAX_KB) : 0; }
__host__ __device__ constexpr int tc_stages(int dreg, bool a_res) { return a_res ? (dreg > 0 ? 5 : 4) : 3; }
__host__ __device__ constexpr size_t tc_smem(int dreg, bool a_res)
{
  return (size_t)tc_a_kb(dreg, a_res) * KB_BYTES + (size_t)tc_stages(dreg, a_res) * (a_res ? 1 : 2) * KB_BYTES + LIST_BYTES + PARK_BYTES +
         AORIG_BYTES + SLOT_STATE_BYTES + 1024 + 256;
}

struct TcJob {
  int a_map, b_map;  // tensor maps: A-form of a_map, B-form of b_map
  int na, nb;
  const float* normA;  // ||a - mean||^2
  const float* origA;  // na x D original descriptors
  const float* origB;
  const float4* padA;  // rows padded to a multiple of 4 floats (D = 33 only), else null
  const float4* padB;
  const float* cmaxB;  // max ||b - mean||^2 over every 32 rows of B
  int k;
  int* idx;     // na x k
  float* dist;  // na x k
  float* audit; // optional na x nb: the raw accumulator (tests: error-bound audit)
  float* thr;   // na: per row, upper bound (accumulator space) of the k-th nearest distance — written by pass 0, read by pass 1
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
// FP32 accumulation inside the tensor core (18 chained MMAs of 8 products; each partial sum is bounded by ||b'||^2 + 2 ||a'|| ||b'||),
// and the FP32 rounding of the norms and of d itself (<= 35 * 2^-24 relative).  With 2 ||a'|| ||b'|| <= ||a'||^2 + ||b'||^2 the sum
// of the three stays below EG (||a'||^2 + ||b'||^2); ES = EG + margin, so v is a lower bound and v + 2 ES (...) an upper bound.
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

// KCAP = capacity of the register top lists (>= k; the first k are written out); DREG = descriptor length when the query
// rows sit in shared memory and rows are read as float4 from the padded copies, 0 = scalar reads from global memory;
// A_RES = the A tile stays in shared memory for the whole CTA (kblocks <= A_RES_MAX_KB); AUDIT = also dump the accumulators.
// PASS 0 = threshold pass: the same GEMM with a minimal epilogue that only derives, per row, an upper bound of the k-th nearest
// distance (job.thr); PASS 1 = candidate pass: columns whose lower bound is within that bound are evaluated exactly.
// Two passes instead of one running threshold: a streaming top-k meets ~k ln(n / k) record breakers per row that all need
// an exact evaluation and a list insertion (measured: 2/3 of the single-pass epilogue's instructions); with the bound known
// up front only the columns inside the error margin of the k-th distance are ever touched, and the tensor pipe — 5 % busy
// in the single-pass kernel — pays for the second GEMM.
template <int KCAP, int DREG, bool A_RES, bool AUDIT, int PASS>
__global__ void __launch_bounds__(TC_THREADS, 1) knn_tc_kernel(const TcJob* __restrict__ jobs, const CUtensorMap* __restrict__ mapsA,
                                                              const CUtensorMap* __restrict__ mapsB, int kblocks, int D,
                                                              unsigned long long* __restrict__ stats)
{
  constexpr int STAGES = tc_stages(DREG, A_RES);
  constexpr int STAGE_BYTES = A_RES ? KB_BYTES : 2 * KB_BYTES;
  constexpr int A_RES_BYTES = tc_a_kb(DREG, A_RES) * KB_BYTES;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by OFFSET, so that every pointer below stays a shared-memory pointer for the compiler (rounding the
  // address through uintptr_t turned all list / query-row accesses into generic LD.E / ST.E)
  uint8_t* a_res = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* tiles = a_res + A_RES_BYTES;
  unsigned short* list = (unsigned short*)(tiles + (size_t)STAGES * STAGE_BYTES);  // [TM][2][LIST_STRIDE]
  float* park_d = (float*)(list + TM * 2 * LIST_STRIDE);                           // [N_SLOTS][KCAP] distances, then indices
  int* park_i = (int*)(park_d + N_SLOTS * KMAXTC);
  float* a_orig = (float*)(park_i + N_SLOTS * KMAXTC);                             // [TM][APAD] (D = 33 path)
  uint32_t* s_tail = (uint32_t*)(a_orig + TM * APAD);                              // [N_SLOTS] entries published by the producer
  uint32_t* s_head = s_tail + N_SLOTS;                                             // [N_SLOTS] entries consumed by the evaluators
  float* s_thr = (float*)(s_head + N_SLOTS);                                       // [N_SLOTS] exact k-th distance so far, accumulator space
  uint32_t* s_busy = (uint32_t*)(s_thr + N_SLOTS);                                 // [N_SLOTS] an evaluator warp owns the slot
  float* s_nalow = (float*)(s_busy + N_SLOTS);                                     // [TM]
  uint32_t* s_done = (uint32_t*)(s_nalow + TM);                                    // producers that have finished
  uint64_t* full_bar = (uint64_t*)(s_done + 16);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + ACC_BUFS;
  uint64_t* a_full = tmem_empty + ACC_BUFS;
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
    for (int b = 0; b < ACC_BUFS; ++b) {
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_empty[b], EPI_WARPS);  // one arrival per epilogue warp
    }
    mbar_init(a_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)(ACC_BUFS * TN)) : "memory");
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
        const int buf = nt % ACC_BUFS;
        const uint32_t acc_phase = (uint32_t)(nt / ACC_BUFS) & 1u;
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
  } else if (warp < 2 + EPI_WARPS) {
    // ===== epilogue: one query row per thread, two warps per TMEM lane quarter (each takes 64 of a tile's 128 columns) =====
    // Per (row, column half):
    //   * t5[]: the KCAP smallest UPPER bounds seen so far, fed with one value per 32-column chunk — the chunk's smallest
    //     accumulator plus the error term of the chunk's largest norm.  Chunk minima belong to distinct columns, so their
    //     KCAP-th smallest upper bound bounds the row's k-th nearest distance; thr = min(that, the exact k-th distance so far);
    //   * a column can be among the exact nearest only if its lower bound acc <= thr: such columns are appended (index
    //     only) to the row's pending list in shared memory.  The common case — no column of the chunk passes — costs the
    //     min tree (16 three-input minima) and ten min/max for the bound, no branch per column;
    //   * when a list holds 32 candidates the WARP evaluates them together: lane c computes the exact sequential FP32
    //     distance of candidate c (query row broadcast from shared memory, B row as float4 loads), the few that beat the
    //     row's current k-th distance are inserted by the owning lane in ascending column order with strict <, which is
    //     the brute-force scan's (distance, index) order bit for bit.  Tie-heavy rows (clustered descriptors: hundreds of
    //     columns inside the error margin of the k-th distance) therefore cost one full-warp evaluation per 32 ties
    //     instead of diverged per-lane work or a brute-force fallback.
    const int ew = warp - 2;
    const int q = warp & 3;        // TMEM lane quarter this warp may access
    const int half = ew >> 2;      // which 64 columns of every tile
    const int lrow = q * 32 + lane;
    const int row = m0 + lrow;
    const bool live = row < job.na;
    const float na = live ? job.normA[row] : 0.f;
    const float na_low = (1.0f - TC_ERR_STORE) * na;  // acc + na_low <= exact distance
    const float INF = __int_as_float(0x7f800000);
    if constexpr (PASS == 0) {
      // ---- threshold pass: K-th smallest upper bound over one value per 32-column chunk (chunk minima are distinct columns)
      float t5[KCAP];
#pragma unroll
      for (int i = 0; i < KCAP; ++i) t5[i] = INF;
      const uint32_t tmem_row = tmem_base + ((uint32_t)(q * 32) << 16);
      const int n_chunks = n_tiles * 2;
      uint32_t r[32], rn[32];
      mbar_wait(&tmem_full[0], 0);
      tc_fence_after();
      tmem_ld_32x32_issue(tmem_row + (uint32_t)(half * 64), r);
      tmem_ld_wait(r);
      for (int ch = 0; ch < n_chunks; ++ch) {
        const int nt = ch >> 1, c0 = half * 64 + (ch & 1) * 32;
        const bool tile_end = (ch & 1) == 1;
        if (ch + 1 < n_chunks) {
          const int nt1 = (ch + 1) >> 1, buf1 = nt1 % ACC_BUFS;
          if (tile_end) {
            mbar_wait(&tmem_full[buf1], (uint32_t)(nt1 / ACC_BUFS) & 1u);
            tc_fence_after();
          }
          tmem_ld_32x32_issue(tmem_row + (uint32_t)(buf1 * TN + half * 64 + ((ch + 1) & 1) * 32), rn);
        }
        const int jbase = nt * TN + c0;
        const int left = job.nb - jbase;  // columns past nb are zero rows of the B form: keep them out of the minimum
        if (left < 32) {
#pragma unroll
          for (int c = 0; c < 32; ++c)
            if (c >= left) r[c] = 0x7f800000u;
        }
        if (left > 0) {
          float m = fmin3(__uint_as_float(r[0]), __uint_as_float(r[1]), __uint_as_float(r[2]));
#pragma unroll
          for (int c = 3; c < 31; c += 2) m = fmin3(m, __uint_as_float(r[c]), __uint_as_float(r[c + 1]));
          m = fminf(m, __uint_as_float(r[31]));
          const float cmax = __ldg(&job.cmaxB[jbase >> 5]);
          float ub = m + ((2.0f * TC_ERR_STORE) * (na + cmax) + 1e-6f * (1.0f + fabsf(m)));
#pragma unroll
          for (int t = 0; t < KCAP; ++t) {
            const float lo_ = fminf(t5[t], ub);
            ub = fmaxf(t5[t], ub);
            t5[t] = lo_;
          }
        }
        tmem_ld_wait(rn);
        if (tile_end) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[nt % ACC_BUFS]);
        }
#pragma unroll
        for (int c = 0; c < 32; ++c) r[c] = rn[c];
      }
      // K-th smallest of the two halves' sorted lists
      if (half == 1) {
#pragma unroll
        for (int t = 0; t < KCAP; ++t) park_d[lrow * KCAP + t] = t5[t];
      }
      named_bar_sync<EPI_WARPS * 32>();
      if (half == 0 && live) {
        int ia = 0, ib = 0;
        float kth = INF;
        for (int t = 0; t < KCAP; ++t) {
          float da = INF;
#pragma unroll
          for (int u = 0; u < KCAP; ++u)
            if (u == ia) da = t5[u];
          const float db = ib < KCAP ? park_d[lrow * KCAP + ib] : INF;
          if (db < da) { kth = db; ++ib; } else { kth = da; ++ia; }
        }
        job.thr[row] = kth;
      }
    } else {
      // ---- candidate pass, producer side: columns whose accumulator is within the row's bound go — index only — into the
      // row's pending ring in shared memory; the evaluator warps (below) pick full rings up.  The producer never evaluates
      // anything itself, so its work per chunk is uniform and the eight epilogue warps stay in step with the MMA.
      const int slot = lrow * 2 + half;
      unsigned short* my_list = list + slot * LIST_STRIDE;
      if (half == 0) s_nalow[lrow] = na_low;
      s_tail[slot] = 0;
      s_head[slot] = 0;
      s_busy[slot] = 0;
      s_thr[slot] = INF;
#pragma unroll
      for (int t = 0; t < KCAP; ++t) {
        park_d[slot * KCAP + t] = INF;
        park_i[slot * KCAP + t] = -1;
      }
      if (ew == 0 && lane == 0) *s_done = 0;
      if (DREG > 0) {
        for (int e = (ew * 32 + lane); e < TM * (APAD / 4); e += EPI_WARPS * 32) {
          const int r_ = e / (APAD / 4), t = e - r_ * (APAD / 4);
          const float4 v = (m0 + r_ < job.na) ? job.padA[(size_t)(m0 + r_) * (APAD / 4) + t] : make_float4(0.f, 0.f, 0.f, 0.f);
          reinterpret_cast<float4*>(a_orig)[r_ * (APAD / 4) + t] = v;
        }
      }
      named_bar_sync<(EPI_WARPS + EVAL_WARPS) * 32>();  // producers + evaluators: rings, bounds and the query rows are in place
      float thr = live ? job.thr[row] : -INF;  // acc-space filter bound from the threshold pass, only ever shrinks
      uint32_t tail = 0;
      const uint32_t tmem_row = tmem_base + ((uint32_t)(q * 32) << 16);
      const int n_chunks = n_tiles * 2;  // this warp's chunks: two per tile
      uint32_t r[32], rn[32];
      mbar_wait(&tmem_full[0], 0);
      tc_fence_after();
      tmem_ld_32x32_issue(tmem_row + (uint32_t)(half * 64), r);
      tmem_ld_wait(r);
      for (int ch = 0; ch < n_chunks; ++ch) {
        const int nt = ch >> 1, c0 = half * 64 + (ch & 1) * 32;
        const bool tile_end = (ch & 1) == 1;
        if (ch + 1 < n_chunks) {
          const int nt1 = (ch + 1) >> 1, buf1 = nt1 % ACC_BUFS;
          if (tile_end) {
            mbar_wait(&tmem_full[buf1], (uint32_t)(nt1 / ACC_BUFS) & 1u);
            tc_fence_after();
          }
          tmem_ld_32x32_issue(tmem_row + (uint32_t)(buf1 * TN + half * 64 + ((ch + 1) & 1) * 32), rn);
        }
        const int jbase = nt * TN + c0;
        const int left = job.nb - jbase;  // columns past nb are zero rows of the B form: keep them out of the minimum
        if (AUDIT) {
          if (live) {
#pragma unroll
            for (int c = 0; c < 32; ++c)
              if (c < left) job.audit[(size_t)row * job.nb + jbase + c] = __uint_as_float(r[c]);
          }
        }
        if (left < 32) {
#pragma unroll
          for (int c = 0; c < 32; ++c)
            if (c >= left) r[c] = 0x7f800000u;
        }
        if (left > 0) {
          thr = fminf(thr, *(volatile float*)&s_thr[slot]);  // what the evaluators have learned about this row so far
          // smallest accumulator of the chunk: 16 three-input minima
          float m = fmin3(__uint_as_float(r[0]), __uint_as_float(r[1]), __uint_as_float(r[2]));
#pragma unroll
          for (int c = 3; c < 31; c += 2) m = fmin3(m, __uint_as_float(r[c]), __uint_as_float(r[c + 1]));
          m = fminf(m, __uint_as_float(r[31]));
          if (m <= thr) {
            // which columns pass (one compare + one bit each), then one append per set bit
            uint32_t m0_ = 0, m1_ = 0, m2_ = 0, m3_ = 0;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              m0_ |= (__uint_as_float(r[c]) <= thr) ? (1u << c) : 0u;
              m1_ |= (__uint_as_float(r[8 + c]) <= thr) ? (1u << (8 + c)) : 0u;
              m2_ |= (__uint_as_float(r[16 + c]) <= thr) ? (1u << (16 + c)) : 0u;
              m3_ |= (__uint_as_float(r[24 + c]) <= thr) ? (1u << (24 + c)) : 0u;
            }
            uint32_t msk = (m0_ | m1_) | (m2_ | m3_);
            if (left < 32) msk &= (1u << left) - 1u;  // columns past nb (masked to +inf above) would still pass an infinite bound
            // room for the whole chunk (the evaluators free 32 entries at a time; a full ring means they are behind)
            const uint32_t need = (uint32_t)__popc(msk);
            while (tail + need - *(volatile uint32_t*)&s_head[slot] > (uint32_t)LIST_CAP) __nanosleep(64);
            while (msk) {
              const int c = __ffs(msk) - 1;
              msk &= msk - 1;
              my_list[tail & (LIST_CAP - 1)] = (unsigned short)(jbase + c);
              ++tail;
            }
            __threadfence_block();
            *(volatile uint32_t*)&s_tail[slot] = tail;  // publish
          }
        }
        tmem_ld_wait(rn);  // rn has landed (and, at a tile end, every read of this tile's accumulator is done)
        if (tile_end) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[nt % ACC_BUFS]);
        }
#pragma unroll
        for (int c = 0; c < 32; ++c) r[c] = rn[c];
      }
      __syncwarp();
      if (lane == 0) {
        __threadfence_block();
        atomicAdd(s_done, 1u);
      }
    }
  } else {
    // ===== evaluator warps (candidate pass only) =====
    // Any evaluator warp serves any row of the CTA: it claims a pending ring that holds 32 candidates (or, once the
    // producers are done, whatever is left), lane c computes the exact sequential FP32 distance of candidate c (query row
    // broadcast from shared memory, B row as float4 loads), and the few that beat the row's current k-th distance are
    // inserted — in ring order = ascending column order, strict < — into the row's list, which lives in shared memory.  The
    // exact k-th distance goes back to the producer as a tighter bound.  Decoupling evaluation from the epilogue warps
    // matters because the candidates are very unevenly spread over the rows (clustered descriptors): inline evaluation
    // made every tile wait for the epilogue warp with the most ties.
    if constexpr (PASS == 1) {
      const int ev = warp - 2 - EPI_WARPS;
      const float INF = __int_as_float(0x7f800000);
      unsigned evals = 0, flushes = 0;
      named_bar_sync<(EPI_WARPS + EVAL_WARPS) * 32>();
      int scan0 = ev * (N_SLOTS / EVAL_WARPS);  // evaluators start their sweeps at different slots
      for (;;) {
        const bool draining = *(volatile uint32_t*)s_done == (uint32_t)EPI_WARPS;
        // sweep: every lane looks at 8 slots
        int found = -1;
        for (int pass_ = 0; pass_ < N_SLOTS / 32 && found < 0; ++pass_) {
          const int s_ = (scan0 + pass_ * 32 + lane) & (N_SLOTS - 1);
          const uint32_t cnt = *(volatile uint32_t*)&s_tail[s_] - *(volatile uint32_t*)&s_head[s_];
          const bool want = (cnt >= 32u || (draining && cnt > 0u)) && *(volatile uint32_t*)&s_busy[s_] == 0u;
          const unsigned vote = __ballot_sync(0xffffffffu, want);
          if (vote) found = (scan0 + pass_ * 32 + (__ffs(vote) - 1)) & (N_SLOTS - 1);
        }
        if (found < 0) {
          if (draining) {
            // nothing pending anywhere and no producer left: done (a ring another evaluator is working on is its business)
            bool any = false;
            for (int s_ = lane; s_ < N_SLOTS; s_ += 32)
              if (*(volatile uint32_t*)&s_tail[s_] != *(volatile uint32_t*)&s_head[s_] && *(volatile uint32_t*)&s_busy[s_] == 0u) any = true;
            if (!__any_sync(0xffffffffu, any)) break;
          } else {
            __nanosleep(200);
          }
          continue;
        }
        scan0 = found + 1;
        int claimed = 0;
        if (lane == 0) claimed = atomicCAS(&s_busy[found], 0u, 1u) == 0u ? 1 : 0;
        claimed = __shfl_sync(0xffffffffu, claimed, 0);
        if (!claimed) continue;
        __threadfence_block();
        const uint32_t head = *(volatile uint32_t*)&s_head[found];
        const uint32_t cnt_all = *(volatile uint32_t*)&s_tail[found] - head;
        const int cnt = (int)min(cnt_all, 32u);
        if (cnt > 0) {
          const int rL = found >> 1;
          const unsigned short* lst = list + found * LIST_STRIDE;
          const int jcol = lane < cnt ? (int)lst[(head + lane) & (LIST_CAP - 1)] : 0;
          float d = INF;
          if (lane < cnt) {
            float acc = 0.f;
            if (DREG > 0) {
              constexpr int Q = DREG > 0 ? (DREG + 3) / 4 : 1;
              const float4* bp = job.padB + (size_t)jcol * Q;
              const float4* ap = reinterpret_cast<const float4*>(a_orig + rL * APAD);
              float a_[Q * 4], b_[Q * 4];
#pragma unroll
              for (int t = 0; t < Q; ++t) {
                const float4 v = __ldg(&bp[t]);
                b_[4 * t] = v.x; b_[4 * t + 1] = v.y; b_[4 * t + 2] = v.z; b_[4 * t + 3] = v.w;
                const float4 w = ap[t];
                a_[4 * t] = w.x; a_[4 * t + 1] = w.y; a_[4 * t + 2] = w.z; a_[4 * t + 3] = w.w;
              }
#pragma unroll
              for (int t = 0; t < DREG; ++t) {
                const float diff = a_[t] - b_[t];
                acc += diff * diff;
              }
            } else {
              const float* a = job.origA + (size_t)(m0 + rL) * D;
              const float* b = job.origB + (size_t)jcol * D;
              for (int t = 0; t < D; ++t) {
                const float diff = __ldg(&a[t]) - __ldg(&b[t]);
                acc += diff * diff;
              }
            }
            d = acc;
          }
          // the row's list: lane 0 works on a register copy
          float bd[KCAP];
          int bi[KCAP];
#pragma unroll
          for (int t = 0; t < KCAP; ++t) {
            bd[t] = park_d[found * KCAP + t];
            bi[t] = park_i[found * KCAP + t];
          }
          unsigned better = __ballot_sync(0xffffffffu, d < bd[KCAP - 1]);
          const bool changed = better != 0u;
          while (better) {
            const int b = __ffs(better) - 1;
            const float dv = __shfl_sync(0xffffffffu, d, b);
            const int jv = __shfl_sync(0xffffffffu, jcol, b);
            topk_insert<KCAP>(bd, bi, dv, jv);  // every lane keeps the same copy: no broadcast of the new k-th needed
            better &= __ballot_sync(0xffffffffu, d < bd[KCAP - 1]) & ~((2u << b) - 1u);
          }
          if (changed && lane == 0) {
#pragma unroll
            for (int t = 0; t < KCAP; ++t) {
              park_d[found * KCAP + t] = bd[t];
              park_i[found * KCAP + t] = bi[t];
            }
            const float nl = s_nalow[rL];
            // acc <= exact k-th - (1 - ES)||a'||^2 (+ rounding pad)
            *(volatile float*)&s_thr[found] = (bd[KCAP - 1] - nl) + 1e-6f * (1.0f + fabsf(bd[KCAP - 1]) + nl);
          }
          evals += (unsigned)cnt;
          ++flushes;
        }
        __syncwarp();
        if (lane == 0) {
          __threadfence_block();
          *(volatile uint32_t*)&s_head[found] = head + (uint32_t)cnt;  // frees the ring entries
          __threadfence_block();
          *(volatile uint32_t*)&s_busy[found] = 0u;
        }
        __syncwarp();
      }
      // all evaluators done: merge the two column halves of every row and write the result
      asm volatile("bar.sync 2, %0;" ::"n"(EVAL_WARPS * 32) : "memory");
      for (int lr = (int)threadIdx.x - (2 + EPI_WARPS) * 32; lr < TM; lr += EVAL_WARPS * 32) {
        const int row = m0 + lr;
        if (row >= job.na) continue;
        const int k = job.k;
        const float* da_ = park_d + (lr * 2) * KCAP;
        const int* ja_ = park_i + (lr * 2) * KCAP;
        const float* db_ = park_d + (lr * 2 + 1) * KCAP;
        const int* jb_ = park_i + (lr * 2 + 1) * KCAP;
        int ia = 0, ib = 0;
        for (int t = 0; t < k; ++t) {
          const float da = ia < KCAP ? da_[ia] : INF, db = ib < KCAP ? db_[ib] : INF;
          const int ja = ia < KCAP ? ja_[ia] : -1, jb = ib < KCAP ? jb_[ib] : -1;
          const bool take_b = jb >= 0 && (ja < 0 || db < da || (db == da && jb < ja));
          const float dsel = take_b ? db : da;
          const int jsel = take_b ? jb : ja;
          if (take_b) ++ib; else ++ia;
          job.idx[(size_t)row * k + t] = jsel;
          job.dist[(size_t)row * k + t] = jsel >= 0 ? dsel : 0.f;
        }
      }
      if (stats) {
        const unsigned fl = __reduce_add_sync(0xffffffffu, lane == 0 ? flushes : 0u);
        const unsigned evs = __reduce_add_sync(0xffffffffu, lane == 0 ? evals : 0u);
        if (lane == 0) {
          if (ev == 0) atomicAdd(&stats[0], (unsigned long long)min(TM, job.na - m0));
          atomicAdd(&stats[1], (unsigned long long)fl);
          atomicAdd(&stats[2], (unsigned long long)evs);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(ACC_BUFS * TN)) : "memory");
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

// largest centred norm of every 32 rows (one epilogue chunk): bounds the error term of all columns of the chunk
struct CmaxJob {
  const float* norm;
  float* cmax;
  int n;
};
__global__ void __launch_bounds__(128) knn_tc_cmax_kernel(const CmaxJob* __restrict__ jobs)
{
  const CmaxJob& j = jobs[blockIdx.y];
  const int chunk = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (chunk * 32 >= j.n) return;
  const int r = chunk * 32 + lane;
  float v = r < j.n ? j.norm[r] : 0.f;
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  if (lane == 0) j.cmax[chunk] = v;
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
  const int Kp = kblocks * TK;
  const bool reg_path = D == 33;
  const int padw = reg_path ? APAD : 0;
  std::vector<char> used(M, 0);
  for (const KnnProblem& p : probs) { used[p.a] = 1; used[p.b] = 1; }
  std::vector<DBuf<float>> formA(M), formB(M), norm(M), pad(M), cmax(M);
  std::vector<PrepJob> pj;
  std::vector<CmaxJob> cj;
  int mxn = 0;
  for (int m = 0; m < M; ++m) {
    if (!used[m] || n_rows[m] == 0) continue;
    formA[m].alloc(c, (size_t)n_rows[m] * Kp);
    formB[m].alloc(c, (size_t)n_rows[m] * Kp);
    norm[m].alloc(c, n_rows[m]);
    cmax[m].alloc(c, (size_t)(n_rows[m] + 31) / 32);
    if (reg_path) pad[m].alloc(c, (size_t)n_rows[m] * padw);
    pj.push_back(PrepJob{desc[m], formA[m].p, formB[m].p, norm[m].p, reg_path ? pad[m].p : nullptr, n_rows[m]});
    cj.push_back(CmaxJob{norm[m].p, cmax[m].p, n_rows[m]});
    mxn = std::max(mxn, n_rows[m]);
  }
  if (pj.empty()) return;
  DBuf<PrepJob> dpj = to_device(c, pj);
  DBuf<CmaxJob> dcj = to_device(c, cj);
  size_t rows_all = 0;
  for (const PrepJob& j : pj) rows_all += (size_t)j.n;
  DBuf<double> partial(c, pj.size() * MEAN_CHUNKS * (size_t)D);
  DBuf<float> mean(c, D);
  MM_LAUNCH(c, knn_tc_mean_partial_kernel, dim3(MEAN_CHUNKS, (unsigned)pj.size()), 128, 0, dpj.p, D, partial.p);
  MM_LAUNCH(c, knn_tc_mean_kernel, (D + 127) / 128, 128, 0, partial.p, (int)pj.size() * MEAN_CHUNKS, D, 1.0 / (double)rows_all, mean.p);
  MM_LAUNCH(c, knn_tc_prep_kernel, dim3(mxn, (unsigned)pj.size()), 128, 0, dpj.p, mean.p, D, Kp, padw);
  MM_LAUNCH(c, knn_tc_cmax_kernel, dim3((mxn + 127) / 128, (unsigned)cj.size()), 128, 0, dcj.p);
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
  DBuf<float> audit_acc;
  size_t rows_total = 0;
  for (const KnnProblem& p : probs)
    if (p.na > 0 && n_rows[p.b] > 0) rows_total += (size_t)p.na;
  DBuf<float> thr(c, rows_total);
  size_t row_off = 0;
  for (const KnnProblem& p : probs) {
    if (p.na == 0 || n_rows[p.b] == 0) continue;
    TcJob j;
    j.thr = thr.p + row_off;
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
    j.cmaxB = cmax[p.b].p;
    j.k = p.k;
    j.idx = p.idx;
    j.dist = p.dist;
    j.audit = nullptr;
    tj.push_back(j);
    max_na = std::max(max_na, p.na);
    kmax = std::max(kmax, p.k);
    bytes += 4.0 * D * ((double)j.na + j.nb) + 8.0 * j.k * j.na;
  }
  if (tj.empty()) return;
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
  const dim3 grid((max_na + TM - 1) / TM, (unsigned)tj.size());
#define MM_TC(KCAP, DREG, RES, AUD)                                                                                                    \
  do {                                                                                                                                 \
    const size_t smem = tc_smem(DREG, RES);                                                                                            \
    /* per device and cheap: set on every call (a process may drive several GPUs) */                                                  \
    MM_CUDA(cudaFuncSetAttribute(knn_tc_kernel<KCAP, DREG, RES, false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
    MM_CUDA(cudaFuncSetAttribute(knn_tc_kernel<KCAP, DREG, RES, AUD, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
    MM_BYTES(c, bytes);                                                                                                                \
    MM_LAUNCH(c, (knn_tc_kernel<KCAP, DREG, RES, false, 0>), grid, TC_THREADS, smem, dtj.p, dA.p, dB.p, kblocks, D, nullptr);          \
    MM_BYTES(c, bytes);                                                                                                                \
    MM_LAUNCH(c, (knn_tc_kernel<KCAP, DREG, RES, AUD, 1>), grid, TC_THREADS, smem, dtj.p, dA.p, dB.p, kblocks, D, c.knn_stats);        \
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
