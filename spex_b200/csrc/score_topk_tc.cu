// score_topk_tc.cu — full-ranking top-k on the 5th-generation tensor cores (tcgen05 / TMEM).
//
// north_star (3): getUsersRating (abstract at LightGCN_SPEX/code/utility1/model.py:14-15; the only
// user x all-items matmul in the reference is NGCF_SPEX/code/utility/batch_test.py:158) followed by
// top-k, is a dense [users,64] x [64,items] contraction, so it goes on the tensor cores:
//
//   * operands: bf16, K-major, pre-packed by spex_pack_bf16 into the UMMA "no-swizzle" canonical
//     layout (8-row x 16-byte core matrices; one 8-row group = 1 KB contiguous).  A tile of any
//     multiple of 8 rows is therefore ONE contiguous block of global memory and is staged by a
//     single 1-D bulk async copy (cp.async.bulk -> UBLKCP) that completes on an mbarrier; no
//     tensor map, no swizzle bookkeeping;
//   * tile: one CTA owns 128 users (UMMA M=128) and streams the item table in tiles of N=128
//     through a multi-stage smem ring.  Per tile: 4 tcgen05.mma.cta_group::1.kind::f16 (K = 4 x 16)
//     issued by one thread; fp32 accumulators in TMEM, double buffered (2 x 128 columns).  A
//     4-MMA chain has ~400 cycles of fixed issue->commit->wake latency on B200 (measured,
//     profiles/microbench/mma_lat.cu), far more than its 256 cycles of math, so TWO CTAs are
//     resident per SM (256 TMEM columns each): four tiles are in flight per SM and the tensor
//     pipe of one CTA works while the other CTA's chain drains;
//   * epilogue (4 warps, thread = user row): with K = 64 every accumulator element is read after
//     only 64 MACs, so the TMEM -> register read (tcgen05.ld, ~170 B/cycle/SM while the MMA pipe
//     is busy) is the co-bottleneck of the MMA pipe.  Loads are software-pipelined (chunk c+1 in
//     flight while chunk c is reduced); a 32-score chunk is rejected with a 3-input-max tree
//     (FMNMX3) and ONE compare against the row's threshold tau;
//   * survivors (about k*ln(m/k) per row over the whole sweep) must not stall the pipeline: a
//     survivor is only APPENDED to the row's unsorted candidate buffer in shared memory after a
//     merge-cursor test against the user's training items (CSR row of R; item ids arrive in
//     ascending order, so the cursor only moves forward: no binary search, no dependent global
//     loads).  When a buffer cannot take the next group of survivors the whole warp compacts it:
//     rank-by-counting over the <= CAP entries (exchanged by warp shuffles), the best k are
//     rewritten in sorted order and tau becomes the k-th best.
//     Scores never touch HBM;
//   * warp roles: warp 0 = bulk-copy producer, warp 1 = TMEM allocator + MMA issuer, warps 2-5 =
//     epilogue (TMEM lane quarter = warp % 4).  The two single-thread roles are entered through
//     elect.sync (straight-line UTCHMMA/UTCBAR/UBLKCP instead of one ELECT + BRA.U.ANY loop per
//     instruction) and the MMA thread issues its two mbarrier polls of a tile back to back:
//     together 1134 -> 1699 TFLOP/s for the MMA side alone, 895 -> 935 with the epilogue.
#include "topk.cuh"
#include <cuda_bf16.h>
#include <stdlib.h>

namespace spex {
namespace tc {

constexpr int BM = 128;            // users per CTA  (UMMA M)
constexpr int BN = 128;            // items per tile (UMMA N)
constexpr int DK = 64;             // embedding dim  (4 x UMMA K)
constexpr int UMMA_K = 16;
constexpr int A_BYTES = BM * DK * 2;
constexpr int B_BYTES = BN * DK * 2;
constexpr int kThreads = 32 * 6;
constexpr int kMaxStages = 4;
constexpr int kTmemCols = 256;     // 2 accumulator buffers x 128 columns; two CTAs share an SM
constexpr int KMAX_TC = 64;
constexpr int kGroup = 8;          // appends are capacity-checked every 8 scores
constexpr uint32_t kSpinLimit = 1u << 26;   // bounded waits: trap instead of hanging the GPU

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > kSpinLimit) __trap();
  }
}
// 1-D bulk async copy global -> shared, completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// Elect one thread of a converged warp.  With elect.sync the compiler knows that exactly one thread
// runs the guarded block and issues tcgen05.mma / tcgen05.commit / cp.async.bulk straight from
// uniform registers; with `lane == 0` it wraps every such instruction in an ELECT + BRA.U.ANY
// loop with R2UR moves (~61 cycles per MMA issue, measured: profiles/microbench/mma_pipe.cu).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "elect.sync _|P1, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> fp32
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                       uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// UMMA shared-memory descriptor, SWIZZLE_NONE, K-major canonical layout
//   bits [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout=0
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// tcgen05.ld 32 lanes x 32 columns (one 32-bit word per lane per column) WITHOUT waiting: the
// registers are only valid after tmem_wait32() on the same array.
#define SPEX_R32(r)                                                                            \
  r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7], r[8], r[9], r[10], r[11], r[12], r[13], r[14], \
      r[15], r[16], r[17], r[18], r[19], r[20], r[21], r[22], r[23], r[24], r[25], r[26], r[27], \
      r[28], r[29], r[30], r[31]
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// wait for every outstanding tcgen05.ld of this thread; the "+r" operands tie the loaded
// registers to the wait so that no consumer can be scheduled above it.
__device__ __forceinline__ void tmem_wait32(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]),
                 "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]),
                 "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]),
                 "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                 "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]),
                 "+r"(r[31])
               :
               : "memory");
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));  // FMNMX3
  return d;
}
// max of 32 scores in 16 FMNMX3
__device__ __forceinline__ float max32(const uint32_t (&r)[32]) {
#define F(i) __uint_as_float(r[i])
  const float a0 = max3(F(0), F(1), F(2)), a1 = max3(F(3), F(4), F(5)), a2 = max3(F(6), F(7), F(8));
  const float a3 = max3(F(9), F(10), F(11)), a4 = max3(F(12), F(13), F(14));
  const float a5 = max3(F(15), F(16), F(17)), a6 = max3(F(18), F(19), F(20));
  const float a7 = max3(F(21), F(22), F(23)), a8 = max3(F(24), F(25), F(26));
  const float a9 = max3(F(27), F(28), F(29));
  const float b0 = max3(a0, a1, a2), b1 = max3(a3, a4, a5), b2 = max3(a6, a7, a8);
  const float b3 = max3(a9, F(30), F(31));
  return fmaxf(max3(b0, b1, b2), b3);
#undef F
}

// Per-row mask cursor, kept in shared memory because only the rare survivor path touches it.
// Items are offered to a row in strictly ascending id order (tile, chunk, column) and the row's
// training items (CSR row of R) are ascending too, so the mask test is a merge: the cursor only
// moves forward and a row performs at most |train(u)| sequential mask loads over the whole sweep.
struct MaskCursor {
  const int32_t* cur[BM];
  const int32_t* end[BM];
  int next[BM];                // smallest training item id not yet passed (INT_MAX: none left)
};

__device__ __forceinline__ unsigned long long pack_state(int n, float tau) {
  return ((unsigned long long)(unsigned)n << 32) | (unsigned long long)__float_as_uint(tau);
}

// One column of a chunk has at least one survivor in this warp (lanes in `b`).  Called by the
// whole warp (convergent), out of line.  Each surviving lane runs the mask cursor and appends
// (v, id) to its row's candidate buffer (row-major in shared memory: entry p of row r at
// cv[r * CAP + p]); rows whose buffer became full are compacted cooperatively: every lane takes
// EPL entries into registers, ranks them by counting the entries that beat them (entries travel
// by warp shuffle; (score desc, id asc) is a strict total order, so ranks are a permutation), the
// best k are written back in sorted order and the row's threshold becomes its k-th best score.
// Returns this lane's (n, tau).  `force` compacts the rows in `b` without appending (final pass).
template <int EPL>
__device__ __noinline__ unsigned long long tc_column(unsigned b, float v, int id, int n, float tau,
                                                     int k, int m_items, int lane, int row, float* cv_warp,
                                                     int* ci_warp, MaskCursor* mc, bool force) {
  constexpr int CAP = 32 * EPL;
  unsigned need = b;
  if (!force) {
    if (((b >> lane) & 1u) && id < m_items) {   // id >= m_items: zero padding of the last tile
      int mnext = mc->next[row];
      if (mnext < id) {
        const int32_t* c = mc->cur[row];
        const int32_t* e = mc->end[row];
        do {
          ++c;
          mnext = (c < e) ? __ldg(c) : 0x7fffffff;
        } while (mnext < id);
        mc->cur[row] = c;
        mc->next[row] = mnext;
      }
      if (mnext != id) {                        // == id: training item of this user, excluded
        cv_warp[lane * CAP + n] = v;
        ci_warp[lane * CAP + n] = id;
        ++n;
      }
    }
    need = __ballot_sync(kFull, n == CAP);
  }
  while (need) {
    const int src = __ffs(need) - 1;
    need &= need - 1;
    const int ns = __shfl_sync(kFull, n, src);
    float* cvr = cv_warp + src * CAP;
    int* cir = ci_warp + src * CAP;
    float ev[EPL];
    int ei[EPL], rank[EPL];
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
      const int p = lane + 32 * e;
      const bool ok = p < ns;
      ev[e] = ok ? cvr[p] : -INFINITY;
      ei[e] = ok ? cir[p] : 0x7fffffff;
      rank[e] = 0;
    }
#pragma unroll
    for (int e2 = 0; e2 < EPL; ++e2) {
#pragma unroll 4
      for (int j = 0; j < 32; ++j) {   // kept rolled: cold code must stay small (i-cache)
        const float vj = __shfl_sync(kFull, ev[e2], j);
        const int ij = __shfl_sync(kFull, ei[e2], j);
#pragma unroll
        for (int e = 0; e < EPL; ++e) rank[e] += beats(vj, ij, ev[e], ei[e]) ? 1 : 0;
      }
    }
    __syncwarp();
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
      if (lane + 32 * e < ns && rank[e] < k) {
        cvr[rank[e]] = ev[e];
        cir[rank[e]] = ei[e];
      }
    }
    __syncwarp();
    const float tau_new = (ns >= k) ? cvr[k - 1] : -INFINITY;
    if (lane == src) {
      n = ns < k ? ns : k;
      tau = tau_new;
    }
  }
  return pack_state(n, tau);
}

__device__ __forceinline__ float tmem_ld1(uint32_t taddr) {
  uint32_t x;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(x) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(x) : : "memory");
  return __uint_as_float(x);
}

// reduce one 32-score chunk against the row's threshold.  Fast path: 16 FMNMX3 + one compare +
// one warp vote.  The kernel is instruction-fetch sensitive (the hot loop must stay resident in
// the SM's instruction cache while two CTAs and three warp roles share it), so the survivor path
// is written as LOOPS, not unrolled code: each lane builds a 32-bit mask of its surviving
// columns, the warp ORs the masks (REDUX) and visits only the columns that have a survivor,
// re-reading that column from TMEM (the accumulator buffer is still owned by the epilogue).
template <int EPL>
__device__ __forceinline__ void consume_chunk(const uint32_t (&r)[32], uint32_t taddr, int item0, int& n,
                                              float& tau, int k, int m_items, int lane, int row,
                                              float* cv_warp, int* ci_warp, MaskCursor* mc) {
  if (__any_sync(kFull, max32(r) > tau)) {
    unsigned mine = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) mine |= (__uint_as_float(r[i]) > tau) ? (1u << i) : 0u;
    unsigned cols = __reduce_or_sync(kFull, mine);
#pragma unroll 1
    while (cols) {
      const int i = __ffs(cols) - 1;
      cols &= cols - 1;
      const float v = tmem_ld1(taddr + i);
      const unsigned b = __ballot_sync(kFull, v > tau);   // tau may have risen since the mask was built
      if (b) {
        const unsigned long long st = tc_column<EPL>(b, v, item0 + i, n, tau, k, m_items, lane, row,
                                                     cv_warp, ci_warp, mc, false);
        tau = __uint_as_float((unsigned)(st & 0xffffffffull));
        n = (int)(st >> 32);
      }
    }
  }
}

template <int EPL>
__global__ void __launch_bounds__(kThreads, 2)
score_topk_tc_kernel(const uint8_t* __restrict__ Ub, const uint8_t* __restrict__ Ib, int64_t B,
                     int m_items, int n_item_tiles, const int64_t* __restrict__ user_ids,
                     const int64_t* __restrict__ mask_rowptr, const int32_t* __restrict__ mask_col,
                     int k, int32_t* __restrict__ out_idx, float* __restrict__ out_val, int stages,
                     long long* __restrict__ trace, int dbg) {
  constexpr int CAP = 32 * EPL;
  // bring-up instrumentation (trace == nullptr in production): CTA 0 records clock64 stamps of
  // tiles [kTraceT0, kTraceT0 + 64): MMA thread [t][5] loop top, [6] B tile landed, [0] TMEM
  // buffer free, [1] chain issued; epilogue warp 2 [t][2] accumulator ready, [3] drained
  // (compiled in only with -DSPEX_TC_TRACE: the stamps lengthen the MMA thread's loop; the same
  // builds read SPEX_TC_DBG: bit 0 = nothing survives, i.e. the cost of the survivor path)
#ifdef SPEX_TC_TRACE
  constexpr int kTraceT0 = 2000;
  const bool tracing = trace != nullptr && blockIdx.x == 0;
#define SPEX_TC_STAMP(t, slot)                                                       \
  do {                                                                               \
    if (tracing && (t) >= kTraceT0 && (t) < kTraceT0 + 64)                           \
      trace[((t) - kTraceT0) * 8 + (slot)] = clock64();                              \
  } while (0)
#else
#define SPEX_TC_STAMP(t, slot) do { } while (0)
#endif
  constexpr uint32_t lbo = 128, sbo = 1024;   // layout written by spex_pack_bf16 (D = 64)
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[kMaxStages];
  __shared__ __align__(8) uint64_t bar_empty[kMaxStages];
  __shared__ __align__(8) uint64_t bar_tfull[2];
  __shared__ __align__(8) uint64_t bar_tempty[2];
  __shared__ __align__(8) uint64_t bar_a;
  __shared__ uint32_t tmem_slot;
  __shared__ MaskCursor mc;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) &
                                             ~(uintptr_t)127);
  uint8_t* sA = smem;                                                    // [128 users][64] bf16
  uint8_t* sB = smem + A_BYTES;                                          // stages x [128 items][64]
  float* cv = reinterpret_cast<float*>(sB + (size_t)stages * B_BYTES);   // [BM][CAP]
  int* ci = reinterpret_cast<int*>(cv + (size_t)CAP * BM);               // [BM][CAP]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t tile_m = blockIdx.x;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&bar_full[s], 1);
      mbar_init(&bar_empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bar_tfull[b], 1);
      mbar_init(&bar_tempty[b], 4);   // one arrive per epilogue warp
    }
    mbar_init(&bar_a, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&tmem_slot)),
                 "r"((uint32_t)kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    // ===== producer: one thread streams the user tile once, then every item tile =====
    if (elect_one()) {
      mbar_arrive_expect_tx(&bar_a, A_BYTES);
      bulk_g2s(sA, Ub + tile_m * A_BYTES, A_BYTES, &bar_a);
      int s = 0;
      uint32_t ph = 0;
      for (int t = 0; t < n_item_tiles; ++t) {
        mbar_wait(&bar_empty[s], ph ^ 1u);
        mbar_arrive_expect_tx(&bar_full[s], B_BYTES);
        bulk_g2s(sB + (size_t)s * B_BYTES, Ib + (size_t)t * B_BYTES, B_BYTES, &bar_full[s]);
        if (++s == stages) {
          s = 0;
          ph ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: one thread drives the tensor core =====
    if (elect_one()) {
      // instruction descriptor: c=f32 (1<<4), a=bf16 (1<<7), b=bf16 (1<<10), K-major A and B,
      // N>>3 at [17,23), M>>4 at [24,29)
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) |
                             ((uint32_t)(BM >> 4) << 24);
      const uint32_t a_addr = smem_u32(sA);
      constexpr uint32_t kstep = 2 * lbo;  // 16 bf16 along K = two 8-element core matrices
      mbar_wait(&bar_a, 0);
      int s = 0;
      uint32_t ph = 0;
      for (int t = 0; t < n_item_tiles; ++t) {
        const int buf = t & 1;
        const uint32_t use = (uint32_t)(t >> 1) & 1u;
        SPEX_TC_STAMP(t, 5);
        // While the tensor pipe is busy a poll of an mbarrier takes ~200 cycles even when its
        // phase is long complete (profiles/microbench/mma_pipe.cu), so the two polls of a tile
        // (operands landed, accumulator drained) are issued back to back and overlap.
        bool have_b = mbar_try_wait(&bar_full[s], ph);
        bool have_d = mbar_try_wait(&bar_tempty[buf], use ^ 1u);
        uint32_t spins = 0;
        while (!have_b) {
          have_b = mbar_try_wait(&bar_full[s], ph);
          if (++spins > kSpinLimit) __trap();
        }
        SPEX_TC_STAMP(t, 6);
        while (!have_d) {
          have_d = mbar_try_wait(&bar_tempty[buf], use ^ 1u);
          if (++spins > kSpinLimit) __trap();
        }
        tc_fence_after();
        SPEX_TC_STAMP(t, 0);
        const uint32_t b_addr = smem_u32(sB + (size_t)s * B_BYTES);
#pragma unroll
        for (int kk = 0; kk < DK / UMMA_K; ++kk) {
          tc_mma(tmem_base + (uint32_t)(buf * BN), make_desc(a_addr + kk * kstep, lbo, sbo),
                 make_desc(b_addr + kk * kstep, lbo, sbo), idesc, kk > 0 ? 1u : 0u);
        }
        tc_commit(&bar_empty[s]);     // smem stage reusable once these MMAs have read it
        tc_commit(&bar_tfull[buf]);   // accumulator ready for the epilogue
        SPEX_TC_STAMP(t, 1);
        if (++s == stages) {
          s = 0;
          ph ^= 1u;
        }
      }
    }
  } else {
    // ===== epilogue: thread owns one user row; warp = TMEM lane quarter =====
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int64_t grow = tile_m * BM + row;
    float* cv_warp = cv + q * 32 * CAP;   // + lane * CAP = this thread's row
    int* ci_warp = ci + q * 32 * CAP;
    int n = 0;
    float tau = (grow < B) ? -INFINITY : INFINITY;   // +inf: padded row, nothing survives
    if (dbg & 1) tau = INFINITY;                     // bring-up builds only
    {
      const int32_t* c = mask_col;
      const int32_t* e = mask_col;
      int mnext = 0x7fffffff;
      if (grow < B && mask_rowptr) {
        const int64_t uid = user_ids ? user_ids[grow] : grow;
        c = mask_col + mask_rowptr[uid];
        e = mask_col + mask_rowptr[uid + 1];
        if (c < e) mnext = __ldg(c);
      }
      mc.cur[row] = c;
      mc.end[row] = e;
      mc.next[row] = mnext;
    }
    uint32_t ra[32], rb[32], rc[32];
    for (int t = 0; t < n_item_tiles; ++t) {
      const int buf = t & 1;
      const uint32_t use = (uint32_t)(t >> 1) & 1u;
      mbar_wait(&bar_tfull[buf], use);
      tc_fence_after();
      if (warp == 2 && lane == 0) SPEX_TC_STAMP(t, 2);
      const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN);
      const int item0 = t * BN;
      // software pipeline over the four 32-column chunks with three register buffers: two
      // tcgen05.ld are in flight while a chunk is reduced (TMEM read latency is ~190 cycles per
      // 4 KB while the MMA pipe is busy, so the read rate scales with the loads in flight)
      tmem_ld32_issue(tbase, ra);
      tmem_ld32_issue(tbase + 32, rb);
      tmem_wait32(ra);
      tmem_wait32(rb);
      tmem_ld32_issue(tbase + 64, rc);
      consume_chunk<EPL>(ra, tbase, item0, n, tau, k, m_items, lane, row, cv_warp, ci_warp, &mc);
      tmem_ld32_issue(tbase + 96, ra);
      consume_chunk<EPL>(rb, tbase + 32, item0 + 32, n, tau, k, m_items, lane, row, cv_warp, ci_warp, &mc);
      tmem_wait32(rc);
      tmem_wait32(ra);
      consume_chunk<EPL>(rc, tbase + 64, item0 + 64, n, tau, k, m_items, lane, row, cv_warp, ci_warp, &mc);
      consume_chunk<EPL>(ra, tbase + 96, item0 + 96, n, tau, k, m_items, lane, row, cv_warp, ci_warp, &mc);
      // every column of this accumulator has been reduced: hand the TMEM buffer back
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_tempty[buf]);
      if (warp == 2 && lane == 0) SPEX_TC_STAMP(t, 3);
    }
    // final compaction of every row (sorted best-first), then each thread writes its own row
    n = (int)(tc_column<EPL>(kFull, 0.f, 0, n, tau, k, m_items, lane, row, cv_warp, ci_warp, &mc, true) >> 32);
    if (grow < B) {
      for (int p = 0; p < k; ++p) {
        const bool ok = p < n;
        out_idx[grow * k + p] = ok ? ci_warp[lane * CAP + p] : -1;
        out_val[grow * k + p] = ok ? cv_warp[lane * CAP + p] : -INFINITY;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)kTmemCols)
                 : "memory");
  }
}

// fp32 [rows, D] -> bf16 core-matrix-tiled layout:
//   byte offset of (r, k) = (r/8) * (16*D) + (k/8) * 128 + (r%8) * 16 + (k%8) * 2
// optional row gather; rows in [n, n_pad) are zero.  Thread per (row, 8-element chunk).
__global__ void __launch_bounds__(256)
pack_bf16_kernel(const float* __restrict__ src, const int64_t* __restrict__ rows, int64_t n,
                 int64_t n_pad, int D, uint8_t* __restrict__ dst) {
  const int D8 = D >> 3;
  const int64_t total = n_pad * D8;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    const int64_t r = i / D8;
    const int kc = (int)(i - r * D8);
    float4 a = f4_zero(), b = f4_zero();
    if (r < n) {
      const int64_t sr = rows ? rows[r] : r;
      a = *reinterpret_cast<const float4*>(src + sr * D + kc * 8);
      b = *reinterpret_cast<const float4*>(src + sr * D + kc * 8 + 4);
    }
    uint4 pk;
    pk.x = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(a.x)) |
           ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(a.y)) << 16);
    pk.y = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(a.z)) |
           ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(a.w)) << 16);
    pk.z = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(b.x)) |
           ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(b.y)) << 16);
    pk.w = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(b.z)) |
           ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(b.w)) << 16);
    const int64_t off = (r >> 3) * (16 * (int64_t)D) + (int64_t)kc * 128 + (r & 7) * 16;
    *reinterpret_cast<uint4*>(dst + off) = pk;
  }
}

}  // namespace tc
}  // namespace spex

using namespace spex;

extern "C" int spex_pack_bf16(const float* src, const int64_t* rows, int64_t n, int64_t n_pad,
                              int32_t D, void* dst_bf16, void* stream) {
  SPEX_RETURN_IF(!src || !dst_bf16 || n < 0 || n_pad < n || (n_pad & 7), SPEX_E_BADARG);
  SPEX_RETURN_IF(D <= 0 || (D & 7) || D > 512, SPEX_E_BADDIM);
  SPEX_RETURN_IF(!aligned16(src) || !aligned16(dst_bf16), SPEX_E_ALIGN);
  if (n_pad == 0) return 0;
  int64_t blocks = (n_pad * (D / 8) + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  tc::pack_bf16_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      src, rows, n, n_pad, D, (uint8_t*)dst_bf16);
  count_launch();
  return check_last();
}

static long long* g_tc_trace = nullptr;
// bring-up only (not part of the declared ABI): device buffer of 64*8 int64 clock stamps
extern "C" void spex_debug_tc_trace(void* dev_buf) { g_tc_trace = (long long*)dev_buf; }

static size_t tc_smem_bytes(int cap, int stages) {
  return 128 + (size_t)tc::A_BYTES + (size_t)stages * tc::B_BYTES + (size_t)cap * tc::BM * 8;
}

template <int EPL>
static int tc_launch(const void* Ub, const void* Ib, int64_t B, int64_t B_pad, int64_t m_items,
                     int64_t m_pad, const int64_t* user_ids, const int64_t* mask_rowptr,
                     const int32_t* mask_col, int32_t k, int32_t* out_idx, float* out_val,
                     cudaStream_t st) {
  // two CTAs per SM need <= ~113 KB each (228 KB per SM, 1 KB reserved per CTA, static smem)
  int stages = tc::kMaxStages;
  while (stages > 2 && tc_smem_bytes(32 * EPL, stages) + 2800 > 115712) --stages;
  if (const char* e = getenv("SPEX_TC_STAGES")) stages = atoi(e);   // bring-up experiment
  const size_t smem = tc_smem_bytes(32 * EPL, stages);
  SPEX_RETURN_IF(smem > 226 * 1024, SPEX_E_TOOBIG);
  // per device and per function, and the size depends on `stages`: set on every launch (cheap)
  {
    cudaError_t e = cudaFuncSetAttribute(tc::score_topk_tc_kernel<EPL>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  const int64_t grid = B_pad / tc::BM;
  SPEX_RETURN_IF(grid > 0x7fffffffLL, SPEX_E_TOOBIG);
  int dbg = 0;
#ifdef SPEX_TC_TRACE
  if (const char* e = getenv("SPEX_TC_DBG")) dbg = atoi(e);
#endif
  tc::score_topk_tc_kernel<EPL><<<(unsigned)grid, tc::kThreads, smem, st>>>(
      (const uint8_t*)Ub, (const uint8_t*)Ib, B, (int)m_items, (int)(m_pad / tc::BN), user_ids,
      mask_rowptr, mask_col, k, out_idx, out_val, stages, g_tc_trace, dbg);
  count_launch();
  return check_last();
}

extern "C" int spex_score_topk_bf16(const void* Ub, const void* Ib, int64_t B, int64_t B_pad,
                                    int64_t m_items, int64_t m_pad, const int64_t* user_ids,
                                    const int64_t* mask_rowptr, const int32_t* mask_col, int32_t k,
                                    int32_t* out_idx, float* out_val, void* stream) {
  SPEX_RETURN_IF(!Ub || !Ib || !out_idx || !out_val || B < 0 || B_pad < B || m_items < 0 ||
                     m_pad < m_items,
                 SPEX_E_BADARG);
  SPEX_RETURN_IF((mask_rowptr == nullptr) != (mask_col == nullptr), SPEX_E_BADARG);
  SPEX_RETURN_IF((B_pad % tc::BM) || (m_pad % tc::BN), SPEX_E_BADARG);
  SPEX_RETURN_IF(k < 1 || k > tc::KMAX_TC || m_pad > 0x7fffffffLL, SPEX_E_TOOBIG);
  SPEX_RETURN_IF(!aligned16(Ub) || !aligned16(Ib), SPEX_E_ALIGN);
  if (B == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  // candidate buffer capacity CAP = 32*EPL must hold k kept entries + one group of appends
  if (k <= 32 - tc::kGroup)
    return tc_launch<1>(Ub, Ib, B, B_pad, m_items, m_pad, user_ids, mask_rowptr, mask_col, k, out_idx, out_val, st);
  if (k <= 64 - tc::kGroup)
    return tc_launch<2>(Ub, Ib, B, B_pad, m_items, m_pad, user_ids, mask_rowptr, mask_col, k, out_idx, out_val, st);
  return tc_launch<3>(Ub, Ib, B, B_pad, m_items, m_pad, user_ids, mask_rowptr, mask_col, k, out_idx, out_val, st);
}
