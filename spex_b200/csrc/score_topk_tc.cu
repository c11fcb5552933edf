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
//   * math: tcgen05.mma.cta_group::1.kind::f16, M=128 users x N=256 items x K=16, four per tile
//     (D = 64), issued by one thread; fp32 accumulators in TMEM, double buffered (2 x 256 cols);
//   * epilogue (8 warps): tcgen05.ld 32 columns at a time; a thread owns one user row, keeps the
//     row's running k-th score in a register and rejects a 32-score batch with one max-tree and one
//     compare.  Survivors are checked against the user's training items (binary search in the CSR
//     row of R) and inserted into the row's sorted list in shared memory.  Scores never touch HBM;
//   * warp roles: warp 0 = bulk-copy producer, warp 1 = TMEM allocator + MMA issuer, warps 2-9 =
//     epilogue (TMEM lane quarter = warp % 4, column half = (warp - 2) / 4).
#include "topk.cuh"
#include <cuda_bf16.h>
#include <stdlib.h>

namespace spex {
namespace tc {

constexpr int BM = 128;            // users per CTA  (UMMA M)
constexpr int BN = 256;            // items per tile (UMMA N)
constexpr int DK = 64;             // embedding dim  (4 x UMMA K)
constexpr int UMMA_K = 16;
constexpr int A_BYTES = BM * DK * 2;
constexpr int B_BYTES = BN * DK * 2;
constexpr int kThreads = 32 * 10;
constexpr int kMaxStages = 4;
constexpr int KMAX_TC = 64;
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
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
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
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// Offer one surviving score to a thread-owned sorted list (position-major in shared memory:
// entry p of this thread lives at lv[p * BM], li[p * BM]).  Rare path, kept out of line; the
// caller's (tau, n) stay in registers: the new pair is returned packed as (n << 32) | bits(tau).
__device__ __noinline__ unsigned long long tc_offer(float tau, int n, int64_t mlo, int64_t mhi,
                                                    float v, int id, int k, int m_items, float* lv,
                                                    int* li, const int32_t* __restrict__ mask_col) {
  if (id < m_items && !mask_contains(mask_col, mlo, mhi, id)) {
    int p = (n < k) ? n : k - 1;
    while (p > 0 && beats(v, id, lv[(p - 1) * BM], li[(p - 1) * BM])) {
      lv[p * BM] = lv[(p - 1) * BM];
      li[p * BM] = li[(p - 1) * BM];
      --p;
    }
    lv[p * BM] = v;
    li[p * BM] = id;
    if (n < k) ++n;
    if (n == k) tau = lv[(k - 1) * BM];
  }
  return ((unsigned long long)(unsigned)n << 32) | (unsigned long long)__float_as_uint(tau);
}

__global__ void __launch_bounds__(kThreads, 1)
score_topk_tc_kernel(const uint8_t* __restrict__ Ub, const uint8_t* __restrict__ Ib, int64_t B,
                     int m_items, int n_item_tiles, const int64_t* __restrict__ user_ids,
                     const int64_t* __restrict__ mask_rowptr, const int32_t* __restrict__ mask_col,
                     int k, int32_t* __restrict__ out_idx, float* __restrict__ out_val, int stages,
                     uint32_t lbo, uint32_t sbo) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[kMaxStages];
  __shared__ __align__(8) uint64_t bar_empty[kMaxStages];
  __shared__ __align__(8) uint64_t bar_tfull[2];
  __shared__ __align__(8) uint64_t bar_tempty[2];
  __shared__ __align__(8) uint64_t bar_a;
  __shared__ uint32_t tmem_slot;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + A_BYTES;
  float* lv = reinterpret_cast<float*>(sB + (size_t)stages * B_BYTES);  // [2][k][BM]
  int* li = reinterpret_cast<int*>(lv + 2 * k * BM);                    // [2][k][BM]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t tile_m = blockIdx.x;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&bar_full[s], 1);
      mbar_init(&bar_empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bar_tfull[b], 1);
      mbar_init(&bar_tempty[b], 8);   // one arrive per epilogue warp
    }
    mbar_init(&bar_a, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&tmem_slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    // ===== producer: one lane streams the user tile once, then every item tile =====
    if (lane == 0) {
      mbar_arrive_expect_tx(&bar_a, A_BYTES);
      bulk_g2s(sA, Ub + tile_m * A_BYTES, A_BYTES, &bar_a);
      for (int t = 0; t < n_item_tiles; ++t) {
        const int s = t % stages;
        const uint32_t ph = (uint32_t)(t / stages) & 1u;
        mbar_wait(&bar_empty[s], ph ^ 1u);
        mbar_arrive_expect_tx(&bar_full[s], B_BYTES);
        bulk_g2s(sB + (size_t)s * B_BYTES, Ib + (size_t)t * B_BYTES, B_BYTES, &bar_full[s]);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: one thread drives the tensor core =====
    if (lane == 0) {
      // instruction descriptor: c=f32 (1<<4), a=bf16 (1<<7), b=bf16 (1<<10), K-major A and B,
      // N>>3 at [17,23), M>>4 at [24,29)
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) |
                             ((uint32_t)(BM >> 4) << 24);
      const uint32_t a_addr = smem_u32(sA);
      const uint32_t kstep = 2 * lbo;  // 16 bf16 along K = two 8-element core matrices
      mbar_wait(&bar_a, 0);
      for (int t = 0; t < n_item_tiles; ++t) {
        const int s = t % stages;
        const uint32_t ph = (uint32_t)(t / stages) & 1u;
        const int buf = t & 1;
        const uint32_t use = (uint32_t)(t >> 1) & 1u;
        mbar_wait(&bar_tempty[buf], use ^ 1u);
        mbar_wait(&bar_full[s], ph);
        tc_fence_after();
        const uint32_t b_addr = smem_u32(sB + (size_t)s * B_BYTES);
#pragma unroll
        for (int kk = 0; kk < DK / UMMA_K; ++kk) {
          tc_mma(tmem_base + (uint32_t)buf * BN, make_desc(a_addr + kk * kstep, lbo, sbo),
                 make_desc(b_addr + kk * kstep, lbo, sbo), idesc, kk > 0 ? 1u : 0u);
        }
        tc_commit(&bar_empty[s]);     // smem stage reusable once these MMAs have read it
        tc_commit(&bar_tfull[buf]);   // accumulator ready for the epilogue
      }
    }
  } else {
    // ===== epilogue: thread owns (row, column half) =====
    const int q = warp & 3;             // TMEM lane quarter this warp may access
    const int h = (warp - 2) >> 2;      // column half
    const int row = q * 32 + lane;
    const int64_t grow = tile_m * BM + row;
    float* mylv = lv + (size_t)h * k * BM + row;
    int* myli = li + (size_t)h * k * BM + row;
    int n = 0;
    int64_t mlo = 0, mhi = 0;
    float tau = (grow < B) ? -INFINITY : INFINITY;   // +inf: padded row, nothing survives
    if (grow < B && mask_rowptr) {
      const int64_t uid = user_ids ? user_ids[grow] : grow;
      mlo = mask_rowptr[uid];
      mhi = mask_rowptr[uid + 1];
    }
    for (int t = 0; t < n_item_tiles; ++t) {
      const int buf = t & 1;
      const uint32_t use = (uint32_t)(t >> 1) & 1u;
      mbar_wait(&bar_tfull[buf], use);
      tc_fence_after();
      const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + h * (BN / 2));
      const int item0 = t * BN + h * (BN / 2);
#pragma unroll 1
      for (int c = 0; c < (BN / 2) / 32; ++c) {
        float v[32];
        __syncwarp();
        tmem_ld32(tbase + c * 32, v);
        float m0 = fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3]));
        float m1 = fmaxf(fmaxf(v[4], v[5]), fmaxf(v[6], v[7]));
        float m2 = fmaxf(fmaxf(v[8], v[9]), fmaxf(v[10], v[11]));
        float m3 = fmaxf(fmaxf(v[12], v[13]), fmaxf(v[14], v[15]));
        float m4 = fmaxf(fmaxf(v[16], v[17]), fmaxf(v[18], v[19]));
        float m5 = fmaxf(fmaxf(v[20], v[21]), fmaxf(v[22], v[23]));
        float m6 = fmaxf(fmaxf(v[24], v[25]), fmaxf(v[26], v[27]));
        float m7 = fmaxf(fmaxf(v[28], v[29]), fmaxf(v[30], v[31]));
        const float mx = fmaxf(fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)), fmaxf(fmaxf(m4, m5), fmaxf(m6, m7)));
        if (mx > tau) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (v[i] > tau) {
              const unsigned long long r = tc_offer(tau, n, mlo, mhi, v[i], item0 + c * 32 + i, k,
                                                    m_items, mylv, myli, mask_col);
              tau = __uint_as_float((unsigned)(r & 0xffffffffull));
              n = (int)(r >> 32);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_tempty[buf]);
    }
    // publish list lengths, then the h == 0 thread of each row merges the two halves
    __shared__ int nlist[2][BM];
    nlist[h][row] = n;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (h == 0 && grow < B) {
      const float* av = lv + row;
      const int* ai = li + row;
      const float* bv = lv + (size_t)k * BM + row;
      const int* bi = li + (size_t)k * BM + row;
      const int na = nlist[0][row], nb = nlist[1][row];
      int pa = 0, pb = 0;
      for (int p = 0; p < k; ++p) {
        int id = -1;
        float val = -INFINITY;
        const bool ha = pa < na, hb = pb < nb;
        if (ha && (!hb || beats(av[pa * BM], ai[pa * BM], bv[pb * BM], bi[pb * BM]))) {
          val = av[pa * BM];
          id = ai[pa * BM];
          ++pa;
        } else if (hb) {
          val = bv[pb * BM];
          id = bi[pb * BM];
          ++pb;
        }
        out_idx[grow * k + p] = id;
        out_val[grow * k + p] = val;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u)
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

static size_t tc_smem_bytes(int k, int stages) {
  return 1024 + (size_t)tc::A_BYTES + (size_t)stages * tc::B_BYTES + (size_t)2 * k * tc::BM * 8;
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
  int stages = tc::kMaxStages;
  while (stages > 2 && tc_smem_bytes(k, stages) > 220 * 1024) --stages;
  const size_t smem = tc_smem_bytes(k, stages);
  SPEX_RETURN_IF(smem > 227 * 1024, SPEX_E_TOOBIG);
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(tc::score_topk_tc_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    configured = smem;
  }
  // K-major no-swizzle canonical layout as packed by spex_pack_bf16 (D = 64):
  //   LBO = 128 B between core matrices adjacent in K, SBO = 1024 B between 8-row groups.
  uint32_t lbo = 128, sbo = 1024;
  if (const char* e = getenv("SPEX_TC_LBO")) lbo = (uint32_t)atoi(e);   // bring-up overrides only
  if (const char* e = getenv("SPEX_TC_SBO")) sbo = (uint32_t)atoi(e);
  const int64_t grid = B_pad / tc::BM;
  SPEX_RETURN_IF(grid > 0x7fffffffLL, SPEX_E_TOOBIG);
  tc::score_topk_tc_kernel<<<(unsigned)grid, tc::kThreads, smem, (cudaStream_t)stream>>>(
      (const uint8_t*)Ub, (const uint8_t*)Ib, B, (int)m_items, (int)(m_pad / tc::BN), user_ids,
      mask_rowptr, mask_col, k, out_idx, out_val, stages, lbo, sbo);
  count_launch();
  return check_last();
}
