// f16acc.cu — can the K=64 scorer halve its TMEM drain with fp16 accumulators?  (B200, sm_100a)
//
// The scorer (spex_b200/csrc/score_topk_tc.cu) reads every fp32 accumulator element after only 64
// MACs; r01 measured ~190 cycles per tcgen05.ld.32x32b.x32 (4 KB) per warp while the MMA pipe is
// busy, i.e. the TMEM -> register read co-bounds the kernel.  tcgen05.mma.kind::f16 can also
// accumulate in fp16 (instruction descriptor c_format = 0); the accumulator still occupies one
// 32-bit TMEM cell per element, but tcgen05.ld ... .pack::16b packs two adjacent columns into one
// register.  This microbenchmark answers, with numbers:
//   (1) is the fp16-accumulator result of a 128x128x64 tile right (fp16 operands; bf16 operands
//       with `bf16` on the command line - that combination may be illegal, so it is separate)?
//   (2) what does a tcgen05.ld cost per warp, alone and under MMA load, for
//         x32 (32 columns -> 32 regs), x32.pack::16b (64 columns -> 32 regs),
//         x64 (64 columns -> 64 regs), x64.pack::16b (128 columns -> 64 regs)?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f16acc f16acc.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) { if (++spins > (1u << 24)) __trap(); }
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}

#define R32(r, o)                                                                                         \
  "=r"(r[o + 0]), "=r"(r[o + 1]), "=r"(r[o + 2]), "=r"(r[o + 3]), "=r"(r[o + 4]), "=r"(r[o + 5]),          \
      "=r"(r[o + 6]), "=r"(r[o + 7]), "=r"(r[o + 8]), "=r"(r[o + 9]), "=r"(r[o + 10]), "=r"(r[o + 11]),    \
      "=r"(r[o + 12]), "=r"(r[o + 13]), "=r"(r[o + 14]), "=r"(r[o + 15]), "=r"(r[o + 16]), "=r"(r[o + 17]), \
      "=r"(r[o + 18]), "=r"(r[o + 19]), "=r"(r[o + 20]), "=r"(r[o + 21]), "=r"(r[o + 22]), "=r"(r[o + 23]), \
      "=r"(r[o + 24]), "=r"(r[o + 25]), "=r"(r[o + 26]), "=r"(r[o + 27]), "=r"(r[o + 28]), "=r"(r[o + 29]), \
      "=r"(r[o + 30]), "=r"(r[o + 31])
#define L32A "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
#define L64A "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31," \
             "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"

// variant 0: x32   1: x32.pack::16b   2: x64   3: x64.pack::16b
template <int V>
__device__ __forceinline__ uint32_t ld_variant(uint32_t taddr) {
  uint32_t x = 0;
  if constexpr (V == 0 || V == 1) {
    uint32_t r[32];
    if constexpr (V == 0)
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 " L32A : R32(r, 0) : "r"(taddr) : "memory");
    else
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 " L32A : R32(r, 0) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) x ^= r[i];
  } else {
    uint32_t r[64];
    if constexpr (V == 2)
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 " L64A : R32(r, 0), R32(r, 32) : "r"(taddr) : "memory");
    else
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.pack::16b.b32 " L64A : R32(r, 0), R32(r, 32) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 64; ++i) x ^= r[i];
  }
  return x;
}

// ---- (1) correctness -------------------------------------------------------------------------------
// one CTA: 128x128x64 tile, operands written by the kernel in the canonical no-swizzle K-major layout
// (byte offset of (r,k) = (r/8)*1024 + (k/8)*128 + (r%8)*16 + (k%8)*2), accumulator read back by the
// four warps (x32 loads, raw words) and with pack::16b.
__global__ void __launch_bounds__(128, 1)
check_kernel(const uint16_t* __restrict__ A, const uint16_t* __restrict__ B, uint32_t idesc,
             uint32_t* __restrict__ raw, uint32_t* __restrict__ packed) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 128 * 64; i += blockDim.x) {
    const int r = i / 64, k = i % 64;
    const int off = (r / 8) * 1024 + (k / 8) * 128 + (r % 8) * 16 + (k % 8) * 2;
    *reinterpret_cast<uint16_t*>(smem + off) = A[i];
    *reinterpret_cast<uint16_t*>(smem + 16384 + off) = B[i];
  }
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 16384);
    for (int kk = 0; kk < 4; ++kk)
      tc_mma(tmem, make_desc(a_addr + kk * 256, 128, 1024), make_desc(b_addr + kk * 256, 128, 1024), idesc, kk > 0);
    tc_commit(&bar);
  }
  mbar_wait(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tmem + ((uint32_t)(warp * 32) << 16);
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < 128; c0 += 32) {
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 " L32A : R32(r, 0) : "r"(base + c0) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 32; ++i) raw[row * 128 + c0 + i] = r[i];
  }
  for (int c0 = 0; c0 < 128; c0 += 64) {
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 " L32A : R32(r, 0) : "r"(base + c0) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 32; ++i) packed[row * 64 + c0 / 2 + i] = r[i];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
  }
}

// ---- (2) timing ------------------------------------------------------------------------------------
// warp 0 lane 0 keeps the tensor core busy (8 MMAs of N=128 per batch = two tiles into columns
// [0,128) and [128,256)); warps 2.. read columns [256, 256+128) with one of the load variants.
template <int V>
__global__ void __launch_bounds__(320, 1)
rate_kernel(int with_mma, uint32_t idesc, int iters, int n_readers, long long* cyc, uint32_t* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 32768 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    stop = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 16384);
  uint32_t x = 0;
  if (warp == 0 && lane == 0 && with_mma) {
    long long t0 = clock64();
    uint32_t ph = 0;
    for (int it = 0; it < iters; ++it) {
      for (int m = 0; m < 8; ++m)
        tc_mma(tmem + (m >= 4 ? 128 : 0), make_desc(a_addr + (m & 3) * 256, 128, 1024),
               make_desc(b_addr + (m & 3) * 256, 128, 1024), idesc, (m & 3) ? 1u : 0u);
      tc_commit(&bar);
      mbar_wait(&bar, ph);
      ph ^= 1u;
    }
    long long t1 = clock64();
    cyc[blockIdx.x * 2] = t1 - t0;
    stop = 1;
  }
  if (warp >= 2 && warp < 2 + n_readers) {
    // two readers per lane quarter take different column ranges
    const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 256 + (((warp - 2) >> 2) & 1) * 128;
    long long t0 = clock64();
    int n = 0;
    if (with_mma) {
      while (!stop) { x ^= ld_variant<V>(base); ++n; }
    } else {
      for (; n < iters; ++n) x ^= ld_variant<V>(base);
    }
    long long t1 = clock64();
    if (warp == 2 && lane == 0) {
      cyc[blockIdx.x * 2 + 1] = (t1 - t0);
      sink[blockIdx.x] = n;
    }
  }
  if (x == 0x12345) sink[1000 + threadIdx.x] = x;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

static uint32_t make_idesc(int c_fmt, int ab_fmt, int N) {
  // c_format [4,6): 0 f16, 1 f32; a_format [7,10), b_format [10,13): 0 f16, 1 bf16; N>>3 at [17,23), M>>4 at [24,29)
  return ((uint32_t)c_fmt << 4) | ((uint32_t)ab_fmt << 7) | ((uint32_t)ab_fmt << 10) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(128 >> 4) << 24);
}

template <int V>
static void run_rate(const char* name, int cols, int c_fmt, long long* cyc, uint32_t* sink) {
  const int iters = 20000;
  long long h[2];
  uint32_t hn;
  const uint32_t idesc = make_idesc(c_fmt, 0, 128);
  for (int readers : {4, 8}) {
    rate_kernel<V><<<148, 320, 32768 + 1024>>>(0, idesc, iters, readers, cyc, sink);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost);
    const double alone = (double)h[1] / iters;
    rate_kernel<V><<<148, 320, 32768 + 1024>>>(1, idesc, iters, readers, cyc, sink);
    cudaError_t e2 = cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost);
    cudaMemcpy(&hn, sink, 4, cudaMemcpyDeviceToHost);
    const double busy = (double)h[1] / (hn ? hn : 1);
    printf("%-22s D=%s readers=%d : %6.1f cycles/ld alone (%5.2f cyc/column), %6.1f under MMA (%5.2f cyc/column); "
           "issuer %7.1f cycles per 8-MMA batch (ideal 512)  (%s, %s)\n",
           name, c_fmt ? "f32" : "f16", readers, alone, alone / cols, busy, busy / cols, (double)h[0] / iters,
           cudaGetErrorString(e), cudaGetErrorString(e2));
  }
}

int main(int argc, char** argv) {
  const bool bf16_in = argc > 1 && !strcmp(argv[1], "bf16");
  // ---- (1) ----
  uint16_t hA[128 * 64], hB[128 * 64];
  static float fA[128 * 64], fB[128 * 64];
  for (int r = 0; r < 128; ++r)
    for (int k = 0; k < 64; ++k) {
      const float a = (float)((r * 7 + k * 3) % 17 - 8) / 8.f + (float)(k % 5) / 64.f;
      const float b = (float)((r * 5 + k) % 13 - 6) / 4.f + (float)(r % 3) / 32.f;
      if (bf16_in) {
        __nv_bfloat16 xa = __float2bfloat16(a), xb = __float2bfloat16(b);
        memcpy(&hA[r * 64 + k], &xa, 2);
        memcpy(&hB[r * 64 + k], &xb, 2);
        fA[r * 64 + k] = __bfloat162float(xa);
        fB[r * 64 + k] = __bfloat162float(xb);
      } else {
        __half xa = __float2half(a), xb = __float2half(b);
        memcpy(&hA[r * 64 + k], &xa, 2);
        memcpy(&hB[r * 64 + k], &xb, 2);
        fA[r * 64 + k] = __half2float(xa);
        fB[r * 64 + k] = __half2float(xb);
      }
    }
  uint16_t *dA, *dB;
  uint32_t *draw, *dpk;
  cudaMalloc(&dA, sizeof(hA));
  cudaMalloc(&dB, sizeof(hB));
  cudaMalloc(&draw, 128 * 128 * 4);
  cudaMalloc(&dpk, 128 * 64 * 4);
  cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(check_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 + 1024);
  static uint32_t hraw[128 * 128], hpk[128 * 64];
  for (int c_fmt : {1, 0}) {
    cudaMemset(draw, 0xff, 128 * 128 * 4);
    cudaMemset(dpk, 0xff, 128 * 64 * 4);
    check_kernel<<<1, 128, 32768 + 1024>>>(dA, dB, make_idesc(c_fmt, bf16_in ? 1 : 0, 128), draw, dpk);
    cudaError_t e = cudaDeviceSynchronize();
    printf("check: operands %s, accumulator %s: %s\n", bf16_in ? "bf16" : "f16", c_fmt ? "f32" : "f16", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    cudaMemcpy(hraw, draw, sizeof(hraw), cudaMemcpyDeviceToHost);
    cudaMemcpy(hpk, dpk, sizeof(hpk), cudaMemcpyDeviceToHost);
    double max_err = 0, max_err_pk = 0, max_ref = 0;
    int upper_nonzero = 0;
    for (int r = 0; r < 128; ++r)
      for (int c = 0; c < 128; ++c) {
        double ref = 0;
        for (int k = 0; k < 64; ++k) ref += (double)fA[r * 64 + k] * fB[c * 64 + k];
        max_ref = fmax(max_ref, fabs(ref));
        const uint32_t w = hraw[r * 128 + c];
        float got;
        if (c_fmt) {
          memcpy(&got, &w, 4);
        } else {
          __half hh;
          uint16_t lo = (uint16_t)(w & 0xffff);
          memcpy(&hh, &lo, 2);
          got = __half2float(hh);
          if (w >> 16) ++upper_nonzero;
        }
        max_err = fmax(max_err, fabs(got - ref));
        // packed: register i of a 64-column load holds columns (2i, 2i+1) as (lo, hi)?
        const uint32_t pw = hpk[r * 64 + c / 2];
        uint16_t ph = (c & 1) ? (uint16_t)(pw >> 16) : (uint16_t)(pw & 0xffff);
        __half hh2;
        memcpy(&hh2, &ph, 2);
        max_err_pk = fmax(max_err_pk, fabs(__half2float(hh2) - ref));
      }
    printf("  max |ref| %.3f  max abs err (x32 raw) %.5f", max_ref, max_err);
    if (!c_fmt)
      printf("  upper halves non-zero: %d of 16384; max abs err via pack::16b (lo = even column) %.5f", upper_nonzero, max_err_pk);
    printf("\n  raw row 0, columns 0..3: %08x %08x %08x %08x   packed regs 0..1: %08x %08x\n", hraw[0], hraw[1], hraw[2],
           hraw[3], hpk[0], hpk[1]);
  }
  if (bf16_in) return 0;
  // ---- (2) ----
  long long* cyc;
  uint32_t* sink;
  cudaMalloc(&cyc, 148 * 16);
  cudaMalloc(&sink, 8192);
  cudaFuncSetAttribute(rate_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 + 1024);
  cudaFuncSetAttribute(rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 + 1024);
  cudaFuncSetAttribute(rate_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 + 1024);
  cudaFuncSetAttribute(rate_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 + 1024);
  for (int c_fmt : {1, 0}) {
    run_rate<0>("x32 (32 col, 32 reg)", 32, c_fmt, cyc, sink);
    run_rate<1>("x32.pack16 (64 col)", 64, c_fmt, cyc, sink);
    run_rate<2>("x64 (64 col, 64 reg)", 64, c_fmt, cyc, sink);
    run_rate<3>("x64.pack16 (128 col)", 128, c_fmt, cyc, sink);
  }
  return 0;
}
