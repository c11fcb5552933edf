// mma_lat.cu — latency of a short tcgen05.mma chain + commit -> mbarrier -> waiting thread, and the
// tcgen05.ld rate while the tensor core is busy (B200, sm_100a).  Explains the per-tile hand-off
// cost of the K=64 scoring kernel (4 MMAs per accumulator per tile).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

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
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) { while (!mbar_try_wait(bar, parity)) {} }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
__device__ __forceinline__ uint32_t ld32(uint32_t taddr) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  uint32_t x = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) x ^= r[i];
  return x;
}

// mode 0: chain latency (issue n_mma MMAs + commit, wait, repeat)
// mode 1: readers (warps 2..2+n_readers) run tcgen05.ld on buffer 1 while the issuer keeps the
//         tensor core busy on buffer 0; mode 2: readers alone
__global__ void __launch_bounds__(320, 1) k(int mode, int n_mma, int N, int iters, int n_readers,
                                          long long* cyc, uint32_t* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];   // A 32 KB | B 32 KB (zeros: values are irrelevant)
  __shared__ __align__(8) uint64_t bar;
  __shared__ __align__(8) uint64_t bar2;
  __shared__ uint32_t slot;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    mbar_init(&bar2, 1);
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
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 32768);
  uint32_t x = 0;
  if (warp == 0 && lane == 0 && mode != 2) {
    long long t0 = clock64();
    uint32_t ph = 0;
    for (int it = 0; it < iters; ++it) {
      for (int m = 0; m < n_mma; ++m)
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
  if (mode >= 1 && warp >= 2 && warp < 2 + n_readers) {
    const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 256 + ((warp - 2) >> 2) * 64;
    long long t0 = clock64();
    int n = 0;
    if (mode == 1) {
      while (!stop) { x ^= ld32(base); ++n; }
    } else {
      for (; n < iters; ++n) x ^= ld32(base);
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

int main() {
  long long* cyc;
  uint32_t* sink;
  cudaMalloc(&cyc, 148 * 16);
  cudaMalloc(&sink, 8192);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 1024);
  const int iters = 20000;
  long long h[2];
  uint32_t hn;
  for (int N : {64, 128, 256}) {
    for (int n_mma : {1, 4, 8}) {
      k<<<148, 320, 65536 + 1024>>>(0, n_mma, N, iters, 0, cyc, sink);
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost);
      printf("chain: N=%3d n_mma=%d  %8.1f cycles per (issue+commit+wait)   ideal MMA time %4d   (%s)\n", N, n_mma,
             (double)h[0] / iters, n_mma * N / 2, cudaGetErrorString(e));
    }
  }
  for (int readers : {4, 8}) {
    k<<<148, 320, 65536 + 1024>>>(2, 0, 128, iters, readers, cyc, sink);
    cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost);
    printf("ld alone:      readers=%d  %7.1f cycles per 4 KB ld per warp\n", readers, (double)h[1] / iters);
    k<<<148, 320, 65536 + 1024>>>(1, 8, 128, iters, readers, cyc, sink);
    cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost);
    cudaMemcpy(&hn, sink, 4, cudaMemcpyDeviceToHost);
    printf("ld under MMA:  readers=%d  %7.1f cycles per 4 KB ld per warp   (issuer: %.1f cycles per 8-MMA batch)\n",
           readers, (double)h[1] / (hn ? hn : 1), (double)h[0] / iters);
  }
  return 0;
}
