// mma_pipe.cu — sustained tcgen05.mma throughput of the K=64 scoring shape on B200 (sm_100a):
// chains of 4 MMAs (M=128, N, K=16) into a ring of `nbuf` TMEM accumulators, one commit per chain,
// the issuer only waits for the chain that used the same accumulator `nbuf` chains ago.  No
// epilogue, no global traffic: this is the ceiling of the MMA side of score_topk_tc_kernel for a
// given tile shape / CTAs per SM, in cycles per chain and in TFLOP/s by wall clock (power capping
// included).  Operands are random bf16 (realistic switching power).
#include <cstdio>
#include <cstdint>
#include <cstdlib>
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
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// elect.sync: the compiler then knows exactly one thread runs the block and issues every tcgen05
// instruction straight from uniform registers (with `lane == 0` it wraps each one in an ELECT /
// BRA.U.ANY loop plus R2UR moves: ~61 cycles per MMA issue instead of a few).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
  return pred != 0;
}
// A operand read from TMEM (row = lane, two bf16 per 32-bit column): only B comes from smem
__device__ __forceinline__ void tc_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}

// probe: 0 = pipelined chains; 1 = cost of mbarrier.try_wait on a completed phase; 2 = clock64 pair
__global__ void __launch_bounds__(320) k(int N, int nbuf, int iters, int nstage, long long* cyc, long long* stamps, int pollers, int ts, int alloc_cols, int by_mode) {
  extern __shared__ __align__(1024) uint8_t smem[];   // A 16 KB | nstage x B (N x 128 B)
  __shared__ __align__(8) uint64_t bar[8];
  __shared__ uint32_t slot;
  __shared__ volatile int stop;
  __shared__ volatile uint32_t chase[4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) stop = 0;
  const int bbytes = N * 128;
  uint32_t seed = blockIdx.x * 7919u + threadIdx.x;
  for (int i = threadIdx.x; i < (16384 + nstage * bbytes) / 4; i += blockDim.x) {
    seed = seed * 1664525u + 1013904223u;
    // two bf16 with exponent 0x7e/0x7d region: |x| in [0.25, 1)
    reinterpret_cast<uint32_t*>(smem)[i] = (seed & 0x807f807fu) | 0x3e803e80u;
  }
  if (threadIdx.x == 0) {
    for (int b = 0; b < 8; ++b) mbar_init(&bar[b], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const uint32_t cols = (uint32_t)alloc_cols;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t a_addr = smem_u32(smem);
  if (warp == 0) {
    if (elect_one()) {
      long long t0 = clock64();
      int st = 0;
      const int lg = nbuf == 1 ? 0 : nbuf == 2 ? 1 : 2;
      for (int it = 0; it < iters; ++it) {
        const int buf = it & (nbuf - 1);
        if (it >= nbuf) mbar_wait(&bar[buf], (uint32_t)((it >> lg) - 1) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t b_addr = smem_u32(smem + 16384 + st * bbytes);
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          if (ts)
            tc_mma_ts(tmem + buf * N, tmem + nbuf * N + m * 8, make_desc(b_addr + m * 256, 128, 1024), idesc, m ? 1u : 0u);
          else
            tc_mma(tmem + buf * N, make_desc(a_addr + m * 256, 128, 1024), make_desc(b_addr + m * 256, 128, 1024),
                   idesc, m ? 1u : 0u);
        }
        tc_commit(&bar[buf]);
        if (++st == nstage) st = 0;
      }
      for (int b = 0; b < nbuf; ++b) {   // drain
        const int last = iters - 1 - ((iters - 1 - b) % nbuf);   // last iteration that used buffer b
        if (last >= 0) mbar_wait(&bar[b], (uint32_t)(last / nbuf) & 1u);
      }
      long long t1 = clock64();
      cyc[blockIdx.x] = t1 - t0;
      stop = 1;
    }
  } else if (warp == 1) {
    // a bystander thread while the tensor pipe is busy: what "free" operations cost the MMA /
    // producer threads of the real kernel.  by_mode 0: try_wait on a completed phase; 1: two
    // independent try_waits issued back to back; 2: dependent integer adds (64 per iteration);
    // 3: dependent ld.shared chain (8 per iteration)
    if (lane == 0) {
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar[7])) : "memory");
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar[5])) : "memory");
      chase[0] = 0;
      long long t0 = clock64();
      int n = 0;
      uint32_t x = threadIdx.x;
      while (!stop) {
        if (by_mode == 0) {
          mbar_wait(&bar[7], 0);
        } else if (by_mode == 1) {
          bool a = mbar_try_wait(&bar[7], 0);
          bool b = mbar_try_wait(&bar[5], 0);
          while (!a) a = mbar_try_wait(&bar[7], 0);
          while (!b) b = mbar_try_wait(&bar[5], 0);
        } else if (by_mode == 2) {
#pragma unroll
          for (int i = 0; i < 64; ++i) asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(n));
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) x = chase[x];
        }
        ++n;
      }
      long long t1 = clock64();
      if (blockIdx.x == 0) {
        stamps[0] = t1 - t0;
        stamps[1] = n;
        stamps[2] = x;
      }
    }
  } else if (warp < 2 + pollers) {
    // warps spinning on a barrier that never completes (like idle epilogue warps)
    while (!stop) {
      if (mbar_try_wait(&bar[6], 0)) break;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(cols) : "memory");
  }
}

// cost of the synchronisation primitives themselves, one thread, barrier already complete
__global__ void probe(long long* out) {
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar)) : "memory");   // phase 0 done
    long long t0 = clock64();
    for (int i = 0; i < 1000; ++i) mbar_wait(&bar, 0);
    long long t1 = clock64();
    long long acc = 0;
    for (int i = 0; i < 1000; ++i) acc += clock64();
    long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t1;
    out[2] = acc;
  }
}

int main() {
  long long* cyc;
  cudaMalloc(&cyc, 148 * 4 * 8);
  long long h[148 * 4];
  long long* stamps;
  cudaMalloc(&stamps, 4 * 8 * 8);
  long long hs[32];
  probe<<<1, 32>>>(cyc);
  cudaDeviceSynchronize();
  cudaMemcpy(h, cyc, 24, cudaMemcpyDeviceToHost);
  printf("mbarrier.try_wait on a completed phase: %.1f cycles; clock64(): %.1f cycles\n", h[0] / 1000.0, h[1] / 1000.0);
  const int iters = 40000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  struct Cfg { int N, nbuf, ctas, nstage, pollers, ts, cols, by; };
  const Cfg cfgs[] = {{128, 2, 2, 3, 0, 0, 256, 0}, {128, 2, 2, 3, 0, 0, 256, 1}, {128, 2, 2, 3, 0, 0, 256, 2},
                      {128, 2, 2, 3, 0, 0, 256, 3}, {256, 2, 1, 3, 0, 0, 512, 0}, {256, 2, 1, 3, 0, 0, 512, 2},
                      {256, 2, 1, 3, 0, 0, 512, 3}, {112, 2, 2, 3, 0, 1, 256, 0}, {240, 2, 1, 3, 0, 1, 512, 0}};
  for (const Cfg& c : cfgs) {
    const size_t smem = 16384 + (size_t)c.nstage * c.N * 128 + 1024;
    // force the requested residency with a shared-memory footprint: 1 CTA/SM -> > half of the SM
    const size_t req = c.ctas == 1 ? (smem > 120 * 1024 ? smem : 120 * 1024) : smem;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)req);
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      k<<<148 * c.ctas, 64 + 32 * c.pollers, req>>>(c.N, c.nbuf, iters, c.nstage, cyc, stamps, c.pollers, c.ts, c.cols, c.by);
      cudaEventRecord(e1);
      cudaError_t e = cudaDeviceSynchronize();
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      cudaMemcpy(h, cyc, 8 * 148 * c.ctas, cudaMemcpyDeviceToHost);
      cudaMemcpy(hs, stamps, 16, cudaMemcpyDeviceToHost);
      double mean = 0;
      for (int i = 0; i < 148 * c.ctas; ++i) mean += h[i];
      mean /= 148 * c.ctas;
      const double flop = 2.0 * 128 * c.N * 64 * (double)iters * 148 * c.ctas;
      if (rep == 1)
        printf("%s N=%3d nbuf=%d ctas/SM=%d pollers=%d: %7.1f cycles per 4-MMA chain per CTA (ideal %4d), %7.1f TFLOP/s by "
               "wall clock at %.0f MHz; bystander mode %d: %.0f cycles per iteration (%s)\n",
               c.ts ? "A-in-TMEM" : "A-in-smem", c.N, c.nbuf, c.ctas, c.pollers, mean / iters, 4 * c.N / 2 * c.ctas, flop / ms / 1e9, mean / ms / 1e3,
               c.by, (double)hs[0] / (hs[1] ? hs[1] : 1), cudaGetErrorString(e));
    }
  }
  return 0;
}
