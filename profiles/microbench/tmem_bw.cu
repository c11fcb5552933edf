// tmem_bw.cu — how fast can registers be filled from TMEM on B200 (sm_100a)?
// The full-ranking scorer reads every fp32 accumulator element after only 64 MACs, so the
// tcgen05.ld rate is a hard bound of that kernel.  This microbenchmark measures it per SM for
// several instruction shapes and warp counts.   nvcc -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int SHAPE>
__device__ __forceinline__ uint32_t do_ld(uint32_t taddr) {
  uint32_t x = 0;
  if constexpr (SHAPE == 0) {  // 32x32b.x32 : 32 lanes x 32 columns = 4 KB
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
#pragma unroll
    for (int i = 0; i < 32; ++i) x ^= r[i];
  } else if constexpr (SHAPE == 1) {  // 32x32b.x8 : 1 KB
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) x ^= r[i];
  } else if constexpr (SHAPE == 2) {  // 16x256b.x4 : 16 lanes x 256 bit x 4 = 32 regs/thread? (x4 -> 16 regs)
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) x ^= r[i];
  } else if constexpr (SHAPE == 3) {  // 32x32b.x32 issued twice back to back before one wait (8 KB)
    uint32_t r[64];
#define LD32(o, a)                                                                                             \
    asm volatile(                                                                                              \
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                              \
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
        : "=r"(r[o+0]), "=r"(r[o+1]), "=r"(r[o+2]), "=r"(r[o+3]), "=r"(r[o+4]), "=r"(r[o+5]), "=r"(r[o+6]), "=r"(r[o+7]), \
          "=r"(r[o+8]), "=r"(r[o+9]), "=r"(r[o+10]), "=r"(r[o+11]), "=r"(r[o+12]), "=r"(r[o+13]), "=r"(r[o+14]), "=r"(r[o+15]), \
          "=r"(r[o+16]), "=r"(r[o+17]), "=r"(r[o+18]), "=r"(r[o+19]), "=r"(r[o+20]), "=r"(r[o+21]), "=r"(r[o+22]), "=r"(r[o+23]), \
          "=r"(r[o+24]), "=r"(r[o+25]), "=r"(r[o+26]), "=r"(r[o+27]), "=r"(r[o+28]), "=r"(r[o+29]), "=r"(r[o+30]), "=r"(r[o+31]) \
        : "r"(a) : "memory")
    LD32(0, taddr);
    LD32(32, taddr + 32);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 64; ++i) x ^= r[i];
  }
  return x;
}

template <int SHAPE>
__global__ void __launch_bounds__(512, 1) tmem_ld_kernel(int iters, uint32_t* out, long long* cycles) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 64) % 448;
  uint32_t x = 0;
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) x ^= do_ld<SHAPE>(base);
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u) : "memory");
  }
}

template <int SHAPE>
void run(const char* name, int bytes_per_warp_ld, int warps) {
  uint32_t* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4);
  cudaMalloc(&cyc, 148 * 8);
  const int iters = 20000;
  tmem_ld_kernel<SHAPE><<<148, warps * 32>>>(100, out, cyc);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  tmem_ld_kernel<SHAPE><<<148, warps * 32>>>(iters, out, cyc);
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  long long c;
  cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  const double bytes = (double)iters * warps * bytes_per_warp_ld;
  printf("%-28s warps=%2d  %8.1f B/cycle/SM  %7.1f cycles per ld per warp  (%.2f ms, %s)\n", name, warps,
         bytes / (double)c, (double)c / iters, ms, cudaGetErrorString(err));
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  for (int w : {1, 4, 8, 16}) run<0>("32x32b.x32 (4 KB/warp)", 4096, w);
  for (int w : {4, 8, 16}) run<1>("32x32b.x8 (1 KB/warp)", 1024, w);
  for (int w : {4, 8, 16}) run<2>("16x256b.x4 (2 KB/warp)", 2048, w);
  for (int w : {4, 8, 16}) run<3>("2 x 32x32b.x32, one wait", 8192, w);
  return 0;
}
