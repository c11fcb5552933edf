// capi.cu — housekeeping entry points of libspex_b200 (version, errors, device check, CUDA-IPC
// helpers for the row-partitioned multi-GPU path and the NGCF dense epilogue).
// See include/spex_b200.h for the contract of each function.
#include "common.cuh"
#include <cuda_bf16.h>
#include <string.h>

namespace spex {

// NGCF dense epilogue, D == 64: 8 warps per CTA, W1^T / W2^T resident in shared memory,
// warp per row (grid-stride), lane owns output features lane and lane+32.
__global__ void __launch_bounds__(256)
ngcf_epilogue_kernel(const float* __restrict__ ego, const float* __restrict__ side,
                     const float* __restrict__ W1, const float* __restrict__ b1,
                     const float* __restrict__ W2, const float* __restrict__ b2, int64_t n,
                     float slope, float* __restrict__ out, float* __restrict__ norm,
                     int64_t norm_stride) {
  __shared__ float w1t[64 * 64];
  __shared__ float w2t[64 * 64];
  for (int i = threadIdx.x; i < 64 * 64; i += 256) {
    const int o = i >> 6, k = i & 63;  // W[o][k] -> Wt[k][o]
    w1t[k * 64 + o] = W1[i];
    w2t[k * 64 + o] = W2[i];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const float bb1[2] = {b1 ? b1[lane] : 0.f, b1 ? b1[lane + 32] : 0.f};
  const float bb2[2] = {b2 ? b2[lane] : 0.f, b2 ? b2[lane + 32] : 0.f};
  int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int64_t stride = (int64_t)gridDim.x * 8;
  for (; row < n; row += stride) {
    const float s0 = side[row * 64 + lane], s1 = side[row * 64 + lane + 32];
    const float e0 = ego[row * 64 + lane] * s0, e1 = ego[row * 64 + lane + 32] * s1;
    float a0 = 0.f, a1 = 0.f, c0 = 0.f, c1 = 0.f;
#pragma unroll 8
    for (int k = 0; k < 64; ++k) {
      const float sk = __shfl_sync(kFull, (k < 32) ? s0 : s1, k & 31);
      const float ek = __shfl_sync(kFull, (k < 32) ? e0 : e1, k & 31);
      a0 = fmaf(sk, w1t[k * 64 + lane], a0);
      a1 = fmaf(sk, w1t[k * 64 + lane + 32], a1);
      c0 = fmaf(ek, w2t[k * 64 + lane], c0);
      c1 = fmaf(ek, w2t[k * 64 + lane + 32], c1);
    }
    a0 += bb1[0]; a1 += bb1[1]; c0 += bb2[0]; c1 += bb2[1];
    const float o0 = (a0 > 0.f ? a0 : a0 * slope) + (c0 > 0.f ? c0 : c0 * slope);
    const float o1 = (a1 > 0.f ? a1 : a1 * slope) + (c1 > 0.f ? c1 : c1 * slope);
    if (out) {
      out[row * 64 + lane] = o0;
      out[row * 64 + lane + 32] = o1;
    }
    if (norm) {
      const float nn = fmaxf(sqrtf(warp_sum(o0 * o0 + o1 * o1)), 1e-12f);
      norm[row * norm_stride + lane] = o0 / nn;
      norm[row * norm_stride + lane + 32] = o1 / nn;
    }
  }
}

}  // namespace spex

using namespace spex;

extern "C" int spex_abi_version(void) { return SPEX_ABI_VERSION; }

extern "C" const char* spex_error_string(int code) {
  switch (code) {
    case 0: return "ok";
    case SPEX_E_BADARG: return "SPEX_E_BADARG: null pointer or negative size";
    case SPEX_E_BADDIM: return "SPEX_E_BADDIM: embedding dim must be a multiple of 4 and <= 512";
    case SPEX_E_ALIGN: return "SPEX_E_ALIGN: pointer not 16-byte aligned";
    case SPEX_E_ARCH: return "SPEX_E_ARCH: device is not sm_100 (no fallback path)";
    case SPEX_E_TOOBIG: return "SPEX_E_TOOBIG: size exceeds a compiled-in limit";
    case SPEX_E_WORKSPACE: return "SPEX_E_WORKSPACE: workspace too small";
    default: break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "unknown spex error";
}

extern "C" int spex_device_check(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  cudaDeviceProp p;
  e = cudaGetDeviceProperties(&p, dev);
  if (e != cudaSuccess) return (int)e;
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  return (p.major == 10) ? 0 : SPEX_E_ARCH;
}

extern "C" int64_t spex_launch_count(void) { return g_launches; }

extern "C" int spex_ngcf_epilogue_f32(const float* ego, const float* side, const float* W1,
                                      const float* b1, const float* W2, const float* b2, int64_t n,
                                      int32_t D, float negative_slope, float* out, float* norm,
                                      int64_t norm_stride, void* stream) {
  SPEX_RETURN_IF(!ego || !side || !W1 || !W2 || n < 0 || (!out && !norm), SPEX_E_BADARG);
  SPEX_RETURN_IF(D != 64, SPEX_E_BADDIM);
  SPEX_RETURN_IF(norm && norm_stride < 64, SPEX_E_BADARG);
  if (n == 0) return 0;
  int64_t blocks = (n + 7) / 8;
  if (blocks > 148 * 4) blocks = 148 * 4;
  ngcf_epilogue_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      ego, side, W1, b1, W2, b2, n, negative_slope, out, norm, norm_stride);
  count_launch();
  return check_last();
}

// ---- CUDA IPC -----------------------------------------------------------------------------------
extern "C" int spex_ipc_alloc(int64_t bytes, void** dev_ptr, void* handle64_host) {
  SPEX_RETURN_IF(bytes <= 0 || !dev_ptr || !handle64_host, SPEX_E_BADARG);
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, (size_t)bytes);
  if (e != cudaSuccess) return (int)e;
  cudaIpcMemHandle_t h;
  e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return (int)e;
  }
  memcpy(handle64_host, &h, 64);
  *dev_ptr = p;
  return 0;
}

extern "C" int spex_ipc_open(const void* handle64_host, void** dev_ptr) {
  SPEX_RETURN_IF(!handle64_host || !dev_ptr, SPEX_E_BADARG);
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64_host, 64);
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return (int)e;
  *dev_ptr = p;
  return 0;
}

extern "C" int spex_ipc_close(void* dev_ptr) {
  SPEX_RETURN_IF(!dev_ptr, SPEX_E_BADARG);
  return (int)cudaIpcCloseMemHandle(dev_ptr);
}

extern "C" int spex_ipc_free(void* dev_ptr) {
  SPEX_RETURN_IF(!dev_ptr, SPEX_E_BADARG);
  return (int)cudaFree(dev_ptr);
}

// copy-engine transfer into an IPC-mapped peer table (E^(0) all-gather without using any SM)
extern "C" int spex_memcpy_peer_async(void* dst, const void* src, int64_t bytes, void* stream) {
  SPEX_RETURN_IF(!dst || !src || bytes < 0, SPEX_E_BADARG);
  if (bytes == 0) return 0;
  return (int)cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
}
