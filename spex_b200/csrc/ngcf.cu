// ngcf.cu — NGCF layer for TRAINING on the library's own kernels.
//
// Reference layer, NGCF_SPEX/code/main_rec.py:76-82 (D = 64):
//     side = A . ego                                     (spex_spmm_csr_f32)
//     z1 = side . W1^T + b1        z2 = (ego * side) . W2^T + b2
//     h  = lrelu(z1) + lrelu(z2)   hd = h * mask         (mask = nn.Dropout's 0 / 1/(1-p) pattern, drawn by
//     y  = hd / max(|hd|_2, 1e-12)                         the caller with the reference's own RNG call)
// Forward (spex_ngcf_layer_fwd_f32): one warp per row, W1^T / W2^T resident in shared memory; writes hd
// (the next layer's ego) and y into the concat buffer.
// Backward (spex_ngcf_layer_bwd_f32), given d(hd) from the next layer (or NULL) and d(y):
//     d_hd = d_hd_next + (dy - y (y . dy)) / n            n = max(|hd|, 1e-12) (dy / n below the clamp)
//     dz1 = d_hd * mask * lrelu'(z1)     dz2 = d_hd * mask * lrelu'(z2)      (z recomputed, not stored)
//     d_side = dz1 . W1 + (dz2 . W2) * ego      d_ego = (dz2 . W2) * side
//     dW1 = dz1^T . side   db1 = sum dz1   dW2 = dz2^T . (ego * side)   db2 = sum dz2
// The weight gradients are reduced WITHOUT atomics: a CTA walks its 32-row tiles in a fixed order, every
// thread owns 32 entries of (dW1 | dW2) in registers and adds the tile's rows in row order; the CTAs'
// partials go to a workspace and a second kernel sums them in CTA order: bit-reproducible.
#include "common.cuh"

namespace spex {
namespace ngcf {

constexpr int D = 64;
constexpr int kWarps = 8;
constexpr int kTile = 32;                  // rows per tile of the backward
constexpr int kBwdBlocks = SPEX_NGCF_BWD_BLOCKS;

__device__ __forceinline__ float lrelu(float x, float s) { return x > 0.f ? x : x * s; }

// ---- forward -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWarps * 32)
layer_fwd_kernel(const float* __restrict__ ego, const float* __restrict__ side, const float* __restrict__ W1,
                 const float* __restrict__ b1, const float* __restrict__ W2, const float* __restrict__ b2,
                 const float* __restrict__ mask, int64_t n, float slope, float* __restrict__ hd,
                 float* __restrict__ norm, int64_t norm_stride) {
  __shared__ float w1t[D * D];
  __shared__ float w2t[D * D];
  for (int i = threadIdx.x; i < D * D; i += kWarps * 32) {
    const int o = i >> 6, k = i & 63;  // W[o][k] -> Wt[k][o]
    w1t[k * D + o] = W1[i];
    w2t[k * D + o] = W2[i];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const float bb1[2] = {b1 ? b1[lane] : 0.f, b1 ? b1[lane + 32] : 0.f};
  const float bb2[2] = {b2 ? b2[lane] : 0.f, b2 ? b2[lane + 32] : 0.f};
  int64_t row = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
  const int64_t stride = (int64_t)gridDim.x * kWarps;
  for (; row < n; row += stride) {
    const float s0 = side[row * D + lane], s1 = side[row * D + lane + 32];
    const float e0 = ego[row * D + lane] * s0, e1 = ego[row * D + lane + 32] * s1;
    float a0 = 0.f, a1 = 0.f, c0 = 0.f, c1 = 0.f;
#pragma unroll 8
    for (int k = 0; k < D; ++k) {
      const float sk = __shfl_sync(kFull, (k < 32) ? s0 : s1, k & 31);
      const float ek = __shfl_sync(kFull, (k < 32) ? e0 : e1, k & 31);
      a0 = fmaf(sk, w1t[k * D + lane], a0);
      a1 = fmaf(sk, w1t[k * D + lane + 32], a1);
      c0 = fmaf(ek, w2t[k * D + lane], c0);
      c1 = fmaf(ek, w2t[k * D + lane + 32], c1);
    }
    a0 += bb1[0]; a1 += bb1[1]; c0 += bb2[0]; c1 += bb2[1];
    float o0 = lrelu(a0, slope) + lrelu(c0, slope), o1 = lrelu(a1, slope) + lrelu(c1, slope);
    if (mask) {
      o0 *= mask[row * D + lane];
      o1 *= mask[row * D + lane + 32];
    }
    if (hd) {
      hd[row * D + lane] = o0;
      hd[row * D + lane + 32] = o1;
    }
    if (norm) {
      const float nn = fmaxf(sqrtf(warp_sum(o0 * o0 + o1 * o1)), 1e-12f);
      norm[row * norm_stride + lane] = o0 / nn;
      norm[row * norm_stride + lane + 32] = o1 / nn;
    }
  }
}

// ---- backward ------------------------------------------------------------------------------------------
// dynamic shared memory: W1t, W2t (forward orientation), W1, W2 (as stored: [o][k]) = 4 x 16 KB,
// tile buffers dz1, dz2, side, es: 4 x [kTile][D] = 32 KB
__global__ void __launch_bounds__(kWarps * 32)
layer_bwd_kernel(const float* __restrict__ ego, const float* __restrict__ side, const float* __restrict__ W1,
                 const float* __restrict__ b1, const float* __restrict__ W2, const float* __restrict__ b2,
                 const float* __restrict__ mask, const float* __restrict__ d_hd_next,
                 const float* __restrict__ d_norm, int64_t d_norm_stride, int64_t n, float slope,
                 float* __restrict__ d_ego, float* __restrict__ d_side, float* __restrict__ work) {
  extern __shared__ __align__(16) float sm[];
  float* w1t = sm;
  float* w2t = sm + D * D;
  float* w1 = sm + 2 * D * D;
  float* w2 = sm + 3 * D * D;
  float* tz1 = sm + 4 * D * D;             // [kTile][D]
  float* tz2 = tz1 + kTile * D;
  float* tsd = tz2 + kTile * D;
  float* tes = tsd + kTile * D;
  for (int i = threadIdx.x; i < D * D; i += kWarps * 32) {
    const int o = i >> 6, k = i & 63;
    const float a = W1[i], b = W2[i];
    w1[i] = a;
    w2[i] = b;
    w1t[k * D + o] = a;
    w2t[k * D + o] = b;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float bb1[2] = {b1 ? b1[lane] : 0.f, b1 ? b1[lane + 32] : 0.f};
  const float bb2[2] = {b2 ? b2[lane] : 0.f, b2 ? b2[lane + 32] : 0.f};
  // weight-gradient accumulators: thread t owns outputs o = t / 4 (of 64) and k in [(t%4)*16, +16) of BOTH
  // matrices: 32 registers; bias sums: thread t < 128 owns (matrix t / 64, output t % 64)
  const int oo = threadIdx.x >> 2, k0 = (threadIdx.x & 3) * 16;
  float acc1[16], acc2[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc1[i] = acc2[i] = 0.f;
  float accb = 0.f;
  const int64_t n_tiles = (n + kTile - 1) / kTile;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    // phase 1: warp per row (4 rows per warp), everything of the row but the weight gradients
    for (int rr = warp; rr < kTile; rr += kWarps) {
      const int64_t row = tile * kTile + rr;
      float z10 = 0.f, z11 = 0.f, z20 = 0.f, z21 = 0.f, s0 = 0.f, s1 = 0.f, es0 = 0.f, es1 = 0.f;
      if (row < n) {
        s0 = side[row * D + lane];
        s1 = side[row * D + lane + 32];
        const float g0 = ego[row * D + lane], g1 = ego[row * D + lane + 32];
        es0 = g0 * s0;
        es1 = g1 * s1;
        float a0 = 0.f, a1 = 0.f, c0 = 0.f, c1 = 0.f;
#pragma unroll 8
        for (int k = 0; k < D; ++k) {
          const float sk = __shfl_sync(kFull, (k < 32) ? s0 : s1, k & 31);
          const float ek = __shfl_sync(kFull, (k < 32) ? es0 : es1, k & 31);
          a0 = fmaf(sk, w1t[k * D + lane], a0);
          a1 = fmaf(sk, w1t[k * D + lane + 32], a1);
          c0 = fmaf(ek, w2t[k * D + lane], c0);
          c1 = fmaf(ek, w2t[k * D + lane + 32], c1);
        }
        a0 += bb1[0]; a1 += bb1[1]; c0 += bb2[0]; c1 += bb2[1];
        const float m0 = mask ? mask[row * D + lane] : 1.f, m1 = mask ? mask[row * D + lane + 32] : 1.f;
        const float h0 = (lrelu(a0, slope) + lrelu(c0, slope)) * m0, h1 = (lrelu(a1, slope) + lrelu(c1, slope)) * m1;
        const float nrm = sqrtf(warp_sum(h0 * h0 + h1 * h1));
        const float nn = fmaxf(nrm, 1e-12f);
        const float y0 = h0 / nn, y1 = h1 / nn;
        const float dy0 = d_norm ? d_norm[row * d_norm_stride + lane] : 0.f;
        const float dy1 = d_norm ? d_norm[row * d_norm_stride + lane + 32] : 0.f;
        // F.normalize backward: x / max(|x|, eps): above the clamp (dy - y (y.dy)) / n, below it dy / eps
        const float dot = (nrm > 1e-12f) ? warp_sum(y0 * dy0 + y1 * dy1) : 0.f;
        float dh0 = (dy0 - y0 * dot) / nn, dh1 = (dy1 - y1 * dot) / nn;
        if (d_hd_next) {
          dh0 += d_hd_next[row * D + lane];
          dh1 += d_hd_next[row * D + lane + 32];
        }
        dh0 *= m0;
        dh1 *= m1;
        z10 = dh0 * (a0 > 0.f ? 1.f : slope);
        z11 = dh1 * (a1 > 0.f ? 1.f : slope);
        z20 = dh0 * (c0 > 0.f ? 1.f : slope);
        z21 = dh1 * (c1 > 0.f ? 1.f : slope);
        // t1[k] = sum_o dz1[o] W1[o][k], t2[k] = sum_o dz2[o] W2[o][k]   (lane owns k = lane, lane + 32)
        float t10 = 0.f, t11 = 0.f, t20 = 0.f, t21 = 0.f;
#pragma unroll 8
        for (int o = 0; o < D; ++o) {
          const float d1 = __shfl_sync(kFull, (o < 32) ? z10 : z11, o & 31);
          const float d2 = __shfl_sync(kFull, (o < 32) ? z20 : z21, o & 31);
          t10 = fmaf(d1, w1[o * D + lane], t10);
          t11 = fmaf(d1, w1[o * D + lane + 32], t11);
          t20 = fmaf(d2, w2[o * D + lane], t20);
          t21 = fmaf(d2, w2[o * D + lane + 32], t21);
        }
        d_side[row * D + lane] = t10 + t20 * g0;
        d_side[row * D + lane + 32] = t11 + t21 * g1;
        d_ego[row * D + lane] = t20 * s0;
        d_ego[row * D + lane + 32] = t21 * s1;
      }
      tz1[rr * D + lane] = z10;
      tz1[rr * D + lane + 32] = z11;
      tz2[rr * D + lane] = z20;
      tz2[rr * D + lane + 32] = z21;
      tsd[rr * D + lane] = s0;
      tsd[rr * D + lane + 32] = s1;
      tes[rr * D + lane] = es0;
      tes[rr * D + lane + 32] = es1;
    }
    __syncthreads();
    // phase 2: the tile's rows into the weight-gradient accumulators, in row order
#pragma unroll 4
    for (int rr = 0; rr < kTile; ++rr) {
      const float d1 = tz1[rr * D + oo], d2 = tz2[rr * D + oo];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        acc1[i] = fmaf(d1, tsd[rr * D + k0 + i], acc1[i]);
        acc2[i] = fmaf(d2, tes[rr * D + k0 + i], acc2[i]);
      }
    }
    if (threadIdx.x < 128) {
      const float* tz = (threadIdx.x < 64) ? tz1 : tz2;
      const int o = threadIdx.x & 63;
#pragma unroll 4
      for (int rr = 0; rr < kTile; ++rr) accb += tz[rr * D + o];
    }
    __syncthreads();
  }
  // CTA partials: work[block][0 .. 8192) = dW1 | dW2 (row-major [o][k]), [8192 .. 8320) = db1 | db2
  float* wk = work + (size_t)blockIdx.x * SPEX_NGCF_BWD_WORK_PER_BLOCK;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    wk[oo * D + k0 + i] = acc1[i];
    wk[D * D + oo * D + k0 + i] = acc2[i];
  }
  if (threadIdx.x < 128) wk[2 * D * D + threadIdx.x] = accb;
}

// sum the CTA partials in CTA order: thread per output element
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ work, int n_blocks, float* __restrict__ dW1,
                    float* __restrict__ db1, float* __restrict__ dW2, float* __restrict__ db2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= SPEX_NGCF_BWD_WORK_PER_BLOCK) return;
  float t = 0.f;
  for (int b = 0; b < n_blocks; ++b) t += work[(size_t)b * SPEX_NGCF_BWD_WORK_PER_BLOCK + i];
  if (i < D * D) dW1[i] = t;
  else if (i < 2 * D * D) dW2[i - D * D] = t;
  else if (i < 2 * D * D + 64) { if (db1) db1[i - 2 * D * D] = t; }
  else if (db2) db2[i - 2 * D * D - 64] = t;
}

}  // namespace ngcf
}  // namespace spex

using namespace spex;

extern "C" int spex_ngcf_layer_fwd_f32(const float* ego, const float* side, const float* W1, const float* b1,
                                       const float* W2, const float* b2, const float* mask, int64_t n,
                                       int32_t D, float negative_slope, float* hd, float* norm,
                                       int64_t norm_stride, void* stream) {
  SPEX_RETURN_IF(!ego || !side || !W1 || !W2 || n < 0 || (!hd && !norm), SPEX_E_BADARG);
  SPEX_RETURN_IF(D != 64, SPEX_E_BADDIM);
  SPEX_RETURN_IF(norm && norm_stride < 64, SPEX_E_BADARG);
  if (n == 0) return 0;
  int64_t blocks = (n + ngcf::kWarps - 1) / ngcf::kWarps;
  if (blocks > 148 * 4) blocks = 148 * 4;
  ngcf::layer_fwd_kernel<<<(unsigned)blocks, ngcf::kWarps * 32, 0, (cudaStream_t)stream>>>(
      ego, side, W1, b1, W2, b2, mask, n, negative_slope, hd, norm, norm_stride);
  count_launch();
  return check_last();
}

extern "C" int spex_ngcf_layer_bwd_f32(const float* ego, const float* side, const float* W1, const float* b1,
                                       const float* W2, const float* b2, const float* mask,
                                       const float* d_hd_next, const float* d_norm, int64_t d_norm_stride,
                                       int64_t n, int32_t D, float negative_slope, float* d_ego, float* d_side,
                                       float* dW1, float* db1, float* dW2, float* db2, float* work,
                                       void* stream) {
  SPEX_RETURN_IF(!ego || !side || !W1 || !W2 || !d_ego || !d_side || !dW1 || !dW2 || !work || n < 0, SPEX_E_BADARG);
  SPEX_RETURN_IF(!d_hd_next && !d_norm, SPEX_E_BADARG);
  SPEX_RETURN_IF(D != 64, SPEX_E_BADDIM);
  SPEX_RETURN_IF(d_norm && d_norm_stride < 64, SPEX_E_BADARG);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n_tiles = (n + ngcf::kTile - 1) / ngcf::kTile;
  int blocks = (int)(n_tiles < ngcf::kBwdBlocks ? n_tiles : ngcf::kBwdBlocks);
  if (blocks < 1) blocks = 1;
  const size_t smem = (size_t)(4 * 64 * 64 + 4 * ngcf::kTile * 64) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(ngcf::layer_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  ngcf::layer_bwd_kernel<<<blocks, ngcf::kWarps * 32, smem, st>>>(ego, side, W1, b1, W2, b2, mask, d_hd_next, d_norm,
                                                                  d_norm_stride, n, negative_slope, d_ego, d_side, work);
  ngcf::wgrad_reduce_kernel<<<(SPEX_NGCF_BWD_WORK_PER_BLOCK + 255) / 256, 256, 0, st>>>(work, blocks, dW1, db1, dW2, db2);
  count_launch(2);
  return check_last();
}
