// topk.cuh — shared pieces of the two full-ranking scorers (exact SIMT fp32 / tcgen05 bf16).
//
// Ordering contract (include/spex_b200.h): score descending, ties by ascending item id.  That is a
// strict total order over (score, id), so the top-k SET and its ORDER do not depend on the order
// in which candidates are offered — the kernels may collect candidates in any order and still be
// bit-reproducible.
#pragma once
#include "common.cuh"

namespace spex {

// does (sa, ia) rank strictly before (sb, ib)?
__device__ __forceinline__ bool beats(float sa, int ia, float sb, int ib) {
  return sa > sb || (sa == sb && ia < ib);
}

// is `item` in the ascending run col[lo, hi)?  (training items of one user: CSR row of R)
__device__ __forceinline__ bool mask_contains(const int32_t* __restrict__ col, int64_t lo, int64_t hi,
                                              int item) {
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    const int c = __ldg(col + mid);
    if (c == item) return true;
    if (c < item) lo = mid + 1; else hi = mid;
  }
  return false;
}

// Warp-cooperative insertion of one candidate into a best-first sorted list of length *n (<= k)
// held in shared memory.  All 32 lanes call it with the same arguments.
__device__ __forceinline__ void warp_list_insert(float* lv, int* li, int& n, int k, float s, int id,
                                                 int lane) {
  int rank = 0;
  for (int p0 = 0; p0 < n; p0 += 32) {
    const int p = p0 + lane;
    const bool b = (p < n) && beats(lv[p], li[p], s, id);
    rank += __popc(__ballot_sync(kFull, b));
  }
  if (rank >= k) return;
  const int newn = (n < k) ? n + 1 : k;
  // shift [rank, newn-1) one slot down, highest chunk first so reads precede overwrites
  for (int p0 = ((newn - 2 - rank) / 32) * 32 + rank; p0 >= rank; p0 -= 32) {
    const int p = p0 + lane;
    float v = 0.f;
    int i = 0;
    const bool mv = (p >= rank) && (p < newn - 1);
    if (mv) {
      v = lv[p];
      i = li[p];
    }
    __syncwarp();
    if (mv) {
      lv[p + 1] = v;
      li[p + 1] = i;
    }
    __syncwarp();
  }
  if (lane == 0) {
    lv[rank] = s;
    li[rank] = id;
  }
  __syncwarp();
  n = newn;
}

}  // namespace spex
