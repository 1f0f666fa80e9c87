// topk_dev.cuh -- block-level selection of the k <= 32 best (score, id) candidates, shared by the
// candidate merge (topk.cu) and the one-kernel exchange (sharded.cu).
// Order everywhere: score descending, then id ascending; id < 0 marks padding and is never selected.
#pragma once
#include "common.cuh"

namespace evdr {

constexpr int SEL_WARPS = 8;                     // 256 threads per CTA

__device__ __forceinline__ bool cand_better(float s, long long i, float os, long long oi) {
  return (s > os) || (s == os && i < oi);
}

// All SEL_WARPS * 32 threads of the CTA call this.  Candidates cs / ci [n] (global or shared memory) ->
// the k best in sel_s / sel_i [k] (shared memory), best first, padded with (-inf, -1).
// ws / wi are shared scratch of SEL_WARPS * 32 entries each.  Ends with a __syncthreads().
__device__ __forceinline__ void block_select_topk(const float* __restrict__ cs, const long long* __restrict__ ci,
                                                  int n, int k, float* ws, long long* wi, float* sel_s,
                                                  long long* sel_i) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // per-warp register list (lane j = j-th best so far) over a strided slice of the candidates
  float my_s = -INFINITY;
  long long my_i = -1;
  float thr_s = -INFINITY;
  long long thr_i = 0x7fffffffffffffffLL;
  int filled = 0;
  for (int base = warp * 32; base < n; base += SEL_WARPS * 32) {
    const int i = base + lane;
    float s = -INFINITY;
    long long id = -1;
    if (i < n) { s = cs[i]; id = ci[i]; }
    bool ok = id >= 0 && cand_better(s, id, thr_s, thr_i);
    unsigned m = __ballot_sync(0xffffffffu, ok);
    while (m) {
      const int src = __ffs(m) - 1;
      m &= m - 1;
      const float xs = __shfl_sync(0xffffffffu, s, src);
      const long long xi = __shfl_sync(0xffffffffu, id, src);
      if (!cand_better(xs, xi, thr_s, thr_i)) continue;           // the bar may have moved
      const bool ahead = my_i >= 0 && cand_better(my_s, my_i, xs, xi);
      const int pos = __popc(__ballot_sync(0xffffffffu, ahead));
      const float us = __shfl_up_sync(0xffffffffu, my_s, 1);
      const long long ui = __shfl_up_sync(0xffffffffu, my_i, 1);
      if (lane == pos) { my_s = xs; my_i = xi; }
      else if (lane > pos) { my_s = us; my_i = ui; }
      if (++filled >= k) {
        thr_s = __shfl_sync(0xffffffffu, my_s, k - 1);
        thr_i = __shfl_sync(0xffffffffu, my_i, k - 1);
      }
    }
  }
  ws[warp * 32 + lane] = my_s;
  wi[warp * 32 + lane] = my_i;
  for (int j = tid; j < k; j += blockDim.x) { sel_s[j] = -INFINITY; sel_i[j] = -1; }
  __syncthreads();
  // rank every real entry of the SEL_WARPS lists (first k of each) among all of them
  const int total = SEL_WARPS * k;
  for (int e = tid; e < total; e += blockDim.x) {
    const int w = e / k, j = e % k;
    const float es = ws[w * 32 + j];
    const long long ei = wi[w * 32 + j];
    if (ei < 0) continue;
    int rank = 0;
    for (int f = 0; f < total; ++f) {
      const int fw = f / k, fj = f % k;
      const float fs = ws[fw * 32 + fj];
      const long long fi = wi[fw * 32 + fj];
      if (fi >= 0 && (cand_better(fs, fi, es, ei) || (fs == es && fi == ei && f < e))) ++rank;
    }
    if (rank < k) { sel_s[rank] = es; sel_i[rank] = ei; }
  }
  __syncthreads();
}

}  // namespace evdr
