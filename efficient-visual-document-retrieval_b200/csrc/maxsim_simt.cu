// maxsim_simt.cu -- general exact MaxSim (fp32 FMA) forward with argmax, and the
// deterministic backward.  This is the path for shapes the tensor-core kernel does
// not cover (d != 128, lq > 32), for the fp32-exact mode and for training (argmax).
//
// Reference semantics: evaluator/retrieval.py:166-213 (forward), SURVEY 8a A5 (backward).
#include "common.cuh"

namespace evdr {

constexpr int BM = 64;   // query-token rows per CTA
constexpr int BN = 64;   // page tokens per inner tile
constexpr int BK = 32;   // reduction chunk
constexpr float kNegFill = -1e4f;  // evaluator/retrieval.py:185

// rowmax[c*R + r], argmax[c*R + r] for r = q*lq + n  (page-major workspace layout)
template <typename TQ, typename TP>
__global__ void __launch_bounds__(256)
rowmax_kernel(const TQ* __restrict__ Q, const TP* __restrict__ P,
              const uint8_t* __restrict__ pmask, int R, int lp, int d,
              float* __restrict__ rowmax, int32_t* __restrict__ argmax) {
  __shared__ float As[BK][BM + 1];
  __shared__ float Bs[BK][BN + 1];
  __shared__ uint8_t Ms[BN];

  const int tid = threadIdx.x;
  const int tx = tid & 15;        // column group
  const int ty = tid >> 4;        // row group
  const int row0 = blockIdx.x * BM;
  const int c = blockIdx.y;
  const TP* Pc = P + (size_t)c * lp * d;
  const uint8_t* mc = pmask + (size_t)c * lp;

  float best[4];
  int bidx[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { best[i] = -INFINITY; bidx[i] = 0; }

  for (int m0 = 0; m0 < lp; m0 += BN) {
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    if (tid < BN) Ms[tid] = (m0 + tid < lp) ? mc[m0 + tid] : (uint8_t)0;

    for (int k0 = 0; k0 < d; k0 += BK) {
#pragma unroll
      for (int i = 0; i < (BM * BK) / 256; ++i) {
        int idx = tid + i * 256;
        int r = idx / BK, k = idx % BK;
        float va = 0.f, vb = 0.f;
        if (k0 + k < d) {
          if (row0 + r < R) va = to_f32(Q[(size_t)(row0 + r) * d + k0 + k]);
          if (m0 + r < lp) vb = to_f32(Pc[(size_t)(m0 + r) * d + k0 + k]);
        }
        As[k][r] = va;
        Bs[k][r] = vb;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        float a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
    // masked running max; page tokens visited in ascending order, strict '>' keeps the first index
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int m = m0 + tx * 4 + j;
      if (m < lp) {
        bool valid = Ms[tx * 4 + j] != 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float v = valid ? acc[i][j] : kNegFill;
          if (v > best[i]) { best[i] = v; bidx[i] = m; }
        }
      }
    }
    __syncthreads();  // Ms reuse
  }
  // reduce across the 16 column groups (contiguous half-warp lanes)
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float v = best[i];
    int ix = bidx[i];
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, v, o);
      int oi = __shfl_xor_sync(0xffffffffu, ix, o);
      if (ov > v || (ov == v && oi < ix)) { v = ov; ix = oi; }
    }
    int r = row0 + ty * 4 + i;
    if (tx == 0 && r < R) {
      rowmax[(size_t)c * R + r] = v;
      if (argmax) argmax[(size_t)c * R + r] = ix;
    }
  }
}

// out[q, c] = has[c] * sum_n qmask[q,n] * rowmax[c*R + q*lq + n]; one warp per page.
__global__ void __launch_bounds__(256)
rowsum_kernel(const float* __restrict__ rowmax, const uint8_t* __restrict__ qmask,
              const uint8_t* __restrict__ pmask, int nq, int np, int lq, int lp,
              float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= np) return;
  int any = 0;
  for (int m = lane; m < lp; m += 32) any |= pmask[(size_t)c * lp + m];
  any = __any_sync(0xffffffffu, any != 0);
  const float has = any ? 1.f : 0.f;
  const int R = nq * lq;
  const float* rm = rowmax + (size_t)c * R;
  for (int q = 0; q < nq; ++q) {
    float s = 0.f;
    for (int n = lane; n < lq; n += 32) {
      float w = qmask[(size_t)q * lq + n] ? 1.f : 0.f;
      s += (rm[q * lq + n] * has) * w;
    }
    s = warp_sum(s);
    if (lane == 0) out[(size_t)q * np + c] = s;
  }
}

// ---------------------------------------------------------------------------
// backward dP: one CTA per page.  Stable bucket of query-token rows by their
// argmax page token (ascending row order inside each bucket), then each warp sums
// its tokens' rows in that order: deterministic, no floating-point atomics.
// dynamic smem: cnt[lp] + off[lp+1] + list[R] ints.
// ---------------------------------------------------------------------------
template <typename TQ>
__global__ void __launch_bounds__(256)
bwd_dp_kernel(const float* __restrict__ gout, const TQ* __restrict__ Q,
              const uint8_t* __restrict__ qmask, const uint8_t* __restrict__ pmask,
              const int32_t* __restrict__ argmax, int nq, int np, int lq, int lp, int d,
              float* __restrict__ dP) {
  extern __shared__ int smem_i[];
  int* cnt = smem_i;            // lp
  int* off = cnt + lp;          // lp + 1
  int* list = off + lp + 1;     // R
  __shared__ int s_has;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c = blockIdx.x;
  const int R = nq * lq;
  const int32_t* am = argmax + (size_t)c * R;

  if (tid == 0) s_has = 0;
  for (int m = tid; m < lp; m += blockDim.x) cnt[m] = 0;
  __syncthreads();
  int any = 0;
  for (int m = tid; m < lp; m += blockDim.x) any |= pmask[(size_t)c * lp + m];
  if (any) s_has = 1;
  __syncthreads();
  float* dPc = dP + (size_t)c * lp * d;
  if (!s_has) {  // empty page: zero gradient (evaluator/retrieval.py:204)
    for (int i = tid; i < lp * d; i += blockDim.x) dPc[i] = 0.f;
    return;
  }
  // a masked winner (possible only when every valid similarity is below the -1e4 fill) gets no
  // gradient: the reference's masked_fill blocks it
  const uint8_t* mc = pmask + (size_t)c * lp;
  for (int r = tid; r < R; r += blockDim.x)
    if (qmask[r] && mc[am[r]]) atomicAdd(&cnt[am[r]], 1);
  __syncthreads();
  if (warp == 0) {  // exclusive scan over lp counters
    int carry = 0;
    for (int m0 = 0; m0 < lp; m0 += 32) {
      int m = m0 + lane;
      int v = (m < lp) ? cnt[m] : 0;
      int inc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      if (m < lp) off[m] = carry + inc - v;
      carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) off[lp] = carry;
    __syncwarp();
    // stable fill: rows in ascending order, 32 at a time
    for (int m = lane; m < lp; m += 32) cnt[m] = 0;
    __syncwarp();
    for (int r0 = 0; r0 < R; r0 += 32) {
      int r = r0 + lane;
      bool act = (r < R) && qmask[r];
      if (act) act = mc[am[r]] != 0;
      int m = act ? am[r] : -1 - lane;  // inactive lanes get unique keys
      unsigned peers = __match_any_sync(0xffffffffu, m);
      if (act) {
        int rank = __popc(peers & ((1u << lane) - 1u));
        int base = cnt[m];
        list[off[m] + base + rank] = r;
      }
      __syncwarp();
      if (act && (__ffs(peers) - 1) == lane) cnt[m] += __popc(peers);
      __syncwarp();
    }
  }
  __syncthreads();
  const int nwarps = blockDim.x >> 5;
  for (int m = warp; m < lp; m += nwarps) {
    const int b = off[m], e = off[m + 1];
    for (int d0 = lane * 4; d0 < d; d0 += 128) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int i = b; i < e; ++i) {
        int r = list[i];
        float w = gout[(size_t)(r / lq) * np + c];
        float4 qv = load4(Q + (size_t)r * d + d0);
        acc.x = fmaf(w, qv.x, acc.x);
        acc.y = fmaf(w, qv.y, acc.y);
        acc.z = fmaf(w, qv.z, acc.z);
        acc.w = fmaf(w, qv.w, acc.w);
      }
      *reinterpret_cast<float4*>(dPc + (size_t)m * d + d0) = acc;
    }
  }
}

// backward dQ: one warp per query-token row, pages in ascending order.
template <typename TP>
__global__ void __launch_bounds__(256)
bwd_dq_kernel(const float* __restrict__ gout, const TP* __restrict__ P,
              const uint8_t* __restrict__ qmask, const uint8_t* __restrict__ pmask,
              const uint8_t* __restrict__ has_page,
              const int32_t* __restrict__ argmax, int nq, int np, int lq, int lp, int d,
              float* __restrict__ dQ) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int R = nq * lq;
  if (r >= R) return;
  const int q = r / lq;
  const bool act = qmask[r] != 0;
  for (int d0 = lane * 4; d0 < d; d0 += 128) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (act) {
      for (int c = 0; c < np; ++c) {
        if (!has_page[c]) continue;
        float w = gout[(size_t)q * np + c];
        int m = argmax[(size_t)c * R + r];
        if (!pmask[(size_t)c * lp + m]) continue;
        float4 pv = load4(P + ((size_t)c * lp + m) * d + d0);
        acc.x = fmaf(w, pv.x, acc.x);
        acc.y = fmaf(w, pv.y, acc.y);
        acc.z = fmaf(w, pv.z, acc.z);
        acc.w = fmaf(w, pv.w, acc.w);
      }
    }
    *reinterpret_cast<float4*>(dQ + (size_t)r * d + d0) = acc;
  }
}

// Rescoring of tensor-core winners.  `am` holds the packed slot chosen by the bf16 tcgen05 search
// (evdr_maxsim_packed_argmax); a slot that is a filler (masked or beyond lp) is mapped back to the
// valid token it duplicates (nearest preceding valid token, else the first valid one -- the pack
// rule), the token index replaces the slot in `am` (the backward scatters through it) and the exact
// fp32 dot product goes to rowmax[c*R + r].
// One warp = one query-token row x 32 consecutive pages: the lanes keep the row's 128 components in
// registers (4 each), every page costs one broadcast index load, one coalesced 512-byte token load
// and 4 FMAs per lane, and the 32 per-page partial sums are reduced together by a 31-shuffle
// transpose-reduction (lane j ends up with page j's dot product) instead of 32 five-step warp sums.
// d == 128 only; other dims use rescore_generic_kernel.
template <typename TQ, typename TP>
__global__ void __launch_bounds__(256)
rescore_kernel(const TQ* __restrict__ Q, const TP* __restrict__ P, const uint8_t* __restrict__ pmask,
               int32_t* __restrict__ am, int R, int np, int lp, float* __restrict__ rowmax) {
  constexpr int d = 128;
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int c0 = blockIdx.y * 32;
  if (r >= R) return;
  const float4 qv = load4(Q + (size_t)r * d + lane * 4);
  float acc[32];
  int t_mine = 0;
  bool valid_mine = false;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const int c = c0 + j;
    float a = 0.f;
    if (c < np) {                                          // warp-uniform
      const uint8_t* mc = pmask + (size_t)c * lp;
      int t = am[(size_t)c * R + r];
      t = t < 0 ? 0 : (t >= lp ? lp - 1 : t);
      if (!mc[t]) {                                        // filler slot: rare, warp-uniform
        int u = t;
        while (u >= 0 && !mc[u]) --u;
        if (u < 0) { u = 0; while (u < lp && !mc[u]) ++u; }
        t = u;                                             // == lp: page without valid tokens
      }
      const bool ok = t < lp;
      if (ok) {
        const float4 pv = load4(P + ((size_t)c * lp + t) * d + lane * 4);
        a = fmaf(qv.x, pv.x, a);
        a = fmaf(qv.y, pv.y, a);
        a = fmaf(qv.z, pv.z, a);
        a = fmaf(qv.w, pv.w, a);
      }
      if (lane == j) { t_mine = ok ? t : 0; valid_mine = ok; }
    }
    acc[j] = a;
  }
  // transpose-reduce: after step `o` lane l keeps the entries whose index has bit `o` equal to l's
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int j = 0; j < o; ++j) {
      const float mine = up ? acc[j + o] : acc[j];
      const float give = up ? acc[j] : acc[j + o];
      acc[j] = mine + __shfl_xor_sync(0xffffffffu, give, o);
    }
  }
  const int c = c0 + lane;
  if (c < np) {
    am[(size_t)c * R + r] = t_mine;
    rowmax[(size_t)c * R + r] = valid_mine ? acc[0] : kNegFill;   // :198 leaves -1e4 on empty pages
  }
}

// Packed slot -> page-token index for every (page, query-token row): a filler slot (masked or beyond lp)
// names the valid token it duplicates (nearest preceding valid token, else the first valid one -- the pack
// rule); a page without valid tokens maps to 0 (it receives no gradient).
__global__ void __launch_bounds__(256)
slot_to_token_kernel(const uint8_t* __restrict__ pmask, int32_t* __restrict__ am, long long n, int R, int lp) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint8_t* mc = pmask + (i / R) * lp;
  int t = am[i];
  t = t < 0 ? 0 : (t >= lp ? lp - 1 : t);
  if (!mc[t]) {
    int u = t;
    while (u >= 0 && !mc[u]) --u;
    if (u < 0) { u = 0; while (u < lp && !mc[u]) ++u; }
    t = u < lp ? u : 0;
  }
  am[i] = t;
}

// any d % 4 == 0: one warp per (query-token row, page)
template <typename TQ, typename TP>
__global__ void __launch_bounds__(256)
rescore_generic_kernel(const TQ* __restrict__ Q, const TP* __restrict__ P, const uint8_t* __restrict__ pmask,
                       int32_t* __restrict__ am, int R, int lp, int d, float* __restrict__ rowmax) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int c = blockIdx.y;
  if (r >= R) return;
  const uint8_t* mc = pmask + (size_t)c * lp;
  int t = am[(size_t)c * R + r];
  if (t < 0) t = 0;
  if (t >= lp) t = lp - 1;
  if (!mc[t]) {
    int u = t;
    while (u >= 0 && !mc[u]) --u;
    if (u < 0) { u = 0; while (u < lp && !mc[u]) ++u; }
    t = u;
  }
  float v = kNegFill;
  if (t < lp) {
    const TQ* q = Q + (size_t)r * d;
    const TP* p = P + ((size_t)c * lp + t) * d;
    float acc = 0.f;
    for (int i = lane * 4; i < d; i += 128) {
      const float4 a = load4(q + i), b = load4(p + i);
      acc = fmaf(a.x, b.x, acc);
      acc = fmaf(a.y, b.y, acc);
      acc = fmaf(a.z, b.z, acc);
      acc = fmaf(a.w, b.w, acc);
    }
    v = warp_sum(acc);
  } else {
    t = 0;
  }
  if (lane == 0) {
    am[(size_t)c * R + r] = t;
    rowmax[(size_t)c * R + r] = v;
  }
}

template <typename TQ, typename TP>
static int launch_rescore(const void* Q, const void* P, const uint8_t* pmask, int32_t* am, int R, int np,
                          int lp, int d, float* rowmax, cudaStream_t st) {
  if (d == 128) {
    dim3 grid((R + 7) / 8, (np + 31) / 32);
    rescore_kernel<TQ, TP><<<grid, 256, 0, st>>>(static_cast<const TQ*>(Q), static_cast<const TP*>(P), pmask,
                                                  am, R, np, lp, rowmax);
  } else {
    dim3 grid((R + 7) / 8, np);
    rescore_generic_kernel<TQ, TP><<<grid, 256, 0, st>>>(static_cast<const TQ*>(Q), static_cast<const TP*>(P),
                                                          pmask, am, R, lp, d, rowmax);
  }
  EVDR_LAUNCH_CHECK("rescore_kernel");
  return EVDR_OK;
}

// Hard-token mining (mainv3_iter_liscore_QA_hardtoken.py:398-425): for a (query, page) pair the page
// token m maximising  max_{valid n} <Q[q,n], P[c,m]>  over the VALID page tokens (the column-max
// variant of MaxSim; masked page tokens are -inf, ties -> lowest index like torch.argmax).
// One CTA per pair: the query's tokens sit in shared memory, each warp walks page tokens.
template <typename TQ, typename TP>
__global__ void __launch_bounds__(256)
hard_token_kernel(const TQ* __restrict__ Q, const TP* __restrict__ P, const uint8_t* __restrict__ qmask,
                  const uint8_t* __restrict__ pmask, const long long* __restrict__ doc_idx, int A, int lq,
                  int lp, int d, long long* __restrict__ out_idx) {
  extern __shared__ float qs[];                 // lq * d
  __shared__ float wbest[8];
  __shared__ int widx[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int q = blockIdx.x / A;
  const long long c = doc_idx[blockIdx.x];
  for (int i = tid; i < lq * d; i += blockDim.x) qs[i] = to_f32(Q[(size_t)q * lq * d + i]);
  __syncthreads();
  const uint8_t* qm = qmask + (size_t)q * lq;
  const uint8_t* pm = pmask + (size_t)c * lp;
  float best = -INFINITY;
  int bidx = 0x7fffffff;
  for (int m = warp; m < lp; m += 8) {
    if (!pm[m]) continue;                       // sim[:, ~doc_mask] = -inf
    const TP* p = P + ((size_t)c * lp + m) * d;
    float colmax = -INFINITY;
    for (int n = 0; n < lq; ++n) {
      if (!qm[n]) continue;                     // q_tokens[q_mask]
      float acc = 0.f;
      for (int i = lane * 4; i < d; i += 128) {
        const float4 b = load4(p + i);
        const float* a = qs + n * d + i;
        acc = fmaf(a[0], b.x, acc);
        acc = fmaf(a[1], b.y, acc);
        acc = fmaf(a[2], b.z, acc);
        acc = fmaf(a[3], b.w, acc);
      }
      colmax = fmaxf(colmax, warp_sum(acc));
    }
    if (colmax > best) { best = colmax; bidx = m; }   // ascending m per warp: first index kept
  }
  if (lane == 0) { wbest[warp] = best; widx[warp] = bidx; }
  __syncthreads();
  if (tid == 0) {
    float v = wbest[0];
    int ix = widx[0];
    for (int w = 1; w < 8; ++w)
      if (wbest[w] > v || (wbest[w] == v && widx[w] < ix)) { v = wbest[w]; ix = widx[w]; }
    out_idx[blockIdx.x] = (ix == 0x7fffffff) ? 0 : ix;  // no valid page token: torch.argmax of all -inf is 0
  }
}

__global__ void has_page_kernel(const uint8_t* __restrict__ pmask, int np, int lp,
                                uint8_t* __restrict__ has) {
  const int lane = threadIdx.x & 31;
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= np) return;
  int any = 0;
  for (int m = lane; m < lp; m += 32) any |= pmask[(size_t)c * lp + m];
  any = __any_sync(0xffffffffu, any != 0);
  if (lane == 0) has[c] = any ? 1 : 0;
}

template <typename TQ, typename TP>
static int launch_fwd(const void* Q, const void* P, const uint8_t* pmask, int R, int np, int lp,
                      int d, float* rowmax, int32_t* argmax, cudaStream_t st) {
  dim3 grid((R + BM - 1) / BM, np);
  rowmax_kernel<TQ, TP><<<grid, 256, 0, st>>>(static_cast<const TQ*>(Q), static_cast<const TP*>(P),
                                              pmask, R, lp, d, rowmax, argmax);
  EVDR_LAUNCH_CHECK("rowmax_kernel");
  return EVDR_OK;
}

}  // namespace evdr

using namespace evdr;

extern "C" {

int evdr_maxsim_fwd(evdr_handle* h, const void* Q, int q_dtype, const void* P, int p_dtype,
                    const uint8_t* qmask, const uint8_t* pmask, int nq, int np, int lq, int lp,
                    int d, float* rowmax_ws, int32_t* argmax, float* out, void* stream) {
  EVDR_REQUIRE_HANDLE(h);
  EVDR_CHECK_ARG(Q && P && qmask && pmask && rowmax_ws && out, "evdr_maxsim_fwd: null pointer");
  EVDR_CHECK_ARG(nq >= 1 && np >= 1 && lq >= 1 && lp >= 1 && d >= 4 && d % 4 == 0,
                 "evdr_maxsim_fwd: bad shape nq=%d np=%d lq=%d lp=%d d=%d", nq, np, lq, lp, d);
  EVDR_CHECK_ARG((q_dtype == EVDR_F32 || q_dtype == EVDR_BF16) && (p_dtype == EVDR_F32 || p_dtype == EVDR_BF16),
                 "evdr_maxsim_fwd: dtype must be EVDR_F32 or EVDR_BF16");
  EVDR_CHECK_ARG(np <= 65535, "evdr_maxsim_fwd: np=%d > 65535 pages per call (chunk the page axis)", np);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int R = nq * lq;
  int rc;
  if (q_dtype == EVDR_F32 && p_dtype == EVDR_F32)
    rc = launch_fwd<float, float>(Q, P, pmask, R, np, lp, d, rowmax_ws, argmax, st);
  else if (q_dtype == EVDR_F32)
    rc = launch_fwd<float, __nv_bfloat16>(Q, P, pmask, R, np, lp, d, rowmax_ws, argmax, st);
  else if (p_dtype == EVDR_F32)
    rc = launch_fwd<__nv_bfloat16, float>(Q, P, pmask, R, np, lp, d, rowmax_ws, argmax, st);
  else
    rc = launch_fwd<__nv_bfloat16, __nv_bfloat16>(Q, P, pmask, R, np, lp, d, rowmax_ws, argmax, st);
  if (rc != EVDR_OK) return rc;
  rowsum_kernel<<<(np + 7) / 8, 256, 0, st>>>(rowmax_ws, qmask, pmask, nq, np, lq, lp, out);
  EVDR_LAUNCH_CHECK("rowsum_kernel");
  return EVDR_OK;
}

int evdr_hard_token_argmax(evdr_handle* h, const void* Q, int q_dtype, const void* P, int p_dtype,
                           const uint8_t* qmask, const uint8_t* pmask, const int64_t* doc_idx, int nq,
                           int n_docs_per_query, int lq, int lp, int d, int64_t* out_idx, void* stream) {
  EVDR_REQUIRE_HANDLE(h);
  EVDR_CHECK_ARG(Q && P && qmask && pmask && doc_idx && out_idx, "evdr_hard_token_argmax: null pointer");
  EVDR_CHECK_ARG(nq >= 1 && n_docs_per_query >= 1 && lq >= 1 && lp >= 1 && d >= 4 && d % 4 == 0,
                 "evdr_hard_token_argmax: bad shape");
  const size_t smem = sizeof(float) * (size_t)lq * d;
  if (smem > h->smem_optin) return set_error(EVDR_ERR_UNSUPPORTED, "evdr_hard_token_argmax: lq*d too large");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned grid = (unsigned)nq * (unsigned)n_docs_per_query;
  const long long* di = reinterpret_cast<const long long*>(doc_idx);
  long long* oi = reinterpret_cast<long long*>(out_idx);
#define EVDR_HT(TQ, TP)                                                                              \
  do {                                                                                               \
    EVDR_CUDA(cudaFuncSetAttribute(hard_token_kernel<TQ, TP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    hard_token_kernel<TQ, TP><<<grid, 256, smem, st>>>(static_cast<const TQ*>(Q), static_cast<const TP*>(P), qmask, \
                                                        pmask, di, n_docs_per_query, lq, lp, d, oi);  \
  } while (0)
  if (q_dtype == EVDR_F32 && p_dtype == EVDR_F32) EVDR_HT(float, float);
  else if (q_dtype == EVDR_F32 && p_dtype == EVDR_BF16) EVDR_HT(float, __nv_bfloat16);
  else if (q_dtype == EVDR_BF16 && p_dtype == EVDR_F32) EVDR_HT(__nv_bfloat16, float);
  else if (q_dtype == EVDR_BF16 && p_dtype == EVDR_BF16) EVDR_HT(__nv_bfloat16, __nv_bfloat16);
  else return set_error(EVDR_ERR_INVALID, "evdr_hard_token_argmax: dtype must be EVDR_F32 or EVDR_BF16");
#undef EVDR_HT
  EVDR_LAUNCH_CHECK("hard_token_kernel");
  return EVDR_OK;
}

int evdr_slots_to_tokens(evdr_handle* h, const uint8_t* pmask, int32_t* argmax, int np, int n_rows, int lp,
                         void* stream) {
  EVDR_REQUIRE_HANDLE(h);
  EVDR_CHECK_ARG(pmask && argmax && np >= 1 && n_rows >= 1 && lp >= 1, "evdr_slots_to_tokens: bad arguments");
  const long long n = (long long)np * n_rows;
  slot_to_token_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(pmask, argmax, n,
                                                                                                n_rows, lp);
  EVDR_LAUNCH_CHECK("slot_to_token_kernel");
  return EVDR_OK;
}

int evdr_maxsim_rescore(evdr_handle* h, const void* Q, int q_dtype, const void* P, int p_dtype,
                        const uint8_t* qmask, const uint8_t* pmask, int32_t* argmax, int nq, int np,
                        int lq, int lp, int d, float* rowmax_ws, float* out, void* stream) {
  EVDR_REQUIRE_HANDLE(h);
  EVDR_CHECK_ARG(Q && P && qmask && pmask && argmax && rowmax_ws && out, "evdr_maxsim_rescore: null pointer");
  EVDR_CHECK_ARG(nq >= 1 && np >= 1 && lq >= 1 && lp >= 1 && d >= 4 && d % 4 == 0, "evdr_maxsim_rescore: bad shape");
  EVDR_CHECK_ARG(np <= 65535, "evdr_maxsim_rescore: np=%d > 65535 pages per call (chunk the page axis)", np);
  EVDR_CHECK_ARG((q_dtype == EVDR_F32 || q_dtype == EVDR_BF16) && (p_dtype == EVDR_F32 || p_dtype == EVDR_BF16),
                 "evdr_maxsim_rescore: dtype must be EVDR_F32 or EVDR_BF16");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int R = nq * lq;
  int rc;
  if (q_dtype == EVDR_F32 && p_dtype == EVDR_F32)
    rc = launch_rescore<float, float>(Q, P, pmask, argmax, R, np, lp, d, rowmax_ws, st);
  else if (q_dtype == EVDR_F32)
    rc = launch_rescore<float, __nv_bfloat16>(Q, P, pmask, argmax, R, np, lp, d, rowmax_ws, st);
  else if (p_dtype == EVDR_F32)
    rc = launch_rescore<__nv_bfloat16, float>(Q, P, pmask, argmax, R, np, lp, d, rowmax_ws, st);
  else
    rc = launch_rescore<__nv_bfloat16, __nv_bfloat16>(Q, P, pmask, argmax, R, np, lp, d, rowmax_ws, st);
  if (rc != EVDR_OK) return rc;
  rowsum_kernel<<<(np + 7) / 8, 256, 0, st>>>(rowmax_ws, qmask, pmask, nq, np, lq, lp, out);
  EVDR_LAUNCH_CHECK("rowsum_kernel");
  return EVDR_OK;
}

int evdr_maxsim_bwd(evdr_handle* h, const float* gout, const void* Q, int q_dtype, const void* P,
                    int p_dtype, const uint8_t* qmask, const uint8_t* pmask, const int32_t* argmax,
                    int nq, int np, int lq, int lp, int d, float* dP, float* dQ, void* stream) {
  EVDR_REQUIRE_HANDLE(h);
  EVDR_CHECK_ARG(gout && Q && P && qmask && pmask && argmax, "evdr_maxsim_bwd: null pointer");
  EVDR_CHECK_ARG(dP || dQ, "evdr_maxsim_bwd: both dP and dQ are null");
  EVDR_CHECK_ARG(nq >= 1 && np >= 1 && lq >= 1 && lp >= 1 && d >= 4 && d % 4 == 0,
                 "evdr_maxsim_bwd: bad shape");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int R = nq * lq;
  if (dP) {
    size_t smem = sizeof(int) * ((size_t)lp + lp + 1 + R);
    if (smem > h->smem_optin)
      return set_error(EVDR_ERR_UNSUPPORTED,
                       "evdr_maxsim_bwd: nq*lq=%d rows + lp=%d need %zu B of shared memory (> %zu)",
                       R, lp, smem, h->smem_optin);
    if (q_dtype == EVDR_F32) {
      EVDR_CUDA(cudaFuncSetAttribute(bwd_dp_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      bwd_dp_kernel<float><<<np, 256, smem, st>>>(gout, static_cast<const float*>(Q), qmask, pmask,
                                                  argmax, nq, np, lq, lp, d, dP);
    } else {
      EVDR_CUDA(cudaFuncSetAttribute(bwd_dp_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      bwd_dp_kernel<__nv_bfloat16><<<np, 256, smem, st>>>(gout, static_cast<const __nv_bfloat16*>(Q),
                                                          qmask, pmask, argmax, nq, np, lq, lp, d, dP);
    }
    EVDR_LAUNCH_CHECK("bwd_dp_kernel");
  }
  if (dQ) {
    uint8_t* has = nullptr;
    EVDR_CUDA(evdr::ws_alloc(h, &has, np, st));
    has_page_kernel<<<(np + 7) / 8, 256, 0, st>>>(pmask, np, lp, has);
    EVDR_LAUNCH_CHECK("has_page_kernel");
    if (p_dtype == EVDR_F32)
      bwd_dq_kernel<float><<<(R + 7) / 8, 256, 0, st>>>(gout, static_cast<const float*>(P), qmask, pmask, has,
                                                        argmax, nq, np, lq, lp, d, dQ);
    else
      bwd_dq_kernel<__nv_bfloat16><<<(R + 7) / 8, 256, 0, st>>>(
          gout, static_cast<const __nv_bfloat16*>(P), qmask, pmask, has, argmax, nq, np, lq, lp, d, dQ);
    EVDR_LAUNCH_CHECK("bwd_dq_kernel");
    EVDR_CUDA(cudaFreeAsync(has, st));
  }
  return EVDR_OK;
}

}  // extern "C"
