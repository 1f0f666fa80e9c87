// topk.cu -- evaluation tail on device: warp-level top-k selection, candidate
// merge, and trec_eval-style measures.
//
// Replaces the reference's Nq*Np `.item()` dict build (mainv2_iter_listwise.py:312-318)
// plus mteb/pytrec_eval behind evaluator/retrieval.py:239-246.
//
// Ordering key everywhere: score descending, then tie key ascending, then position
// ascending (only to separate padding sentinels).  The tie key is tie_rank[col] when
// given (host passes the rank of the docid string in descending order = trec_eval's
// rule) else the column index / candidate id.
#include "common.cuh"
#include "topk_dev.cuh"

namespace evdr {

constexpr int TOPK_MAX = 128;
constexpr int TOPK_WARPS = 8;

__device__ __forceinline__ bool better(float s, long long t, float os, long long ot) {
  return (s > os) || (s == os && t < ot);
}

// Per-warp sorted list in shared memory (best first), k entries.
struct WarpList {
  float* s;
  long long* tie;
  long long* id;
};

// Insert candidate (cs, ct, cid) held by all lanes (already known to beat the worst).
__device__ __forceinline__ void warp_insert(const WarpList& L, int k, int lane, float cs,
                                            long long ct, long long cid) {
  // position = number of entries better than the candidate
  int pos = 0;
  for (int j0 = 0; j0 < k; j0 += 32) {
    int j = j0 + lane;
    bool b = (j < k) && better(L.s[j], L.tie[j], cs, ct);
    pos += __popc(__ballot_sync(0xffffffffu, b));
  }
  // shift [pos, k-2] down by one, from the tail, 32 entries at a time
  for (int hi = k - 1; hi > pos; hi -= 32) {
    int j = hi - lane;  // destination index
    float ts = 0.f; long long tt = 0, ti = 0;
    bool mv = (j > pos);
    if (mv) { ts = L.s[j - 1]; tt = L.tie[j - 1]; ti = L.id[j - 1]; }
    __syncwarp();
    if (mv) { L.s[j] = ts; L.tie[j] = tt; L.id[j] = ti; }
    __syncwarp();
  }
  if (lane == 0) { L.s[pos] = cs; L.tie[pos] = ct; L.id[pos] = cid; }
  __syncwarp();
}

// trec_eval-style measures of ONE query from its ranked ids; double accumulators like trec_eval.
// out[m][j][q], m: 0 ndcg, 1 map, 2 recall, 3 precision, 4 mrr
struct MetricArgs {
  const int32_t* qrel_off;
  const long long* qrel_ids;
  const int32_t* qrel_rels;
  const int32_t* k_values;
  int nk;
  double* out;      // nullptr: no measures
};

__device__ void query_metrics(const long long* __restrict__ ids, int kmax, int q, int nq,
                              const MetricArgs& ma) {
  const int b = ma.qrel_off[q], e = ma.qrel_off[q + 1];
  int n_rel = 0;
  for (int i = b; i < e; ++i) n_rel += ma.qrel_rels[i] > 0;
  for (int j = 0; j < ma.nk; ++j) {
    const int k = ma.k_values[j];
    const int kk = min(k, kmax);
    double dcg = 0.0, ap = 0.0, rr = 0.0;
    int hits = 0;
    for (int i = 0; i < kk; ++i) {
      long long id = ids[i];
      int rel = 0;
      if (id >= 0)
        for (int t = b; t < e; ++t)
          if (ma.qrel_ids[t] == id) { rel = ma.qrel_rels[t]; break; }
      if (rel > 0) {
        dcg += (double)rel / log2((double)i + 2.0);
        ++hits;
        ap += (double)hits / (double)(i + 1);
        if (rr == 0.0) rr = 1.0 / (double)(i + 1);
      }
    }
    // ideal DCG: positive rels in descending order, cut at k (selection without storage)
    double idcg = 0.0;
    {
      int last_rel = 0x7fffffff;  // rels emitted so far are >= last_rel
      int emitted = 0;
      while (emitted < k && emitted < n_rel) {
        int best = 0, cnt = 0;
        for (int t = b; t < e; ++t) {
          int r = ma.qrel_rels[t];
          if (r > 0 && r < last_rel) {
            if (r > best) { best = r; cnt = 1; }
            else if (r == best) ++cnt;
          }
        }
        if (best == 0) break;
        for (int c = 0; c < cnt && emitted < k; ++c, ++emitted)
          idcg += (double)best / log2((double)emitted + 2.0);
        last_rel = best;
      }
    }
    const size_t o = (size_t)j * nq + q;
    const size_t plane = (size_t)ma.nk * nq;
    ma.out[0 * plane + o] = idcg > 0.0 ? dcg / idcg : 0.0;
    ma.out[1 * plane + o] = n_rel > 0 ? ap / (double)n_rel : 0.0;
    ma.out[2 * plane + o] = n_rel > 0 ? (double)hits / (double)n_rel : 0.0;
    ma.out[3 * plane + o] = (double)hits / (double)k;
    ma.out[4 * plane + o] = rr;
  }
}

// grid (nsplit, nq).  Source is either implicit columns of a score row
// (cand_ids == nullptr: id = col + id_offset, tie = tie_rank ? tie_rank[col] : col)
// or explicit candidates (tie = cand_tie ? cand_tie[i] : cand_ids[i]).
__global__ void __launch_bounds__(TOPK_WARPS * 32)
topk_select_kernel(const float* __restrict__ scores, long long ld, int n,
                   const int32_t* __restrict__ tie_rank, long long id_offset,
                   const long long* __restrict__ cand_ids, const long long* __restrict__ cand_tie,
                   int k, float* __restrict__ out_s, long long* __restrict__ out_tie,
                   long long* __restrict__ out_id, const MetricArgs ma) {
  extern __shared__ unsigned char smem_raw[];
  // layout: per warp k floats, k tie, k id
  float* ws = reinterpret_cast<float*>(smem_raw);
  long long* wt = reinterpret_cast<long long*>(smem_raw + sizeof(float) * TOPK_WARPS * TOPK_MAX);
  long long* wi = wt + TOPK_WARPS * TOPK_MAX;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int q = blockIdx.y;
  const int nsplit = gridDim.x;
  const int seg = (n + nsplit - 1) / nsplit;
  const int beg = blockIdx.x * seg;
  const int end = min(n, beg + seg);
  const float* row = scores + (size_t)q * ld;
  const long long* idrow = cand_ids ? cand_ids + (size_t)q * ld : nullptr;
  const long long* tierow = cand_tie ? cand_tie + (size_t)q * ld : nullptr;

  WarpList L{ws + warp * TOPK_MAX, wt + warp * TOPK_MAX, wi + warp * TOPK_MAX};
  for (int j = lane; j < k; j += 32) {
    L.s[j] = -INFINITY;
    L.tie[j] = 0x7fffffffffffffffLL;
    L.id[j] = -1;
  }
  __syncwarp();
  float thr_s = -INFINITY;
  long long thr_t = 0x7fffffffffffffffLL;

  for (int base = beg + warp * 32; base < end; base += TOPK_WARPS * 32) {
    int i = base + lane;
    float s = -INFINITY;
    long long t = 0x7fffffffffffffffLL, id = -1;
    bool ok = false;
    if (i < end) {
      s = row[i];
      if (idrow) {
        id = idrow[i];
        t = tierow ? tierow[i] : id;
        ok = id >= 0;
      } else {
        id = (long long)i + id_offset;
        t = tie_rank ? (long long)tie_rank[i] : (long long)i;
        ok = true;
      }
      ok = ok && better(s, t, thr_s, thr_t);
    }
    unsigned m = __ballot_sync(0xffffffffu, ok);
    while (m) {
      int src = __ffs(m) - 1;
      m &= m - 1;
      float cs = __shfl_sync(0xffffffffu, s, src);
      long long ct = __shfl_sync(0xffffffffu, t, src);
      long long cid = __shfl_sync(0xffffffffu, id, src);
      if (better(cs, ct, thr_s, thr_t)) {  // threshold may have moved
        warp_insert(L, k, lane, cs, ct, cid);
        thr_s = L.s[k - 1];
        thr_t = L.tie[k - 1];
      }
    }
  }
  __syncthreads();
  // block merge: rank every entry of the 8 lists among all of them
  const int total = TOPK_WARPS * k;
  float* os = out_s + ((size_t)q * nsplit + blockIdx.x) * k;
  long long* ot = out_tie ? out_tie + ((size_t)q * nsplit + blockIdx.x) * k : nullptr;
  long long* oi = out_id + ((size_t)q * nsplit + blockIdx.x) * k;
  for (int e = tid; e < total; e += blockDim.x) {
    int w = e / k, j = e % k;
    float s = ws[w * TOPK_MAX + j];
    long long t = wt[w * TOPK_MAX + j];
    int rank = 0;
    for (int f = 0; f < total; ++f) {
      int fw = f / k, fj = f % k;
      float fs = ws[fw * TOPK_MAX + fj];
      long long ft = wt[fw * TOPK_MAX + fj];
      if (better(fs, ft, s, t) || (fs == s && ft == t && f < e)) ++rank;
    }
    if (rank < k) {
      os[rank] = s;
      if (ot) ot[rank] = t;
      oi[rank] = wi[w * TOPK_MAX + j];
    }
  }
  // fused tail: the CTA that produced a query's final ranking also scores it (nDCG/MAP/Recall/P/MRR)
  if (ma.out != nullptr) {
    __syncthreads();
    if (tid == 0) query_metrics(oi, k, q, gridDim.y, ma);
  }
}

// ---------------------------------------------------------------------------------------------
// k <= 32: the per-warp list lives in REGISTERS (lane j = j-th best so far), an insertion is one ballot
// and three shuffle-ups instead of a shared-memory shift, four independent 32-column loads are in flight
// per warp, and the threshold is bootstrapped from the lane maxima of the first 128 columns (their k-th
// largest is a lower bound of the final k-th value), which removes the insertion storm of a cold list --
// the whole cost when a warp only sees a few hundred columns (single-query search).
// Same sources, outputs and ordering as topk_select_kernel.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_kth_largest(float v, int k, int lane) {
  // bitonic sort of one value per lane, descending; returns the value at position k - 1
#pragma unroll
  for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      const float o = __shfl_xor_sync(0xffffffffu, v, stride);
      const bool desc = (lane & size) == 0;              // direction of this lane's block
      const bool lower = (lane & stride) == 0;           // lower index of the compared pair
      const bool keep_max = (desc == lower);
      v = keep_max ? fmaxf(v, o) : fminf(v, o);
    }
  }
  return __shfl_sync(0xffffffffu, v, k - 1);
}

__global__ void __launch_bounds__(TOPK_WARPS * 32)
topk_small_kernel(const float* __restrict__ scores, long long ld, int n,
                  const int32_t* __restrict__ tie_rank, long long id_offset,
                  const long long* __restrict__ cand_ids, const long long* __restrict__ cand_tie,
                  int k, float* __restrict__ out_s, long long* __restrict__ out_tie,
                  long long* __restrict__ out_id, const MetricArgs ma) {
  __shared__ float ws[TOPK_WARPS * 32];
  __shared__ long long wt[TOPK_WARPS * 32];
  __shared__ long long wi[TOPK_WARPS * 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int q = blockIdx.y;
  const int nsplit = gridDim.x;
  const int seg = (n + nsplit - 1) / nsplit;
  const int beg = blockIdx.x * seg;
  const int end = min(n, beg + seg);
  const float* row = scores + (size_t)q * ld;
  const long long* idrow = cand_ids ? cand_ids + (size_t)q * ld : nullptr;
  const long long* tierow = cand_tie ? cand_tie + (size_t)q * ld : nullptr;
  constexpr long long TIE_MAX = 0x7fffffffffffffffLL;

  // this warp's contiguous slice of the CTA's range (coalesced 128-byte rows per load)
  const int wseg = ((end - beg + TOPK_WARPS - 1) / TOPK_WARPS + 31) & ~31;
  const int wbeg = min(end, beg + warp * wseg);
  const int wend = min(end, wbeg + wseg);

  float my_s = -INFINITY;                    // list entry `lane`
  long long my_t = TIE_MAX, my_i = -1;
  // bootstrap: k-th largest of the lane maxima over the first 128 columns of the slice
  float thr_s;
  long long thr_t = TIE_MAX;
  {
    float m = -INFINITY;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = wbeg + u * 32 + lane;
      if (i < wend && (!idrow || idrow[i] >= 0)) m = fmaxf(m, row[i]);
    }
    thr_s = warp_kth_largest(m, k, lane);
  }
  int filled = 0;                            // insertions so far (the list's k-th entry is real once >= k)
  // software pipeline: the next 128 columns are in flight while the current 128 are filtered
  float nxt[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int i = wbeg + u * 32 + lane;
    nxt[u] = (i < wend) ? __ldcs(row + i) : -INFINITY;
  }
  for (int base = wbeg; base < wend; base += 128) {
    float s[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) s[u] = nxt[u];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = base + 128 + u * 32 + lane;
      nxt[u] = (i < wend) ? __ldcs(row + i) : -INFINITY;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = base + u * 32 + lane;
      long long t = TIE_MAX, id = -1;
      bool ok = (i < wend) && (s[u] >= thr_s);          // cheap pre-test before touching tie / id arrays
      if (ok) {
        if (idrow) {
          id = idrow[i];
          t = tierow ? tierow[i] : id;
          ok = id >= 0;
        } else {
          id = (long long)i + id_offset;
          t = tie_rank ? (long long)tie_rank[i] : (long long)i;
        }
        ok = ok && better(s[u], t, thr_s, thr_t);
      }
      unsigned m = __ballot_sync(0xffffffffu, ok);
      while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        const float cs = __shfl_sync(0xffffffffu, s[u], src);
        const long long ct = __shfl_sync(0xffffffffu, t, src);
        const long long cid = __shfl_sync(0xffffffffu, id, src);
        if (!better(cs, ct, thr_s, thr_t)) continue;    // the threshold may have moved
        const int pos = __popc(__ballot_sync(0xffffffffu, better(my_s, my_t, cs, ct)));
        const float us = __shfl_up_sync(0xffffffffu, my_s, 1);
        const long long ut = __shfl_up_sync(0xffffffffu, my_t, 1);
        const long long ui = __shfl_up_sync(0xffffffffu, my_i, 1);
        if (lane == pos) { my_s = cs; my_t = ct; my_i = cid; }
        else if (lane > pos) { my_s = us; my_t = ut; my_i = ui; }
        if (++filled >= k) {                             // k real entries: the list's k-th is the bar
          const float ls = __shfl_sync(0xffffffffu, my_s, k - 1);
          const long long lt = __shfl_sync(0xffffffffu, my_t, k - 1);
          if (better(ls, lt, thr_s, thr_t)) { thr_s = ls; thr_t = lt; }
        }
      }
    }
  }
  ws[warp * 32 + lane] = my_s;
  wt[warp * 32 + lane] = my_t;
  wi[warp * 32 + lane] = my_i;
  __syncthreads();
  // block merge: rank every entry of the 8 lists (first k of each) among all of them
  const int total = TOPK_WARPS * k;
  float* os = out_s + ((size_t)q * nsplit + blockIdx.x) * k;
  long long* ot = out_tie ? out_tie + ((size_t)q * nsplit + blockIdx.x) * k : nullptr;
  long long* oi = out_id + ((size_t)q * nsplit + blockIdx.x) * k;
  for (int e = tid; e < total; e += blockDim.x) {
    const int w = e / k, j = e % k;
    const float es = ws[w * 32 + j];
    const long long et = wt[w * 32 + j];
    int rank = 0;
    for (int f = 0; f < total; ++f) {
      const int fw = f / k, fj = f % k;
      const float fs = ws[fw * 32 + fj];
      const long long ft = wt[fw * 32 + fj];
      if (better(fs, ft, es, et) || (fs == es && ft == et && f < e)) ++rank;
    }
    if (rank < k) {
      os[rank] = es;
      if (ot) ot[rank] = et;
      oi[rank] = wi[w * 32 + j];
    }
  }
  if (ma.out != nullptr) {
    __syncthreads();
    if (tid == 0) query_metrics(oi, k, q, gridDim.y, ma);
  }
}

__global__ void __launch_bounds__(128)
metrics_kernel(const long long* __restrict__ top_ids, int kmax, int nq, const MetricArgs ma) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  query_metrics(top_ids + (size_t)q * kmax, kmax, q, nq, ma);
}

static int pick_nsplit(const evdr_handle* h, int nq, int n) {
  int want = (2 * h->sm_count + nq - 1) / nq;          // fill the GPU twice over
  // >= 1024 columns per CTA: a warp's scan is a chain of dependent 32-column loads, so a single-query
  // search over 100k pages wants ~100 short chains rather than 25 long ones (the second stage merges
  // nsplit * k candidates in one CTA either way)
  int cap = (n + 1023) / 1024;
  int ns = want < cap ? want : cap;
  return ns < 1 ? 1 : ns;
}

static size_t topk_smem() {
  return (sizeof(float) + 2 * sizeof(long long)) * (size_t)TOPK_WARPS * TOPK_MAX;
}

// two-stage selection shared by evdr_topk / evdr_topk_merge
static int run_topk(evdr_handle* h, const float* scores, long long ld, int nq, int n,
                    const int32_t* tie_rank, long long id_offset, const long long* cand_ids,
                    int k, float* top_s, long long* top_id, cudaStream_t st,
                    const MetricArgs* metrics = nullptr) {
  const MetricArgs none = {nullptr, nullptr, nullptr, nullptr, 0, nullptr};
  const MetricArgs fin = metrics ? *metrics : none;
  const size_t smem = topk_smem();
  EVDR_CUDA(cudaFuncSetAttribute(topk_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int nsplit = pick_nsplit(h, nq, n);
  const bool small = k <= 32;                 // register-list kernel
  if (nsplit == 1) {
    if (small)
      topk_small_kernel<<<dim3(1, nq), TOPK_WARPS * 32, 0, st>>>(
          scores, ld, n, tie_rank, id_offset, cand_ids, nullptr, k, top_s, nullptr, top_id, fin);
    else
      topk_select_kernel<<<dim3(1, nq), TOPK_WARPS * 32, smem, st>>>(
          scores, ld, n, tie_rank, id_offset, cand_ids, nullptr, k, top_s, nullptr, top_id, fin);
    EVDR_LAUNCH_CHECK("topk_select_kernel");
    return EVDR_OK;
  }
  const size_t per = (size_t)nq * nsplit * k;
  unsigned char* ws = nullptr;
  EVDR_CUDA(evdr::ws_alloc(h, &ws, per * (sizeof(float) + 2 * sizeof(long long)), st));
  long long* p_tie = reinterpret_cast<long long*>(ws);
  long long* p_id = p_tie + per;
  float* p_s = reinterpret_cast<float*>(p_id + per);
  if (small) {
    topk_small_kernel<<<dim3(nsplit, nq), TOPK_WARPS * 32, 0, st>>>(
        scores, ld, n, tie_rank, id_offset, cand_ids, nullptr, k, p_s, p_tie, p_id, none);
    EVDR_LAUNCH_CHECK("topk_small_kernel(stage1)");
    topk_small_kernel<<<dim3(1, nq), TOPK_WARPS * 32, 0, st>>>(
        p_s, (long long)nsplit * k, nsplit * k, nullptr, 0, p_id, p_tie, k, top_s, nullptr, top_id, fin);
    EVDR_LAUNCH_CHECK("topk_small_kernel(stage2)");
  } else {
    topk_select_kernel<<<dim3(nsplit, nq), TOPK_WARPS * 32, smem, st>>>(
        scores, ld, n, tie_rank, id_offset, cand_ids, nullptr, k, p_s, p_tie, p_id, none);
    EVDR_LAUNCH_CHECK("topk_select_kernel(stage1)");
    topk_select_kernel<<<dim3(1, nq), TOPK_WARPS * 32, smem, st>>>(
        p_s, (long long)nsplit * k, nsplit * k, nullptr, 0, p_id, p_tie, k, top_s, nullptr, top_id, fin);
    EVDR_LAUNCH_CHECK("topk_select_kernel(stage2)");
  }
  EVDR_CUDA(cudaFreeAsync(ws, st));
  return EVDR_OK;
}

// k <= 32 best of explicit candidate rows: one CTA per query, register lists per warp + rank merge
__global__ void __launch_bounds__(SEL_WARPS * 32)
select_candidates_kernel(const float* __restrict__ cand_s, const long long* __restrict__ cand_i, int ncand, int k,
                         float* __restrict__ top_s, long long* __restrict__ top_i) {
  __shared__ float ws[SEL_WARPS * 32], sel_s[32];
  __shared__ long long wi[SEL_WARPS * 32], sel_i[32];
  const int q = blockIdx.x;
  block_select_topk(cand_s + (size_t)q * ncand, cand_i + (size_t)q * ncand, ncand, k, ws, wi, sel_s, sel_i);
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    top_s[(size_t)q * k + j] = sel_s[j];
    top_i[(size_t)q * k + j] = sel_i[j];
  }
}

int select_candidates(evdr_handle* h, const float* cand_s, const long long* cand_i, int nq, int ncand, int k,
                      float* top_s, long long* top_i, cudaStream_t st) {
  (void)h;
  EVDR_CHECK_ARG(cand_s && cand_i && top_s && top_i && nq >= 1 && ncand >= 1 && k >= 1 && k <= 32,
                 "select_candidates: bad arguments (k=%d must be <= 32)", k);
  select_candidates_kernel<<<nq, SEL_WARPS * 32, 0, st>>>(cand_s, cand_i, ncand, k, top_s, top_i);
  EVDR_LAUNCH_CHECK("select_candidates_kernel");
  return EVDR_OK;
}

}  // namespace evdr

using namespace evdr;

extern "C" {

int evdr_topk(evdr_handle* h, const float* scores, int64_t ld, int nq, int np, int k,
              const int32_t* tie_rank, int64_t id_offset, float* top_scores, int64_t* top_ids,
              void* stream) {
  EVDR_REQUIRE_HANDLE(h);
  EVDR_CHECK_ARG(scores && top_scores && top_ids, "evdr_topk: null pointer");
  EVDR_CHECK_ARG(nq >= 1 && np >= 1 && ld >= np, "evdr_topk: bad shape nq=%d np=%d ld=%lld", nq, np, (long long)ld);
  EVDR_CHECK_ARG(k >= 1 && k <= TOPK_MAX, "evdr_topk: k=%d outside [1,%d]", k, TOPK_MAX);
  EVDR_CHECK_ARG(nq <= 65535, "evdr_topk: nq=%d > 65535 per call", nq);
  return run_topk(h, scores, ld, nq, np, tie_rank, id_offset, nullptr, k, top_scores,
                  reinterpret_cast<long long*>(top_ids), static_cast<cudaStream_t>(stream));
}

int evdr_topk_merge(evdr_handle* h, const float* cand_scores, const int64_t* cand_ids, int nq,
                    int ncand, int k, float* top_scores, int64_t* top_ids, void* stream) {
  EVDR_REQUIRE_HANDLE(h);
  EVDR_CHECK_ARG(cand_scores && cand_ids && top_scores && top_ids, "evdr_topk_merge: null pointer");
  EVDR_CHECK_ARG(nq >= 1 && ncand >= 1, "evdr_topk_merge: bad shape");
  EVDR_CHECK_ARG(k >= 1 && k <= TOPK_MAX, "evdr_topk_merge: k=%d outside [1,%d]", k, TOPK_MAX);
  EVDR_CHECK_ARG(nq <= 65535, "evdr_topk_merge: nq=%d > 65535 per call", nq);
  return run_topk(h, cand_scores, ncand, nq, ncand, nullptr, 0,
                  reinterpret_cast<const long long*>(cand_ids), k, top_scores,
                  reinterpret_cast<long long*>(top_ids), static_cast<cudaStream_t>(stream));
}

int evdr_retrieval_metrics(evdr_handle* h, const int64_t* top_ids, int kmax, int nq,
                           const int32_t* qrel_off, const int64_t* qrel_ids,
                           const int32_t* qrel_rels, const int32_t* k_values, int nk, double* out,
                           void* stream) {
  EVDR_REQUIRE_HANDLE(h);
  EVDR_CHECK_ARG(top_ids && qrel_off && qrel_ids && qrel_rels && k_values && out,
                 "evdr_retrieval_metrics: null pointer");
  EVDR_CHECK_ARG(kmax >= 1 && nq >= 1 && nk >= 1, "evdr_retrieval_metrics: bad shape");
  const MetricArgs ma = {qrel_off, reinterpret_cast<const long long*>(qrel_ids), qrel_rels, k_values, nk, out};
  metrics_kernel<<<(nq + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const long long*>(top_ids), kmax, nq, ma);
  EVDR_LAUNCH_CHECK("metrics_kernel");
  return EVDR_OK;
}

int evdr_topk_metrics(evdr_handle* h, const float* scores, int64_t ld, int nq, int np, int kmax,
                      const int32_t* tie_rank, int64_t id_offset, const int32_t* qrel_off,
                      const int64_t* qrel_ids, const int32_t* qrel_rels, const int32_t* k_values,
                      int nk, float* top_scores, int64_t* top_ids, double* out, void* stream) {
  EVDR_REQUIRE_HANDLE(h);
  EVDR_CHECK_ARG(scores && top_scores && top_ids && qrel_off && qrel_ids && qrel_rels && k_values && out,
                 "evdr_topk_metrics: null pointer");
  EVDR_CHECK_ARG(nq >= 1 && np >= 1 && ld >= np && nk >= 1, "evdr_topk_metrics: bad shape");
  EVDR_CHECK_ARG(kmax >= 1 && kmax <= TOPK_MAX, "evdr_topk_metrics: kmax=%d outside [1,%d]", kmax, TOPK_MAX);
  EVDR_CHECK_ARG(nq <= 65535, "evdr_topk_metrics: nq=%d > 65535 per call", nq);
  const MetricArgs ma = {qrel_off, reinterpret_cast<const long long*>(qrel_ids), qrel_rels, k_values, nk, out};
  return run_topk(h, scores, ld, nq, np, tie_rank, id_offset, nullptr, kmax, top_scores,
                  reinterpret_cast<long long*>(top_ids), static_cast<cudaStream_t>(stream), &ma);
}

}  // extern "C"
