// losses.cu -- the (B, N) score-matrix consumers that sit directly behind the scorer in a training
// step: rank-gap mining and the distillation losses, each as ONE launch that
// also produces the gradient w.r.t. the student scores (the input of evdr_maxsim_bwd).
//
//   rank_gap_kernel      mainv3_iter_liscore_QA_hardtoken.py:368-385 (double argsort -> ranks, gap,
//                        per-page gap mass, teacher top-k, hardest `aux` pages of every query)
//   liscore_kernel       criterion.py:123-142 (listwise_distillation_loss) + :81-83
//                        (score_preserving_loss), the pair every *_liscore recipe combines
//   pairwise_kernel      criterion.py:96-110 (pairwise_distillation_loss) without the three
//                        (B, N, N) intermediates
//
// Every reduction runs in a fixed order (no floating-point atomics except on integer-valued floats
// below 2^24, which are exact), so losses and gradients are bit-reproducible.
#include "common.cuh"

namespace evdr {

constexpr int LOSS_THREADS = 256;

__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
  // red: >= 32 floats of shared memory; result broadcast to every thread
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float x = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, x) : v + x;
  }
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = is_max ? -INFINITY : 0.f;
  for (int w = 0; w < nw; ++w) r = is_max ? fmaxf(r, red[w]) : r + red[w];
  return r;
}

// position of element i in the stable descending order of row x (ties: lower index first)
__device__ __forceinline__ int desc_rank(const float* x, int n, int i) {
  const float v = x[i];
  int r = 0;
  for (int j = 0; j < n; ++j) {
    const float u = x[j];
    r += (u > v) || (u == v && j < i);
  }
  return r;
}

// one CTA per query row
__global__ void __launch_bounds__(LOSS_THREADS)
rank_gap_kernel(const float* __restrict__ sc_t, long long ld_t, const float* __restrict__ sc_s,
                long long ld_s, int n, int kk, int a, float* __restrict__ gap, float* __restrict__ G,
                int32_t* __restrict__ rank_t_out, int32_t* __restrict__ rank_s_out,
                long long* __restrict__ topk_idx, long long* __restrict__ aux_idx) {
  extern __shared__ float sm[];
  float* t = sm;                 // [n]
  float* s = sm + n;             // [n]
  int* cand = reinterpret_cast<int*>(sm + 2 * n);   // [kk] page at teacher rank r
  float* cgap = sm + 2 * n + kk;                    // [kk] |gap| of that page
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    t[i] = sc_t[(size_t)b * ld_t + i];
    s[i] = sc_s[(size_t)b * ld_s + i];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int rt = desc_rank(t, n, i), rs = desc_rank(s, n, i);
    const float g = (float)rt - (float)rs;
    if (gap) gap[(size_t)b * n + i] = g;
    if (G) atomicAdd(&G[i], fabsf(g));        // integer-valued, < 2^24: exact in any order
    if (rank_t_out) rank_t_out[(size_t)b * n + i] = rt;
    if (rank_s_out) rank_s_out[(size_t)b * n + i] = rs;
    if (rt < kk) { cand[rt] = i; cgap[rt] = fabsf(g); }
  }
  __syncthreads();
  for (int r = threadIdx.x; r < kk; r += blockDim.x) {
    if (topk_idx) topk_idx[(size_t)b * kk + r] = cand[r];
    if (aux_idx && a > 0) {
      const int pos = desc_rank(cgap, kk, r);   // largest |gap| first, ties in teacher-rank order
      if (pos < a) aux_idx[(size_t)b * a + pos] = cand[r];
    }
  }
}

// Per-row losses + gradients of listwise distillation (top-k partial cross entropy on temperature
// softmaxes, x T^2) and of the score-preserving MSE.  row_loss[b] = {listwise_b, sum_i (s-t)^2}.
// The last CTA to finish adds the rows up in index order: loss[0] = mean_b listwise_b (x T^2),
// loss[1] = mean_{b,i} (s-t)^2.
__global__ void __launch_bounds__(LOSS_THREADS)
liscore_kernel(const float* __restrict__ sc_s, long long ld_s, const float* __restrict__ sc_t,
               long long ld_t, int B, int n, int k, float T, float* __restrict__ g_list,
               float* __restrict__ g_mse, float* __restrict__ row_loss, float* __restrict__ loss,
               int* __restrict__ done) {
  extern __shared__ float sm[];
  float* t = sm;
  float* s = sm + n;
  __shared__ float red[32];
  __shared__ int last;
  const int b = blockIdx.x;
  float mt = -INFINITY, ms = -INFINITY;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float tv = sc_t[(size_t)b * ld_t + i], sv = sc_s[(size_t)b * ld_s + i];
    t[i] = tv;
    s[i] = sv;
    mt = fmaxf(mt, tv);
    ms = fmaxf(ms, sv);
  }
  mt = block_reduce(mt, red, true);
  ms = block_reduce(ms, red, true);
  const float invT = 1.f / T;
  float zt = 0.f, zs = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    zt += expf((t[i] - mt) * invT);
    zs += expf((s[i] - ms) * invT);
  }
  zt = block_reduce(zt, red, false);
  zs = block_reduce(zs, red, false);
  const float log_zs = logf(zs);
  // teacher top-k membership (criterion.py:132: topk of prob_t == topk of the teacher scores)
  float ce = 0.f, st = 0.f, se = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const bool in_top = desc_rank(t, n, i) < k;
    const float pt = expf((t[i] - mt) * invT) / zt;
    const float lps = (s[i] - ms) * invT - log_zs;
    if (in_top) { ce -= pt * lps; st += pt; }
    const float d = s[i] - t[i];
    se += d * d;
    // park what the gradient pass needs
    g_mse[(size_t)b * n + i] = 2.f * d / ((float)B * (float)n);
    g_list[(size_t)b * n + i] = in_top ? pt : 0.f;
  }
  ce = block_reduce(ce, red, false);
  st = block_reduce(st, red, false);
  se = block_reduce(se, red, false);
  // d/ds_i of  T^2/B * ( -sum_{j in top} pt_j * logp_s_j )  =  T/B * ( p_s_i * sum_top pt - [i in top] pt_i )
  const float scale = T / (float)B;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float ps = expf((s[i] - ms) * invT) / zs;
    const float own = g_list[(size_t)b * n + i];       // written by this same thread above
    g_list[(size_t)b * n + i] = scale * (ps * st - own);
  }
  if (threadIdx.x == 0) {
    row_loss[2 * b] = ce;
    row_loss[2 * b + 1] = se;
    __threadfence();
    last = (atomicAdd(done, 1) == B - 1);
    if (last) *done = 0;                    // hand the counter slot back zeroed (ring slots are shared)
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    float l0 = 0.f, l1 = 0.f;
    const volatile float* rl = row_loss;
    for (int r = 0; r < B; ++r) { l0 += rl[2 * r]; l1 += rl[2 * r + 1]; }
    loss[0] = l0 / (float)B * (T * T);
    loss[1] = l1 / ((float)B * (float)n);
  }
}

// numerically stable pieces of binary_cross_entropy_with_logits(x, y) = max(x,0) - x*y + log1p(exp(-|x|))
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// grid (ceil(n / LOSS_THREADS), B): thread = one page i of query b, loop over all pages j.
//   loss = mean_{b,i,j} bce(s_i - s_j, sigmoid(t_i - t_j));   dloss/ds_i = 2/(B n^2) sum_j (sig(s_i-s_j) - sig(t_i-t_j))
__global__ void __launch_bounds__(LOSS_THREADS)
pairwise_kernel(const float* __restrict__ sc_s, long long ld_s, const float* __restrict__ sc_t,
                long long ld_t, int B, int n, float* __restrict__ grad, float* __restrict__ part,
                float* __restrict__ loss, int* __restrict__ done) {
  extern __shared__ float sm[];
  float* t = sm;
  float* s = sm + n;
  __shared__ float red[32];
  __shared__ int last;
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    t[i] = sc_t[(size_t)b * ld_t + i];
    s[i] = sc_s[(size_t)b * ld_s + i];
  }
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float acc = 0.f, gacc = 0.f;
  if (i < n) {
    const float si = s[i], ti = t[i];
    for (int j = 0; j < n; ++j) {
      const float x = si - s[j];
      const float y = sigmoidf_(ti - t[j]);
      acc += fmaxf(x, 0.f) - x * y + log1pf(expf(-fabsf(x)));
      gacc += sigmoidf_(x) - y;
    }
    grad[(size_t)b * n + i] = gacc * (2.f / ((float)B * (float)n * (float)n));
  }
  acc = block_reduce(acc, red, false);
  const int nblk = gridDim.x * gridDim.y;
  if (threadIdx.x == 0) {
    part[blockIdx.y * gridDim.x + blockIdx.x] = acc;
    __threadfence();
    last = (atomicAdd(done, 1) == nblk - 1);
    if (last) *done = 0;                    // hand the counter slot back zeroed (ring slots are shared)
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    const volatile float* pp = part;
    double tot = 0.0;
    for (int r = 0; r < nblk; ++r) tot += (double)pp[r];
    loss[0] = (float)(tot / ((double)B * (double)n * (double)n));
  }
}

__device__ __forceinline__ float softplusf_(float x) {          // log(1 + exp(x)), stable
  return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x)));
}

// criterion.py:148-189 (lambda_loss).  With p(i) = teacher rank of page i (0 = best) the reference's
// sort + upper-triangle mask is "every pair with p(i) < p(j)":
//   w_ij = 10 * |sig(t_i) - sig(t_j)| * |1/log2(p(i)+2) - 1/log2(p(j)+2)|
//   loss = sum_{b, p(i)<p(j)} w_ij * softplus(-alpha (s_i - s_j)) / (N(N-1)/2 + eps)        (summed over b)
// grid (ceil(n / LOSS_THREADS), B): thread = page i, loop over all j.
__global__ void __launch_bounds__(LOSS_THREADS)
lambda_kernel(const float* __restrict__ sc_s, long long ld_s, const float* __restrict__ sc_t,
              long long ld_t, int B, int n, float alpha, float eps, float* __restrict__ grad,
              float* __restrict__ part, float* __restrict__ loss, int* __restrict__ done) {
  extern __shared__ float sm[];
  float* t = sm;                                   // sigmoid(t) after the rank pass
  float* s = sm + n;
  float* disc = sm + 2 * n;                        // 1 / log2(rank + 2) of page i
  __shared__ float red[32];
  __shared__ int last;
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    t[i] = sc_t[(size_t)b * ld_t + i];
    s[i] = sc_s[(size_t)b * ld_s + i];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) disc[i] = 1.f / log2f((float)desc_rank(t, n, i) + 2.f);
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) t[i] = sigmoidf_(t[i]);
  __syncthreads();
  const float denom = 0.5f * (float)n * (float)(n - 1) + eps;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float acc = 0.f, gacc = 0.f;
  if (i < n) {
    const float si = s[i], ri = t[i], di = disc[i];
    for (int j = 0; j < n; ++j) {
      if (j == i) continue;
      const float dj = disc[j];
      const float w = 10.f * fabsf(ri - t[j]) * fabsf(di - dj);
      // larger discount = better teacher rank (ranks are a permutation: no equal discounts)
      if (di > dj) {                               // i ranked above j: term softplus(-alpha (s_i - s_j))
        const float x = alpha * (si - s[j]);
        acc += w * softplusf_(-x);
        gacc -= w * alpha * sigmoidf_(-x);
      } else {                                     // j ranked above i: term softplus(-alpha (s_j - s_i))
        const float x = alpha * (s[j] - si);
        gacc += w * alpha * sigmoidf_(-x);
      }
    }
    grad[(size_t)b * n + i] = gacc / denom;
  }
  acc = block_reduce(acc, red, false);
  const int nblk = gridDim.x * gridDim.y;
  if (threadIdx.x == 0) {
    part[blockIdx.y * gridDim.x + blockIdx.x] = acc;
    __threadfence();
    last = (atomicAdd(done, 1) == nblk - 1);
    if (last) *done = 0;                    // hand the counter slot back zeroed (ring slots are shared)
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    const volatile float* pp = part;
    double tot = 0.0;
    for (int r = 0; r < nblk; ++r) tot += (double)pp[r];
    loss[0] = (float)(tot / (double)denom);
  }
}

// criterion.py:192-226 (ranknce_loss).  o(0), o(1), ... = pages in descending teacher order:
//   infonce = mean_b ( logsumexp_i(s_i / T) - s_o(0) / T )
//   ranking = mean_{b, k < N-1} sig(t_o(k) - t_o(k+1)) * softplus(-(s_o(k) - s_o(k+1)))
//   loss = infonce + lambda * ranking;   row_loss[b] = {infonce_b, sum_k ranking_bk}
// one CTA per query row.
__global__ void __launch_bounds__(LOSS_THREADS)
ranknce_kernel(const float* __restrict__ sc_s, long long ld_s, const float* __restrict__ sc_t,
               long long ld_t, int B, int n, float T, float lambda_w, float* __restrict__ grad,
               float* __restrict__ row_loss, float* __restrict__ loss, int* __restrict__ done) {
  extern __shared__ float sm[];
  float* t = sm;
  float* s = sm + n;
  int* order = reinterpret_cast<int*>(sm + 2 * n);     // order[p] = page at teacher rank p
  float* g = sm + 3 * n;                                // gradient of the ranking term, per page
  __shared__ float red[32];
  __shared__ int last;
  const int b = blockIdx.x;
  float ms = -INFINITY;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    t[i] = sc_t[(size_t)b * ld_t + i];
    const float sv = sc_s[(size_t)b * ld_s + i];
    s[i] = sv;
    g[i] = 0.f;
    ms = fmaxf(ms, sv);
  }
  ms = block_reduce(ms, red, true);                     // (syncs: t, s, g are visible)
  for (int i = threadIdx.x; i < n; i += blockDim.x) order[desc_rank(t, n, i)] = i;
  const float invT = 1.f / T;
  float zs = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) zs += expf((s[i] - ms) * invT);
  zs = block_reduce(zs, red, false);                    // (syncs: order is visible)
  // adjacent pairs in teacher order; pair k touches pages o(k) and o(k+1): thread k owns pair k and adds
  // to g in two passes (even k, then odd k) so that no two threads touch the same page at once
  float rk = 0.f;
  const float pair_scale = (n > 1) ? lambda_w / ((float)B * (float)(n - 1)) : 0.f;
  for (int par = 0; par < 2; ++par) {
    for (int k = 2 * threadIdx.x + par; k + 1 < n; k += 2 * blockDim.x) {
      const int a = order[k], c = order[k + 1];
      const float w = sigmoidf_(t[a] - t[c]);
      const float x = s[a] - s[c];
      rk += w * softplusf_(-x);
      const float d = -w * sigmoidf_(-x) * pair_scale;  // d/ds_a ; d/ds_c = -d
      g[a] += d;
      g[c] -= d;
    }
    __syncthreads();
  }
  rk = block_reduce(rk, red, false);
  const int top = order[0];
  const float ce_scale = invT / (float)B;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float ps = expf((s[i] - ms) * invT) / zs;
    grad[(size_t)b * n + i] = ce_scale * (ps - (i == top ? 1.f : 0.f)) + g[i];
  }
  if (threadIdx.x == 0) {
    row_loss[2 * b] = (ms * invT + logf(zs)) - s[top] * invT;
    row_loss[2 * b + 1] = rk;
    __threadfence();
    last = (atomicAdd(done, 1) == B - 1);
    if (last) *done = 0;                    // hand the counter slot back zeroed (ring slots are shared)
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    const volatile float* rl = row_loss;
    float l0 = 0.f, l1 = 0.f;
    for (int r = 0; r < B; ++r) { l0 += rl[2 * r]; l1 += rl[2 * r + 1]; }
    l0 /= (float)B;
    l1 = (n > 1) ? l1 / ((float)B * (float)(n - 1)) : 0.f;
    loss[0] = l0 + lambda_w * l1;
    loss[1] = l0;
    loss[2] = l1;
  }
}

static int next_counter(evdr_handle* h, cudaStream_t st, int** out) {
  // slots start zeroed (evdr_create) and every user hands its slot back zeroed, so no memset is needed;
  // the atomic index keeps two host threads that (against the contract) share a handle on distinct slots
  (void)st;
  *out = h->sched_ring + (h->sched_next.fetch_add(1, std::memory_order_relaxed) % EVDR_SCHED_RING);
  return EVDR_OK;
}

}  // namespace evdr

using namespace evdr;

extern "C" int evdr_rank_gap(evdr_handle* h, const float* sc_t, int64_t ld_t, const float* sc_s,
                             int64_t ld_s, int B, int n, int kk, int a, float* gap, float* G,
                             int32_t* rank_t, int32_t* rank_s, int64_t* topk_idx, int64_t* aux_idx,
                             void* stream) {
  EVDR_REQUIRE_HANDLE(h);
  EVDR_CHECK_ARG(sc_t && sc_s, "evdr_rank_gap: null score pointer");
  EVDR_CHECK_ARG(B >= 1 && n >= 1 && ld_t >= n && ld_s >= n, "evdr_rank_gap: bad shape B=%d n=%d", B, n);
  EVDR_CHECK_ARG(kk >= 0 && kk <= n && kk <= 1024 && a >= 0 && a <= kk,
                 "evdr_rank_gap: need 0 <= aux (%d) <= k (%d) <= min(n, 1024)", a, kk);
  EVDR_CHECK_ARG((long long)B * n < (1LL << 24), "evdr_rank_gap: B*n must stay below 2^24 (exact gap mass)");
  const size_t smem = sizeof(float) * (2 * (size_t)n + 2 * (size_t)kk);
  if (smem > h->smem_optin)
    return set_error(EVDR_ERR_UNSUPPORTED, "evdr_rank_gap: n=%d rows do not fit shared memory", n);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (G) EVDR_CUDA(cudaMemsetAsync(G, 0, sizeof(float) * (size_t)n, st));
  EVDR_CUDA(cudaFuncSetAttribute(rank_gap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  rank_gap_kernel<<<B, LOSS_THREADS, smem, st>>>(sc_t, ld_t, sc_s, ld_s, n, kk, a, gap, G, rank_t, rank_s,
                                                 reinterpret_cast<long long*>(topk_idx),
                                                 reinterpret_cast<long long*>(aux_idx));
  EVDR_LAUNCH_CHECK("rank_gap_kernel");
  return EVDR_OK;
}

extern "C" int evdr_liscore_loss(evdr_handle* h, const float* sc_s, int64_t ld_s, const float* sc_t,
                                 int64_t ld_t, int B, int n, int k, float temperature,
                                 float* grad_listwise, float* grad_mse, float* row_loss_ws,
                                 float* loss2, void* stream) {
  EVDR_REQUIRE_HANDLE(h);
  EVDR_CHECK_ARG(sc_s && sc_t && grad_listwise && grad_mse && row_loss_ws && loss2,
                 "evdr_liscore_loss: null pointer");
  EVDR_CHECK_ARG(B >= 1 && n >= 1 && ld_t >= n && ld_s >= n, "evdr_liscore_loss: bad shape B=%d n=%d", B, n);
  // torch.topk raises when k > n (criterion.py:132)
  EVDR_CHECK_ARG(k >= 1 && k <= n, "evdr_liscore_loss: selected index k out of range (k=%d, n=%d)", k, n);
  EVDR_CHECK_ARG(temperature > 0.f, "evdr_liscore_loss: temperature must be positive");
  const size_t smem = sizeof(float) * 2 * (size_t)n;
  if (smem > h->smem_optin)
    return set_error(EVDR_ERR_UNSUPPORTED, "evdr_liscore_loss: n=%d rows do not fit shared memory", n);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int* done = nullptr;
  int rc = next_counter(h, st, &done);
  if (rc != EVDR_OK) return rc;
  EVDR_CUDA(cudaFuncSetAttribute(liscore_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  liscore_kernel<<<B, LOSS_THREADS, smem, st>>>(sc_s, ld_s, sc_t, ld_t, B, n, k, temperature, grad_listwise,
                                                grad_mse, row_loss_ws, loss2, done);
  EVDR_LAUNCH_CHECK("liscore_kernel");
  return EVDR_OK;
}

extern "C" int evdr_pairwise_loss(evdr_handle* h, const float* sc_s, int64_t ld_s, const float* sc_t,
                                  int64_t ld_t, int B, int n, float* grad, float* part_ws, float* loss1,
                                  void* stream) {
  EVDR_REQUIRE_HANDLE(h);
  EVDR_CHECK_ARG(sc_s && sc_t && grad && part_ws && loss1, "evdr_pairwise_loss: null pointer");
  EVDR_CHECK_ARG(B >= 1 && B <= 65535 && n >= 1 && ld_t >= n && ld_s >= n,
                 "evdr_pairwise_loss: bad shape B=%d n=%d", B, n);
  const size_t smem = sizeof(float) * 2 * (size_t)n;
  if (smem > h->smem_optin)
    return set_error(EVDR_ERR_UNSUPPORTED, "evdr_pairwise_loss: n=%d rows do not fit shared memory", n);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int* done = nullptr;
  int rc = next_counter(h, st, &done);
  if (rc != EVDR_OK) return rc;
  EVDR_CUDA(cudaFuncSetAttribute(pairwise_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((n + LOSS_THREADS - 1) / LOSS_THREADS, B);
  pairwise_kernel<<<grid, LOSS_THREADS, smem, st>>>(sc_s, ld_s, sc_t, ld_t, B, n, grad, part_ws, loss1, done);
  EVDR_LAUNCH_CHECK("pairwise_kernel");
  return EVDR_OK;
}

extern "C" int evdr_lambda_loss(evdr_handle* h, const float* sc_s, int64_t ld_s, const float* sc_t, int64_t ld_t,
                                int B, int n, float alpha, float eps, float* grad, float* part_ws, float* loss1,
                                void* stream) {
  EVDR_REQUIRE_HANDLE(h);
  EVDR_CHECK_ARG(sc_s && sc_t && grad && part_ws && loss1, "evdr_lambda_loss: null pointer");
  EVDR_CHECK_ARG(B >= 1 && B <= 65535 && n >= 1 && ld_t >= n && ld_s >= n, "evdr_lambda_loss: bad shape B=%d n=%d", B, n);
  const size_t smem = sizeof(float) * 3 * (size_t)n;
  if (smem > h->smem_optin)
    return set_error(EVDR_ERR_UNSUPPORTED, "evdr_lambda_loss: n=%d rows do not fit shared memory", n);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int* done = nullptr;
  int rc = next_counter(h, st, &done);
  if (rc != EVDR_OK) return rc;
  EVDR_CUDA(cudaFuncSetAttribute(lambda_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((n + LOSS_THREADS - 1) / LOSS_THREADS, B);
  lambda_kernel<<<grid, LOSS_THREADS, smem, st>>>(sc_s, ld_s, sc_t, ld_t, B, n, alpha, eps, grad, part_ws, loss1, done);
  EVDR_LAUNCH_CHECK("lambda_kernel");
  return EVDR_OK;
}

extern "C" int evdr_ranknce_loss(evdr_handle* h, const float* sc_s, int64_t ld_s, const float* sc_t, int64_t ld_t,
                                 int B, int n, float temperature, float lambda_weight, float* grad,
                                 float* row_loss_ws, float* loss3, void* stream) {
  EVDR_REQUIRE_HANDLE(h);
  EVDR_CHECK_ARG(sc_s && sc_t && grad && row_loss_ws && loss3, "evdr_ranknce_loss: null pointer");
  EVDR_CHECK_ARG(B >= 1 && n >= 1 && ld_t >= n && ld_s >= n, "evdr_ranknce_loss: bad shape B=%d n=%d", B, n);
  EVDR_CHECK_ARG(temperature > 0.f, "evdr_ranknce_loss: temperature must be positive");
  const size_t smem = sizeof(float) * 4 * (size_t)n;
  if (smem > h->smem_optin)
    return set_error(EVDR_ERR_UNSUPPORTED, "evdr_ranknce_loss: n=%d rows do not fit shared memory", n);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int* done = nullptr;
  int rc = next_counter(h, st, &done);
  if (rc != EVDR_OK) return rc;
  EVDR_CUDA(cudaFuncSetAttribute(ranknce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ranknce_kernel<<<B, LOSS_THREADS, smem, st>>>(sc_s, ld_s, sc_t, ld_t, B, n, temperature, lambda_weight, grad,
                                                row_loss_ws, loss3, done);
  EVDR_LAUNCH_CHECK("ranknce_kernel");
  return EVDR_OK;
}
