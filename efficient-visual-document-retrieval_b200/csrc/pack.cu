// pack.cu -- HBM layout producers for the tensor-core path, and the fused
// mask * l2_normalize (A4).
//
//   pages  : (np, lp, 128) fp32|bf16 + arbitrary uint8 mask  ->  bf16 token stream in
//            16-token chunks, masked slots replaced by a duplicate valid token
//   queries: (nq, lq<=32, 128) -> bf16 (nq, 32, 128) zero-padded + uint8 (nq, 32)
//
// Reference layout contract: utils/preprocess_data.py:37-105 (zero right padding,
// mask = valid & attn & img); normalisation utils/preprocess_data.py:8-9.
#include "common.cuh"

namespace evdr {

constexpr int CH = EVDR_CHUNK;

__device__ __forceinline__ uint2 pack_bf16x4(float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
  __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
  uint2 o;
  o.x = *reinterpret_cast<uint32_t*>(&a);
  o.y = *reinterpret_cast<uint32_t*>(&b);
  return o;
}
// v - float(bf16(v)): the part of v the bf16 rounding dropped (exact in fp32)
__device__ __forceinline__ float4 residual_bf16x4(float4 v, uint2 hi) {
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hi.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hi.y));
  return make_float4(v.x - a.x, v.y - a.y, v.z - b.x, v.w - b.y);
}

// n_chunks[c]; varlen: ceil((last valid + 1)/16), min 1; else ceil(lp/16)
__global__ void __launch_bounds__(256)
plan_len_kernel(const uint8_t* __restrict__ pmask, const long long* __restrict__ in_off, int np, int lp,
                int varlen, int32_t* __restrict__ nch) {
  const int lane = threadIdx.x & 31;
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= np) return;
  // ragged input (in_off != null): page c owns tokens in_off[c] .. in_off[c+1] of a flat array
  const long long base = in_off ? in_off[c] : (long long)c * lp;
  if (in_off) lp = (int)(in_off[c + 1] - base);
  int n;
  if (!varlen) {
    n = (lp + CH - 1) / CH;
  } else {
    int last = -1;
    for (int m = lane; m < lp; m += 32)
      if (pmask[base + m]) last = m;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
    n = (last + 1 + CH - 1) / CH;
    if (n < 1) n = 1;
  }
  if (lane == 0) nch[c] = n;
}

// single-CTA exclusive scan: off[0..np], in place safe (nch may alias off + 1? no: separate)
__global__ void __launch_bounds__(1024)
plan_scan_kernel(const int32_t* __restrict__ nch, int np, int32_t* __restrict__ off) {
  __shared__ int warp_tot[32];
  __shared__ int carry_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < np; base += 1024) {
    int i = base + tid;
    int v = (i < np) ? nch[i] : 0;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      int w = warp_tot[lane];
      int winc = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc += t;
      }
      warp_tot[lane] = winc - w;  // exclusive
    }
    __syncthreads();
    int carry = carry_s;
    int excl = carry + warp_tot[warp] + inc - v;
    if (i < np) off[i] = excl;
    __syncthreads();
    if (tid == 1023) carry_s = excl + v;
    __syncthreads();
  }
  if (tid == 0) off[np] = carry_s;
}

// One CTA per page.  src[m] = nearest valid index <= m, else first valid index,
// else -1 (page without valid tokens -> zero tokens).
template <typename T>
__global__ void __launch_bounds__(256)
pack_pages_kernel(const T* __restrict__ P, const uint8_t* __restrict__ pmask,
                  const long long* __restrict__ in_off, const int32_t* __restrict__ chunk_off, int np,
                  int lp, int normalize, int kb, __nv_bfloat16* __restrict__ tokens) {
  extern __shared__ int src[];  // n_slots ints
  __shared__ int seg_last[256];
  __shared__ int first_valid_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c = blockIdx.x;
  const int c0 = chunk_off[c];
  const int n_slots = (chunk_off[c + 1] - c0) * CH;
  const long long base = in_off ? in_off[c] : (long long)c * lp;   // ragged input: flat token array
  if (in_off) lp = (int)(in_off[c + 1] - base);
  const uint8_t* mc = pmask + base;

  // segmented "last valid index" scan: thread t owns slots [t*seg, (t+1)*seg)
  const int seg = (n_slots + 255) / 256;
  int my_last = -1, my_first = 0x7fffffff;
  for (int m = tid * seg; m < min((tid + 1) * seg, n_slots); ++m) {
    if (m < lp && mc[m]) { my_last = m; if (my_first == 0x7fffffff) my_first = m; }
  }
  seg_last[tid] = my_last;
  // block min of first valid
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) my_first = min(my_first, __shfl_xor_sync(0xffffffffu, my_first, o));
  if (tid == 0) first_valid_s = 0x7fffffff;
  __syncthreads();
  if (lane == 0 && my_first != 0x7fffffff) atomicMin(&first_valid_s, my_first);
  __syncthreads();
  // inclusive prefix max over seg_last (Hillis-Steele, 256 entries)
  for (int o = 1; o < 256; o <<= 1) {
    int v = seg_last[tid];
    int u = (tid >= o) ? seg_last[tid - o] : -1;
    __syncthreads();
    seg_last[tid] = max(v, u);
    __syncthreads();
  }
  // valid tokens beyond the page's slot range can exist only when lp > n_slots, which the
  // plan excludes (varlen: n_slots covers the last valid token; rect: n_slots >= lp).
  const int first_valid = first_valid_s;
  int run = (tid > 0) ? seg_last[tid - 1] : -1;
  for (int m = tid * seg; m < min((tid + 1) * seg, n_slots); ++m) {
    if (m < lp && mc[m]) run = m;
    src[m] = (run >= 0) ? run : (first_valid != 0x7fffffff ? first_valid : -1);
  }
  __syncthreads();

  const int row = EVDR_TC_DIM * (kb == 3 ? 2 : 1);   // kb = 3 (bf16x3 split): [hi | lo] rows of 256
  __nv_bfloat16* dst = tokens + (size_t)c0 * CH * row;
  const T* Pc = P + base * EVDR_TC_DIM;
  for (int m = warp; m < n_slots; m += (blockDim.x >> 5)) {
    const int s = src[m];
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (s >= 0) {
      v = load4(Pc + (size_t)s * EVDR_TC_DIM + lane * 4);
      if (normalize) {
        float ss = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        ss = warp_sum(ss);
        float inv = 1.f / (sqrtf(ss) + 1e-12f);
        v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
      }
    }
    const uint2 hi = pack_bf16x4(v);
    *reinterpret_cast<uint2*>(dst + (size_t)m * row + lane * 4) = hi;
    if (kb == 3) {
      const uint2 lo = pack_bf16x4(residual_bf16x4(v, hi));
      *reinterpret_cast<uint2*>(dst + (size_t)m * row + EVDR_TC_DIM + lane * 4) = lo;
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
pack_queries_kernel(const T* __restrict__ Q, const uint8_t* __restrict__ qmask, int nq, int lq, int kb,
                    __nv_bfloat16* __restrict__ Qb, uint8_t* __restrict__ qmask32) {
  const int lane = threadIdx.x & 31;
  const long long slot = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (slot >= (long long)nq * EVDR_TC_LQ) return;
  const int q = (int)(slot / EVDR_TC_LQ), n = (int)(slot % EVDR_TC_LQ);
  const bool valid = (n < lq) && qmask[(size_t)q * lq + n] != 0;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (valid) v = load4(Q + ((size_t)q * lq + n) * EVDR_TC_DIM + lane * 4);
  const int row = EVDR_TC_DIM * (kb == 3 ? 2 : 1);   // kb = 3 (bf16x3 split): [hi | lo] rows of 256
  const uint2 hi = pack_bf16x4(v);
  *reinterpret_cast<uint2*>(Qb + (size_t)slot * row + lane * 4) = hi;
  if (kb == 3) {
    const uint2 lo = pack_bf16x4(residual_bf16x4(v, hi));
    *reinterpret_cast<uint2*>(Qb + (size_t)slot * row + EVDR_TC_DIM + lane * 4) = lo;
  }
  if (lane == 0) qmask32[slot] = valid ? 1 : 0;
}

// out[q, c] = sum_g in[(q*G + g), c]: queries longer than 32 tokens are scored as G 32-token groups
// (the max over page tokens is per query token, so the split is exact); fixed order -> deterministic
__global__ void __launch_bounds__(256)
group_sum_kernel(const float* __restrict__ in, long long in_ld, int nq, int G, int np,
                 float* __restrict__ out, long long out_ld) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)nq * np) return;
  const int q = (int)(i / np), c = (int)(i % np);
  float s = 0.f;
  for (int g = 0; g < G; ++g) s += in[((long long)q * G + g) * in_ld + c];
  out[(long long)q * out_ld + c] = s;
}

// y = (x*mask)/(||x*mask|| + eps); one warp per token, any d % 4 == 0
__global__ void __launch_bounds__(256)
l2norm_fwd_kernel(const float* __restrict__ X, const uint8_t* __restrict__ mask, long long n_tok,
                  int d, float* __restrict__ Y) {
  const int lane = threadIdx.x & 31;
  const long long t = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= n_tok) return;
  const bool valid = mask == nullptr || mask[t] != 0;
  const float* x = X + t * d;
  float* y = Y + t * d;
  float ss = 0.f;
  if (valid)
    for (int i = lane * 4; i < d; i += 128) {
      float4 v = *reinterpret_cast<const float4*>(x + i);
      ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
  ss = warp_sum(ss);
  const float inv = 1.f / (sqrtf(ss) + 1e-12f);
  for (int i = lane * 4; i < d; i += 128) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) {
      v = *reinterpret_cast<const float4*>(x + i);
      v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
    }
    *reinterpret_cast<float4*>(y + i) = v;
  }
}

// gx = mask * ( g/(n+eps) - x * (x.g) / (n (n+eps)^2) ),  n = ||x*mask||
__global__ void __launch_bounds__(256)
l2norm_bwd_kernel(const float* __restrict__ X, const uint8_t* __restrict__ mask,
                  const float* __restrict__ gY, long long n_tok, int d, float* __restrict__ gX) {
  const int lane = threadIdx.x & 31;
  const long long t = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= n_tok) return;
  const bool valid = mask == nullptr || mask[t] != 0;
  const float* x = X + t * d;
  const float* g = gY + t * d;
  float* o = gX + t * d;
  float ss = 0.f, xg = 0.f;
  if (valid)
    for (int i = lane * 4; i < d; i += 128) {
      float4 v = *reinterpret_cast<const float4*>(x + i);
      float4 w = *reinterpret_cast<const float4*>(g + i);
      ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
      xg += v.x * w.x + v.y * w.y + v.z * w.z + v.w * w.w;
    }
  ss = warp_sum(ss);
  xg = warp_sum(xg);
  const float n = sqrtf(ss);
  const float a = 1.f / (n + 1e-12f);
  // d/dx of x/(n+eps): g*a - x * (x.g) * a^2 / n ; for n == 0 the norm's subgradient is 0
  const float b = (n > 0.f) ? xg * a * a / n : 0.f;
  for (int i = lane * 4; i < d; i += 128) {
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) {
      float4 v = *reinterpret_cast<const float4*>(x + i);
      float4 w = *reinterpret_cast<const float4*>(g + i);
      r.x = w.x * a - v.x * b; r.y = w.y * a - v.y * b;
      r.z = w.z * a - v.z * b; r.w = w.w * a - v.w * b;
    }
    *reinterpret_cast<float4*>(o + i) = r;
  }
}

// d == 128 (ColPali): one float4 per lane is a whole token, so a warp keeps TPW tokens in registers (TPW
// independent 512-byte loads in flight, no second read of x), reduces their sums interleaved and streams the
// results out; grid-stride over token groups.
constexpr int L2N_TPW = 4;

__global__ void __launch_bounds__(256)
l2norm_fwd128_kernel(const float* __restrict__ X, const uint8_t* __restrict__ mask, long long n_tok,
                     float* __restrict__ Y) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarp = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long t0 = warp0 * L2N_TPW; t0 < n_tok; t0 += nwarp * L2N_TPW) {
    float4 v[L2N_TPW];
    float ss[L2N_TPW];
#pragma unroll
    for (int u = 0; u < L2N_TPW; ++u) {
      const long long t = t0 + u;
      const bool valid = t < n_tok && (mask == nullptr || mask[t] != 0);
      v[u] = valid ? __ldcs(reinterpret_cast<const float4*>(X + t * 128) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
      ss[u] = v[u].x * v[u].x + v[u].y * v[u].y + v[u].z * v[u].z + v[u].w * v[u].w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int u = 0; u < L2N_TPW; ++u) ss[u] += __shfl_xor_sync(0xffffffffu, ss[u], o);
#pragma unroll
    for (int u = 0; u < L2N_TPW; ++u) {
      const long long t = t0 + u;
      if (t >= n_tok) break;
      const float inv = 1.f / (sqrtf(ss[u]) + 1e-12f);
      __stcs(reinterpret_cast<float4*>(Y + t * 128) + lane,
             make_float4(v[u].x * inv, v[u].y * inv, v[u].z * inv, v[u].w * inv));
    }
  }
}

__global__ void __launch_bounds__(256)
l2norm_bwd128_kernel(const float* __restrict__ X, const uint8_t* __restrict__ mask,
                     const float* __restrict__ gY, long long n_tok, float* __restrict__ gX) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarp = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long t0 = warp0 * L2N_TPW; t0 < n_tok; t0 += nwarp * L2N_TPW) {
    float4 v[L2N_TPW], w[L2N_TPW];
    float ss[L2N_TPW], xg[L2N_TPW];
    bool ok[L2N_TPW];
#pragma unroll
    for (int u = 0; u < L2N_TPW; ++u) {
      const long long t = t0 + u;
      ok[u] = t < n_tok && (mask == nullptr || mask[t] != 0);
      v[u] = ok[u] ? __ldcs(reinterpret_cast<const float4*>(X + t * 128) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
      w[u] = ok[u] ? __ldcs(reinterpret_cast<const float4*>(gY + t * 128) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
      ss[u] = v[u].x * v[u].x + v[u].y * v[u].y + v[u].z * v[u].z + v[u].w * v[u].w;
      xg[u] = v[u].x * w[u].x + v[u].y * w[u].y + v[u].z * w[u].z + v[u].w * w[u].w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int u = 0; u < L2N_TPW; ++u) {
        ss[u] += __shfl_xor_sync(0xffffffffu, ss[u], o);
        xg[u] += __shfl_xor_sync(0xffffffffu, xg[u], o);
      }
#pragma unroll
    for (int u = 0; u < L2N_TPW; ++u) {
      const long long t = t0 + u;
      if (t >= n_tok) break;
      const float n = sqrtf(ss[u]);
      const float a = 1.f / (n + 1e-12f);
      const float b = (n > 0.f) ? xg[u] * a * a / n : 0.f;
      float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ok[u])
        r = make_float4(w[u].x * a - v[u].x * b, w[u].y * a - v[u].y * b, w[u].z * a - v[u].z * b,
                        w[u].w * a - v[u].w * b);
      __stcs(reinterpret_cast<float4*>(gX + t * 128) + lane, r);
    }
  }
}

}  // namespace evdr

using namespace evdr;

extern "C" {

static int pack_plan_impl(evdr_handle* h, const uint8_t* pmask, const long long* in_off, int np, int lp,
                          int varlen, int32_t* chunk_off, void* stream) {
  EVDR_REQUIRE_HANDLE(h);
  EVDR_CHECK_ARG(chunk_off && np >= 1 && lp >= 1, "evdr_pack_plan: bad arguments");
  EVDR_CHECK_ARG(!varlen || pmask, "evdr_pack_plan: varlen needs pmask");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int32_t* nch = nullptr;
  EVDR_CUDA(evdr::ws_alloc(h, &nch, sizeof(int32_t) * (size_t)np, st));
  plan_len_kernel<<<(np + 7) / 8, 256, 0, st>>>(pmask, in_off, np, lp, varlen, nch);
  EVDR_LAUNCH_CHECK("plan_len_kernel");
  plan_scan_kernel<<<1, 1024, 0, st>>>(nch, np, chunk_off);
  EVDR_LAUNCH_CHECK("plan_scan_kernel");
  EVDR_CUDA(cudaFreeAsync(nch, st));
  return EVDR_OK;
}

int evdr_pack_plan(evdr_handle* h, const uint8_t* pmask, int np, int lp, int varlen,
                   int32_t* chunk_off, void* stream) {
  return pack_plan_impl(h, pmask, nullptr, np, lp, varlen, chunk_off, stream);
}

int evdr_pack_plan_ragged(evdr_handle* h, const uint8_t* mask_flat, const int64_t* tok_off, int np,
                          int32_t* chunk_off, void* stream) {
  if (!mask_flat || !tok_off) return set_error(EVDR_ERR_INVALID, "evdr_pack_plan_ragged: null pointer");
  return pack_plan_impl(h, mask_flat, reinterpret_cast<const long long*>(tok_off), np, 1, 1, chunk_off, stream);
}

static int pack_pages_impl(evdr_handle* h, const void* P, int p_dtype, const uint8_t* pmask,
                           const long long* in_off, const int32_t* chunk_off, int np, int lp, int d,
                           int normalize, int k_blocks, void* tokens, void* stream) {
  EVDR_REQUIRE_HANDLE(h);
  EVDR_CHECK_ARG(P && pmask && chunk_off && tokens, "evdr_pack_pages: null pointer");
  EVDR_CHECK_ARG(d == EVDR_TC_DIM, "evdr_pack_pages: d=%d, the tensor-core path needs d=128", d);
  EVDR_CHECK_ARG(np >= 1 && lp >= 1, "evdr_pack_pages: bad shape");
  EVDR_CHECK_ARG(k_blocks == 1 || k_blocks == 3, "evdr_pack_pages: k_blocks=%d (1 or 3)", k_blocks);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // upper bound of slots per page: rect plan gives ceil(lp/16)*16, varlen never more
  size_t smem = sizeof(int) * (size_t)((lp + CH - 1) / CH) * CH;
  if (smem > h->smem_optin)
    return set_error(EVDR_ERR_UNSUPPORTED, "evdr_pack_pages: lp=%d too long", lp);
  if (p_dtype == EVDR_F32) {
    EVDR_CUDA(cudaFuncSetAttribute(pack_pages_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pack_pages_kernel<float><<<np, 256, smem, st>>>(static_cast<const float*>(P), pmask, in_off, chunk_off, np,
                                                   lp, normalize, k_blocks, static_cast<__nv_bfloat16*>(tokens));
  } else if (p_dtype == EVDR_BF16) {
    EVDR_CUDA(cudaFuncSetAttribute(pack_pages_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pack_pages_kernel<__nv_bfloat16><<<np, 256, smem, st>>>(static_cast<const __nv_bfloat16*>(P), pmask,
                                                           in_off, chunk_off, np, lp, normalize, k_blocks,
                                                           static_cast<__nv_bfloat16*>(tokens));
  } else {
    return set_error(EVDR_ERR_INVALID, "evdr_pack_pages: bad dtype %d", p_dtype);
  }
  EVDR_LAUNCH_CHECK("pack_pages_kernel");
  return EVDR_OK;
}

int evdr_pack_pages(evdr_handle* h, const void* P, int p_dtype, const uint8_t* pmask,
                    const int32_t* chunk_off, int np, int lp, int d, int normalize, int k_blocks,
                    void* tokens, void* stream) {
  return pack_pages_impl(h, P, p_dtype, pmask, nullptr, chunk_off, np, lp, d, normalize, k_blocks, tokens, stream);
}

int evdr_pack_pages_ragged(evdr_handle* h, const void* tok_flat, int dtype, const uint8_t* mask_flat,
                           const int64_t* tok_off, const int32_t* chunk_off, int np, int max_len, int d,
                           int normalize, int k_blocks, void* tokens, void* stream) {
  if (!tok_off) return set_error(EVDR_ERR_INVALID, "evdr_pack_pages_ragged: null tok_off");
  return pack_pages_impl(h, tok_flat, dtype, mask_flat, reinterpret_cast<const long long*>(tok_off), chunk_off,
                         np, max_len, d, normalize, k_blocks, tokens, stream);
}

int evdr_pack_queries(evdr_handle* h, const void* Q, int q_dtype, const uint8_t* qmask, int nq,
                      int lq, int d, int k_blocks, void* Qb, uint8_t* qmask32, void* stream) {
  EVDR_REQUIRE_HANDLE(h);
  EVDR_CHECK_ARG(Q && qmask && Qb && qmask32, "evdr_pack_queries: null pointer");
  EVDR_CHECK_ARG(d == EVDR_TC_DIM, "evdr_pack_queries: d=%d, the tensor-core path needs d=128", d);
  if (lq > EVDR_TC_LQ)
    return set_error(EVDR_ERR_UNSUPPORTED, "evdr_pack_queries: lq=%d > 32 (use evdr_maxsim_fwd)", lq);
  EVDR_CHECK_ARG(nq >= 1 && lq >= 1, "evdr_pack_queries: bad shape");
  EVDR_CHECK_ARG(k_blocks == 1 || k_blocks == 3, "evdr_pack_queries: k_blocks=%d (1 or 3)", k_blocks);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  long long slots = (long long)nq * EVDR_TC_LQ;
  unsigned grid = (unsigned)((slots + 7) / 8);
  if (q_dtype == EVDR_F32)
    pack_queries_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(Q), qmask, nq, lq, k_blocks,
                                                    static_cast<__nv_bfloat16*>(Qb), qmask32);
  else if (q_dtype == EVDR_BF16)
    pack_queries_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(Q), qmask,
                                                            nq, lq, k_blocks, static_cast<__nv_bfloat16*>(Qb),
                                                            qmask32);
  else
    return set_error(EVDR_ERR_INVALID, "evdr_pack_queries: bad dtype %d", q_dtype);
  EVDR_LAUNCH_CHECK("pack_queries_kernel");
  return EVDR_OK;
}

int evdr_group_sum(evdr_handle* h, const float* in, int64_t in_ld, int nq, int groups, int np,
                   float* out, int64_t out_ld, void* stream) {
  EVDR_REQUIRE_HANDLE(h);
  EVDR_CHECK_ARG(in && out && nq >= 1 && groups >= 1 && np >= 1 && in_ld >= np && out_ld >= np,
                 "evdr_group_sum: bad arguments");
  const long long n = (long long)nq * np;
  group_sum_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      in, in_ld, nq, groups, np, out, out_ld);
  EVDR_LAUNCH_CHECK("group_sum_kernel");
  return EVDR_OK;
}

int evdr_l2norm_fwd(evdr_handle* h, const float* X, const uint8_t* mask, int64_t n_tok, int d,
                    float* Y, void* stream) {
  EVDR_REQUIRE_HANDLE(h);
  EVDR_CHECK_ARG(X && Y && n_tok >= 1 && d >= 4 && d % 4 == 0, "evdr_l2norm_fwd: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (d == 128) {
    const long long groups = (n_tok + 8 * L2N_TPW - 1) / (8 * L2N_TPW);
    const long long cap = (long long)h->sm_count * 16;
    l2norm_fwd128_kernel<<<(unsigned)(groups < cap ? groups : cap), 256, 0, st>>>(X, mask, n_tok, Y);
  } else {
    l2norm_fwd_kernel<<<(unsigned)((n_tok + 7) / 8), 256, 0, st>>>(X, mask, n_tok, d, Y);
  }
  EVDR_LAUNCH_CHECK("l2norm_fwd_kernel");
  return EVDR_OK;
}

int evdr_l2norm_bwd(evdr_handle* h, const float* X, const uint8_t* mask, const float* gY,
                    int64_t n_tok, int d, float* gX, void* stream) {
  EVDR_REQUIRE_HANDLE(h);
  EVDR_CHECK_ARG(X && gY && gX && n_tok >= 1 && d >= 4 && d % 4 == 0, "evdr_l2norm_bwd: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (d == 128) {
    const long long groups = (n_tok + 8 * L2N_TPW - 1) / (8 * L2N_TPW);
    const long long cap = (long long)h->sm_count * 16;
    l2norm_bwd128_kernel<<<(unsigned)(groups < cap ? groups : cap), 256, 0, st>>>(X, mask, gY, n_tok, gX);
  } else {
    l2norm_bwd_kernel<<<(unsigned)((n_tok + 7) / 8), 256, 0, st>>>(X, mask, gY, n_tok, d, gX);
  }
  EVDR_LAUNCH_CHECK("l2norm_bwd_kernel");
  return EVDR_OK;
}

}  // extern "C"
