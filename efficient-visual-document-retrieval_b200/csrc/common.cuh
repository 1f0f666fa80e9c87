// common.cuh -- shared host/device helpers for libevdr.so (sm_100a only).
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cstdlib>

#include "../../include/evdr.h"

#define EVDR_VERSION_INT 100  // 0.1.0
#define EVDR_SCHED_RING 4096

// ---------------------------------------------------------------------------
// error plumbing: thread-local message + status code, no exceptions
// ---------------------------------------------------------------------------
namespace evdr {

char* tls_error_buf();
int set_error(int code, const char* fmt, ...);
// name + template arguments of the scoring kernel this thread launched last (evdr_last_kernel): lets tests and
// benchmarks assert WHICH flavour the dispatch picked
char* tls_kernel_buf();
void note_kernel(const char* fmt, ...);

#define EVDR_CHECK_ARG(cond, ...)                                        \
  do {                                                                   \
    if (!(cond)) return ::evdr::set_error(EVDR_ERR_INVALID, __VA_ARGS__); \
  } while (0)

#define EVDR_CUDA(call)                                                              \
  do {                                                                               \
    cudaError_t e__ = (call);                                                        \
    if (e__ != cudaSuccess)                                                          \
      return ::evdr::set_error(EVDR_ERR_CUDA, "%s failed: %s (%s:%d)", #call,        \
                               cudaGetErrorString(e__), __FILE__, __LINE__);         \
  } while (0)

#define EVDR_LAUNCH_CHECK(name)                                                      \
  do {                                                                               \
    cudaError_t e__ = cudaGetLastError();                                            \
    if (e__ != cudaSuccess)                                                          \
      return ::evdr::set_error(EVDR_ERR_CUDA, "launch of %s failed: %s", name,       \
                               cudaGetErrorString(e__));                             \
  } while (0)

// NCCL function table, resolved with dlsym at evdr_comm_init (sharded.cu)
struct NcclApi;

}  // namespace evdr

typedef CUresult (*evdr_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                         const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                         const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                         CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct evdr_handle {
  int device;
  int sm_count;
  int cc_major, cc_minor;
  size_t smem_optin;
  evdr_encode_tiled_fn encode_tiled;
  // scheduler counters of the tensor-core kernel: a ring of device ints owned by the handle (one per
  // launch, zeroed on the launch stream), so the hot path never allocates
  int* sched_ring;
  std::atomic<unsigned> sched_next;
  // stream-ordered scratch comes from a pool OWNED by the handle (the device's default pool and its
  // release threshold, which other libraries in the process may rely on, are left alone)
  cudaMemPool_t pool;
  // cta_group::2 flavours need clusters of two CTAs with ~224 KB of shared memory each to be schedulable
  // (not under every MPS / partitioning setup): 0 = not probed yet, 1 = yes, -1 = no (single-CTA flavours only)
  int pair_state;
  // multi-GPU
  void* nccl_comm;  // ncclComm_t
  int rank, world;
  // peer-memory candidate exchange over NVLink (sharded.cu): this rank's exchange region, every rank's
  // region as seen from here (device array of `world` pointers; entry `rank` is the local one), the
  // capacity of one slot in candidates and the call sequence number (epoch = seq & 1)
  unsigned char* peer_local;
  unsigned char** peer_ptrs_dev;
  void* peer_opened[16];
  int peer_world, peer_rank, peer_cap;
  unsigned* peer_err_host;              // pinned + mapped: non-zero after a failed wait (sticky)
  unsigned* peer_err_dev;               // the same word as seen from the device
  unsigned long long peer_timeout_ns;   // bound of one wait in the exchange kernel (0 = unbounded)
};

namespace evdr {
// internal cross-file entry points (not part of the C ABI)
// maxsim_tc.cu: scan with the top-k fused into the epilogue; EVDR_ERR_UNSUPPORTED (no message) when the
// launch would have more than one query group or needs the short-page epilogue.  cand_* hold
// nq * n_cta * k entries on return (n_cta <= sm_count).
int scan_topk(evdr_handle* h, const void* Qb, const uint8_t* qmask32, int nq, const void* tokens, int64_t n_tokens,
              const int32_t* chunk_off, int np, int64_t total_chunks, int k, int64_t id_offset, float* cand_s,
              long long* cand_i, int* n_cta, cudaStream_t st);
// maxsim_scan.cu: the operand-swapped streaming scan for 1-2 queries (scores and / or per-CTA top-k lists)
int scan_swapped(evdr_handle* h, const void* Qb, const uint8_t* qmask32, int nq, const void* tokens, int64_t n_tokens,
                 const int32_t* chunk_off, int np, int64_t total_chunks, int pages_per_item, float* out,
                 int64_t out_ld, int tk_k, int64_t id_offset, float* cand_s, long long* cand_i, int* n_cta,
                 cudaStream_t st);
// topk.cu: k <= 32 best of (nq, ncand) candidate rows, one CTA per query, one launch
int select_candidates(evdr_handle* h, const float* cand_s, const long long* cand_i, int nq, int ncand, int k,
                      float* top_s, long long* top_i, cudaStream_t st);
// sharded.cu
int peer_exchange(evdr_handle* h, const float* cand_s, const long long* cand_i, int nq, int ncand, int k,
                  float* top_scores, long long* top_ids, cudaStream_t st);
}  // namespace evdr

namespace evdr {
// stream-ordered workspace from the handle's private pool (freed with cudaFreeAsync)
template <typename T>
inline cudaError_t ws_alloc(evdr_handle* h, T** p, size_t bytes, cudaStream_t st) {
  return cudaMallocFromPoolAsync(reinterpret_cast<void**>(p), bytes, h->pool, st);
}
}  // namespace evdr

#define EVDR_REQUIRE_HANDLE(h)                                                        \
  do {                                                                                \
    if ((h) == nullptr) return ::evdr::set_error(EVDR_ERR_INVALID, "null handle");    \
    cudaError_t e__ = cudaSetDevice((h)->device);                                     \
    if (e__ != cudaSuccess)                                                           \
      return ::evdr::set_error(EVDR_ERR_CUDA, "cudaSetDevice(%d): %s", (h)->device,   \
                               cudaGetErrorString(e__));                              \
  } while (0)

// ---------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------
namespace evdr {

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

// load 4 consecutive elements as float4 (pointer must be aligned to 4 elements)
template <typename T>
__device__ __forceinline__ float4 load4(const T* p);
template <>
__device__ __forceinline__ float4 load4<float>(const float* p) {
  return *reinterpret_cast<const float4*>(p);
}
template <>
__device__ __forceinline__ float4 load4<__nv_bfloat16>(const __nv_bfloat16* p) {
  uint2 raw = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&raw.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&raw.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace evdr
