// sharded.cu -- multi-GPU tail: local top-k -> ncclAllGather of (score, id) candidates over
// NVLink/NVSwitch -> merge on every rank.  Pages are sharded by contiguous range, queries
// replicated.  NCCL is resolved at run time with dlopen/dlsym so that the
// library carries no link-time NCCL dependency and shares torch's libnccl.so.2.
#include <dlfcn.h>

#include <mutex>

#include "common.cuh"
#include "topk_dev.cuh"

namespace evdr {

typedef struct { char internal[EVDR_NCCL_UNIQUE_ID_BYTES]; } nccl_uid_t;
typedef void* nccl_comm_t;
typedef int nccl_result_t;  // ncclSuccess == 0
enum { kNcclChar = 0 };     // ncclInt8 / ncclChar

struct NcclApi {
  nccl_result_t (*GetUniqueId)(nccl_uid_t*);
  nccl_result_t (*CommInitRank)(nccl_comm_t*, int, nccl_uid_t, int);
  nccl_result_t (*CommDestroy)(nccl_comm_t);
  nccl_result_t (*AllGather)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t);
  const char* (*GetErrorString)(nccl_result_t);
  bool ok;
};

static NcclApi* nccl_api() {
  static NcclApi api = {};
  static std::once_flag once;                    // handles of several host threads may race to resolve NCCL
  std::call_once(once, [] {
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return;
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(lib, "ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(lib, "ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(lib, "ncclCommDestroy"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(lib, "ncclAllGather"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(lib, "ncclGetErrorString"));
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather && api.GetErrorString;
  });
  return api.ok ? &api : nullptr;
}

#define EVDR_NCCL(api, call)                                                             \
  do {                                                                                   \
    nccl_result_t r__ = (call);                                                          \
    if (r__ != 0)                                                                        \
      return set_error(EVDR_ERR_NCCL, "%s failed: %s", #call, (api)->GetErrorString(r__)); \
  } while (0)

// candidates are exchanged as one byte blob per rank: [nq*k int64 ids][nq*k floats, padded to 8 B]
__host__ __device__ inline size_t blob_bytes(size_t per) {
  return per * sizeof(long long) + ((per * sizeof(float) + 7) / 8) * 8;
}

// merge after the all-gather, straight from the gathered blobs: one CTA per query loads the world * k
// candidates of its query into shared memory and selects the k best (k <= 32: register lists)
__global__ void __launch_bounds__(SEL_WARPS * 32)
merge_gathered_kernel(const unsigned char* __restrict__ blob, int world, int nq, int k,
                      float* __restrict__ top_s, long long* __restrict__ top_i) {
  extern __shared__ unsigned char sm_raw[];
  long long* cid = reinterpret_cast<long long*>(sm_raw);                  // [world * k]
  long long* wi = cid + (size_t)world * k;                                // [SEL_WARPS * 32]
  long long* sel_i = wi + SEL_WARPS * 32;                                 // [k]
  float* csc = reinterpret_cast<float*>(sel_i + k);                       // [world * k]
  float* ws = csc + (size_t)world * k;                                    // [SEL_WARPS * 32]
  float* sel_s = ws + SEL_WARPS * 32;                                     // [k]
  const int q = blockIdx.x;
  const int per = nq * k, n = world * k;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int r = i / k, t = i % k;
    const unsigned char* base = blob + (size_t)r * blob_bytes(per);
    cid[i] = reinterpret_cast<const long long*>(base)[q * k + t];
    csc[i] = reinterpret_cast<const float*>(base + (size_t)per * sizeof(long long))[q * k + t];
  }
  __syncthreads();
  block_select_topk(csc, cid, n, k, ws, wi, sel_s, sel_i);
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    top_s[(size_t)q * k + j] = sel_s[j];
    top_i[(size_t)q * k + j] = sel_i[j];
  }
}

// unpack for k > 32 (the general two-stage selection of topk.cu takes contiguous candidate rows)
__global__ void unpack_gathered_kernel(const unsigned char* __restrict__ blob, int world, int nq,
                                       int k, float* __restrict__ cs, long long* __restrict__ ci) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int per = nq * k;
  if (i >= world * per) return;
  const int r = i / per, j = i % per, q = j / k, t = j % k;
  const unsigned char* base = blob + (size_t)r * blob_bytes(per);
  const long long* id = reinterpret_cast<const long long*>(base);
  const float* s = reinterpret_cast<const float*>(base + (size_t)per * sizeof(long long));
  const size_t o = (size_t)q * world * k + (size_t)r * k + t;
  cs[o] = s[j];
  ci[o] = id[j];
}

// score column blocks for the training collective: pages are partitioned contiguously (first
// np_total % world ranks own one page more), NCCL wants equal counts, so blocks travel padded to cmax
// columns and are laid out by page offset on arrival
__host__ __device__ inline int part_count(int n, int world, int r) { return n / world + (r < n % world ? 1 : 0); }
__host__ __device__ inline int part_start(int n, int world, int r) {
  return r * (n / world) + (r < n % world ? r : n % world);
}

__global__ void pad_block_kernel(const float* __restrict__ local, long long ld, int B, int cnt, int cmax,
                                 float* __restrict__ send) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * cmax) return;
  const int b = (int)(i / cmax), j = (int)(i % cmax);
  send[i] = j < cnt ? local[(size_t)b * ld + j] : 0.f;
}

__global__ void place_blocks_kernel(const float* __restrict__ recv, int world, int B, int np_total, int cmax,
                                    float* __restrict__ out, long long out_ld) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)world * B * cmax) return;
  const int j = (int)(i % cmax);
  const int b = (int)((i / cmax) % B);
  const int r = (int)(i / ((long long)cmax * B));
  if (j < part_count(np_total, world, r)) out[(size_t)b * out_ld + part_start(np_total, world, r) + j] = recv[i];
}

// ---------------------------------------------------------------------------------------------
// Peer-memory exchange: local selection + all-gather + merge in ONE kernel over NVLink / NVSwitch P2P stores.
//
// Every rank owns an exchange region (cudaMalloc, shared through CUDA IPC):
//     data  [2 epochs][world source ranks][cap candidates] x (int64 id, fp32 score)
//     flags [2][world][PEER_MAXQ] u32      meta [2][world] u32 ((nq << 8) | k of the source's call)
//     ctl   {seq, done, err}               (local use only)
// Per query q (CTA-strided): the CTA reduces this rank's candidates of q to the local top-k (they may be
// the k * n_cta per-CTA lists of the fused scan, or already k), copies them into the slot (epoch, my rank)
// of EVERY rank's region with plain remote stores, fences at system scope and releases the per-(source
// rank, query) flag on every rank; it then acquires the `world` flags of q in its own region and merges
// the world * k candidates (score desc, id asc) straight into the result.  No NCCL launch, no unpack
// pass; the transfer of query q overlaps the merge of the queries before it.
//
// The call sequence number lives ON THE DEVICE (ctl.seq, advanced by the last CTA of a launch), so the
// kernel can be captured in a CUDA graph and replayed, and a host-side early return can never desynchronise
// the ranks.  Two epochs (sequence parity) suffice: a rank can only start call n+2 after every peer has
// pushed call n+1, i.e. after that peer's call-n kernel -- the last reader of the epoch being overwritten
// -- is complete (stream order).
//
// Contract (collective): every rank calls with the same nq and k (checked through `meta`; a mismatch is
// reported like a timeout), on ONE stream per handle.  Progress: a CTA pushes before it waits and grids
// are identical on all ranks and capped at the co-resident CTA count, so the lowest unfinished query of
// every rank always has its CTA running.  Waits are bounded by a wall-clock timeout (EVDR_PEER_TIMEOUT_MS,
// default 30000, 0 = wait forever): on expiry the kernel does NOT trap -- it records the error (device
// word + pinned host word), writes (-inf, -1) results for the affected queries, skips all further
// waiting and exits normally; the next evdr_peer_* call on the handle returns EVDR_ERR_STATE.
// ---------------------------------------------------------------------------------------------
constexpr int PEER_MAXQ = 4096;           // queries per call
constexpr int PEER_MAX_WORLD = 16;
constexpr int PEER_THREADS = SEL_WARPS * 32;

struct PeerCtl { unsigned seq, done, err, pad; };

__host__ __device__ inline size_t peer_slot_bytes(size_t cap) { return cap * 12; }   // ids then scores
__host__ __device__ inline size_t peer_data_bytes(int world, size_t cap) {
  return ((2 * (size_t)world * peer_slot_bytes(cap) + 127) / 128) * 128;
}
__host__ __device__ inline size_t peer_flag_bytes(int world) { return sizeof(unsigned) * 2 * (size_t)world * PEER_MAXQ; }
__host__ __device__ inline size_t peer_meta_off(int world, size_t cap) { return peer_data_bytes(world, cap) + peer_flag_bytes(world); }
__host__ __device__ inline size_t peer_ctl_off(int world, size_t cap) {
  return peer_meta_off(world, cap) + ((sizeof(unsigned) * 2 * (size_t)world + 15) / 16) * 16;
}
__host__ __device__ inline size_t peer_region_bytes(int world, size_t cap) { return peer_ctl_off(world, cap) + sizeof(PeerCtl); }

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

enum { PEER_ERR_TIMEOUT = 1, PEER_ERR_SHAPE = 2 };

__global__ void __launch_bounds__(PEER_THREADS)
exchange_merge_kernel(unsigned char* const* __restrict__ regions, int world, int rank, int cap,
                      const float* __restrict__ cand_s, const long long* __restrict__ cand_i, int ncand,
                      int nq, int k, float* __restrict__ top_s, long long* __restrict__ top_i,
                      unsigned long long timeout_ns, unsigned* __restrict__ err_host) {
  extern __shared__ unsigned char sm_raw[];
  long long* cid = reinterpret_cast<long long*>(sm_raw);                  // [world * k] gathered candidates
  long long* wi = cid + (size_t)world * k;                                // [SEL_WARPS * 32] selection scratch
  long long* loc_i = wi + SEL_WARPS * 32;                                 // [k] this rank's top-k of q
  float* csc = reinterpret_cast<float*>(loc_i + k);                       // [world * k]
  float* ws = csc + (size_t)world * k;                                    // [SEL_WARPS * 32]
  float* loc_s = ws + SEL_WARPS * 32;                                     // [k]
  const size_t slot = peer_slot_bytes(cap);
  const size_t data = peer_data_bytes(world, cap);
  PeerCtl* ctl = reinterpret_cast<PeerCtl*>(regions[rank] + peer_ctl_off(world, cap));
  const unsigned seq = *reinterpret_cast<volatile unsigned*>(&ctl->seq) + 1u;   // same for every CTA of the launch
  const int epoch = seq & 1u;
  const unsigned meta_mine = ((unsigned)nq << 8) | (unsigned)k;
  const int per = nq * k;
  const int n = world * k;
  for (int q = blockIdx.x; q < nq; q += gridDim.x) {
    // ---- this rank's k best of query q
    if (ncand > k) {
      block_select_topk(cand_s + (size_t)q * ncand, cand_i + (size_t)q * ncand, ncand, k, ws, wi, loc_s, loc_i);
    } else {
      for (int t = threadIdx.x; t < k; t += blockDim.x) {
        loc_s[t] = cand_s[(size_t)q * k + t];
        loc_i[t] = cand_i[(size_t)q * k + t];
      }
      __syncthreads();
    }
    // ---- push them into slot (epoch, rank) of every region
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const int r = i / k, t = i % k;
      unsigned char* base = regions[r] + ((size_t)epoch * world + rank) * slot;
      reinterpret_cast<long long*>(base)[q * k + t] = loc_i[t];
      reinterpret_cast<float*>(base + (size_t)per * 8)[q * k + t] = loc_s[t];
    }
    __threadfence_system();
    __syncthreads();
    bool ok = true;
    if (threadIdx.x < world) {
      unsigned char* peer = regions[threadIdx.x];
      reinterpret_cast<unsigned*>(peer + peer_meta_off(world, cap))[epoch * world + rank] = meta_mine;
      st_release_sys(reinterpret_cast<unsigned*>(peer + data) + ((size_t)epoch * world + rank) * PEER_MAXQ + q, seq);
      // ---- wait for source rank threadIdx.x to have delivered query q here
      const unsigned* mine = reinterpret_cast<const unsigned*>(regions[rank] + data) +
                             ((size_t)epoch * world + threadIdx.x) * PEER_MAXQ + q;
      volatile unsigned* derr = &ctl->err;
      unsigned long long t0 = 0;
      while (ld_acquire_sys(mine) != seq) {
        if (*derr != 0u) { ok = false; break; }            // a wait already failed in this launch
        if (t0 == 0) t0 = global_timer_ns();
        __nanosleep(200);
        if (timeout_ns != 0 && global_timer_ns() - t0 > timeout_ns) {
          atomicCAS(&ctl->err, 0u, (unsigned)PEER_ERR_TIMEOUT | ((unsigned)threadIdx.x << 8));
          *reinterpret_cast<volatile unsigned*>(err_host) = (unsigned)PEER_ERR_TIMEOUT | ((unsigned)threadIdx.x << 8);
          ok = false;
          break;
        }
      }
      if (ok) {
        const unsigned m = __ldcv(reinterpret_cast<const unsigned*>(regions[rank] + peer_meta_off(world, cap)) +
                                  epoch * world + threadIdx.x);
        if (m != meta_mine) {                                // the peer called with another nq / k
          atomicCAS(&ctl->err, 0u, (unsigned)PEER_ERR_SHAPE | ((unsigned)threadIdx.x << 8));
          *reinterpret_cast<volatile unsigned*>(err_host) = (unsigned)PEER_ERR_SHAPE | ((unsigned)threadIdx.x << 8);
          ok = false;
        }
      }
    }
    if (!__syncthreads_and(ok ? 1 : 0)) {
      for (int t = threadIdx.x; t < k; t += blockDim.x) {
        top_s[(size_t)q * k + t] = -INFINITY;
        top_i[(size_t)q * k + t] = -1;
      }
      __syncthreads();
      continue;
    }
    // ---- gather the world * k candidates of q (remote writes: bypass L1)
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const int r = i / k, t = i % k;
      const unsigned char* base = regions[rank] + ((size_t)epoch * world + r) * slot;
      cid[i] = __ldcv(reinterpret_cast<const long long*>(base) + q * k + t);
      csc[i] = __ldcv(reinterpret_cast<const float*>(base + (size_t)per * 8) + q * k + t);
    }
    __syncthreads();
    // ---- merge: position = number of better candidates (score desc, id asc; padding (-inf, -1) last)
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const float s = csc[i];
      const long long id = cid[i];
      const bool pad = id < 0;
      int pos = 0;
      for (int j = 0; j < n; ++j) {
        const float os = csc[j];
        const long long oid = cid[j];
        const bool opad = oid < 0;
        const bool better = pad ? (!opad || j < i)
                                : (!opad && ((os > s) || (os == s && (oid < id || (oid == id && j < i)))));
        pos += better;
      }
      if (pos < k) {
        top_s[(size_t)q * k + pos] = pad ? -INFINITY : s;
        top_i[(size_t)q * k + pos] = pad ? -1 : id;
      }
    }
    __syncthreads();
  }
  // ---- the last CTA of the launch advances the call sequence number
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&ctl->done, 1u) == gridDim.x - 1) {
      ctl->done = 0;
      __threadfence();
      *reinterpret_cast<volatile unsigned*>(&ctl->seq) = seq;
    }
  }
}

static size_t exchange_smem(int world, int k) {
  return ((size_t)world * k + SEL_WARPS * 32 + k) * (sizeof(long long) + sizeof(float));
}

// candidates (nq, ncand) of this rank -> global top-k on every rank (collective); ncand == k or k <= 32
int peer_exchange(evdr_handle* h, const float* cand_s, const long long* cand_i, int nq, int ncand, int k,
                  float* top_scores, long long* top_ids, cudaStream_t st) {
  if (!h->peer_ptrs_dev) return set_error(EVDR_ERR_STATE, "peer exchange before evdr_peer_connect");
  EVDR_CHECK_ARG(nq >= 1 && nq <= PEER_MAXQ && k >= 1 && k <= 255 && (long long)nq * k <= h->peer_cap,
                 "peer exchange: nq=%d x k=%d exceeds the exchange region (%d candidates, %d queries)",
                 nq, k, h->peer_cap, PEER_MAXQ);
  EVDR_CHECK_ARG(ncand == k || (ncand > k && k <= 32), "peer exchange: %d candidates per query need k <= 32", ncand);
  if (h->peer_err_host && *reinterpret_cast<volatile unsigned*>(h->peer_err_host) != 0u) {
    const unsigned e = *reinterpret_cast<volatile unsigned*>(h->peer_err_host);
    return set_error(EVDR_ERR_STATE, "peer exchange is disabled on this handle: an earlier call %s (rank %u)",
                     (e & 0xffu) == PEER_ERR_TIMEOUT ? "timed out waiting for a peer" : "met a peer with another nq / k",
                     e >> 8);
  }
  const size_t smem = exchange_smem(h->peer_world, k);
  if (smem > 48 * 1024) {
    if (smem > h->smem_optin) return set_error(EVDR_ERR_UNSUPPORTED, "peer exchange: world * k too large for shared memory");
    EVDR_CUDA(cudaFuncSetAttribute(exchange_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  // identical on every rank (same binary, same device type) and never more than can be co-resident
  int per_sm = 0;
  EVDR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, exchange_merge_kernel, PEER_THREADS, smem));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 2) per_sm = 2;
  const int grid = nq < per_sm * h->sm_count ? nq : per_sm * h->sm_count;
  exchange_merge_kernel<<<grid, PEER_THREADS, smem, st>>>(
      h->peer_ptrs_dev, h->peer_world, h->peer_rank, h->peer_cap, cand_s, cand_i, ncand, nq, k, top_scores,
      top_ids, h->peer_timeout_ns, h->peer_err_dev);
  EVDR_LAUNCH_CHECK("exchange_merge_kernel");
  return EVDR_OK;
}

// candidates of this rank -> global top-k through ONE ncclAllGather; `fill(loc_s, loc_i)` writes this
// rank's nq * k (score, id) pairs straight into the send blob
template <typename Fill>
static int nccl_exchange(evdr_handle* h, int nq, int k, float* top_scores, long long* top_ids, cudaStream_t st,
                         Fill fill) {
  if (!h->nccl_comm) return set_error(EVDR_ERR_STATE, "sharded search before evdr_comm_init");
  NcclApi* api = nccl_api();
  if (!api) return set_error(EVDR_ERR_NCCL, "libnccl.so.2 not loadable");
  const int world = h->world;
  const size_t per = (size_t)nq * k;
  const size_t rank_bytes = blob_bytes(per);
  // stream-ordered scratch: [send blob][recv blobs x world][cand ids world*per][cand scores world*per]
  unsigned char* ws = nullptr;
  EVDR_CUDA(ws_alloc(h, &ws, rank_bytes * (2 * (size_t)world + 1) + 64, st));
  unsigned char* send = ws;
  unsigned char* recv = ws + rank_bytes;
  long long* loc_i = reinterpret_cast<long long*>(send);
  float* loc_s = reinterpret_cast<float*>(send + per * sizeof(long long));
  int rc = fill(loc_s, loc_i);
  if (rc != EVDR_OK) { cudaFreeAsync(ws, st); return rc; }
  nccl_result_t r = api->AllGather(send, recv, rank_bytes, kNcclChar, static_cast<nccl_comm_t>(h->nccl_comm), st);
  if (r != 0) {
    cudaFreeAsync(ws, st);
    return set_error(EVDR_ERR_NCCL, "ncclAllGather failed: %s", api->GetErrorString(r));
  }
  if (k <= 32) {
    const size_t smem = exchange_smem(world, k);
    merge_gathered_kernel<<<nq, SEL_WARPS * 32, smem, st>>>(recv, world, nq, k, top_scores, top_ids);
    cudaError_t le = cudaGetLastError();
    cudaFreeAsync(ws, st);
    if (le != cudaSuccess)
      return set_error(EVDR_ERR_CUDA, "launch of merge_gathered_kernel failed: %s", cudaGetErrorString(le));
    return EVDR_OK;
  }
  long long* cand_i = reinterpret_cast<long long*>(recv + rank_bytes * world);
  float* cand_s = reinterpret_cast<float*>(cand_i + (size_t)world * per);
  const int total = world * (int)per;
  unpack_gathered_kernel<<<(total + 255) / 256, 256, 0, st>>>(recv, world, nq, k, cand_s, cand_i);
  cudaError_t le = cudaGetLastError();
  if (le != cudaSuccess) {
    cudaFreeAsync(ws, st);
    return set_error(EVDR_ERR_CUDA, "launch of unpack_gathered_kernel failed: %s", cudaGetErrorString(le));
  }
  rc = evdr_topk_merge(h, cand_s, reinterpret_cast<const int64_t*>(cand_i), nq, world * k, k, top_scores,
                       reinterpret_cast<int64_t*>(top_ids), st);
  cudaFreeAsync(ws, st);
  return rc;
}

}  // namespace evdr

using namespace evdr;

extern "C" {

int evdr_peer_region_init(evdr_handle* h, int rank, int world, int max_candidates, void* ipc_handle_out) {
  EVDR_REQUIRE_HANDLE(h);
  EVDR_CHECK_ARG(ipc_handle_out && world >= 1 && world <= PEER_MAX_WORLD && rank >= 0 && rank < world &&
                 max_candidates >= 1, "evdr_peer_region_init: bad arguments");
  if (h->peer_local) return set_error(EVDR_ERR_STATE, "evdr_peer_region_init: already initialised");
  static_assert(sizeof(cudaIpcMemHandle_t) == EVDR_IPC_HANDLE_BYTES, "IPC handle size");
  max_candidates += max_candidates & 1;          // slots stay 8-byte aligned
  const size_t bytes = peer_region_bytes(world, (size_t)max_candidates);
  unsigned char* reg = nullptr;
  EVDR_CUDA(cudaMalloc(reinterpret_cast<void**>(&reg), bytes));
  EVDR_CUDA(cudaMemset(reg, 0, bytes));
  cudaIpcMemHandle_t ipc;
  cudaError_t e = cudaIpcGetMemHandle(&ipc, reg);
  if (e != cudaSuccess) {
    cudaFree(reg);
    return set_error(EVDR_ERR_CUDA, "cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
  }
  // error word: pinned + mapped, so the host sees a failed wait without synchronising
  e = cudaHostAlloc(reinterpret_cast<void**>(&h->peer_err_host), sizeof(unsigned), cudaHostAllocMapped);
  if (e == cudaSuccess) {
    *h->peer_err_host = 0u;
    e = cudaHostGetDevicePointer(reinterpret_cast<void**>(&h->peer_err_dev), h->peer_err_host, 0);
  }
  if (e != cudaSuccess) {
    cudaFree(reg);
    return set_error(EVDR_ERR_CUDA, "mapped error word for the peer exchange: %s", cudaGetErrorString(e));
  }
  unsigned long long ms = 30000ull;
  if (const char* t = getenv("EVDR_PEER_TIMEOUT_MS")) ms = strtoull(t, nullptr, 10);
  h->peer_timeout_ns = ms * 1000000ull;
  memcpy(ipc_handle_out, &ipc, sizeof(ipc));
  h->peer_local = reg;
  h->peer_rank = rank;
  h->peer_world = world;
  h->peer_cap = max_candidates;
  return EVDR_OK;
}

int evdr_peer_connect(evdr_handle* h, const void* all_ipc_handles) {
  EVDR_REQUIRE_HANDLE(h);
  EVDR_CHECK_ARG(all_ipc_handles, "evdr_peer_connect: null pointer");
  if (!h->peer_local) return set_error(EVDR_ERR_STATE, "evdr_peer_connect before evdr_peer_region_init");
  if (h->peer_ptrs_dev) return set_error(EVDR_ERR_STATE, "evdr_peer_connect: already connected");
  unsigned char* ptrs[PEER_MAX_WORLD] = {};
  for (int r = 0; r < h->peer_world; ++r) {
    if (r == h->peer_rank) { ptrs[r] = h->peer_local; continue; }
    cudaIpcMemHandle_t ipc;
    memcpy(&ipc, static_cast<const unsigned char*>(all_ipc_handles) + (size_t)r * sizeof(ipc), sizeof(ipc));
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, ipc, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess)
      return set_error(EVDR_ERR_CUDA, "cudaIpcOpenMemHandle(rank %d) failed: %s (GPUs without peer access?)",
                       r, cudaGetErrorString(e));
    h->peer_opened[r] = p;
    ptrs[r] = static_cast<unsigned char*>(p);
  }
  EVDR_CUDA(cudaMalloc(reinterpret_cast<void**>(&h->peer_ptrs_dev), sizeof(ptrs)));
  EVDR_CUDA(cudaMemcpy(h->peer_ptrs_dev, ptrs, sizeof(ptrs), cudaMemcpyHostToDevice));
  return EVDR_OK;
}

int evdr_peer_world(const evdr_handle* h) { return (h && h->peer_ptrs_dev) ? h->peer_world : 0; }

int evdr_peer_sharded_topk(evdr_handle* h, const float* local_scores, int64_t ld, int nq, int np_local,
                           int k, int64_t id_offset, float* top_scores, int64_t* top_ids, void* stream) {
  EVDR_REQUIRE_HANDLE(h);
  EVDR_CHECK_ARG(local_scores && top_scores && top_ids, "evdr_peer_sharded_topk: null pointer");
  if (!h->peer_ptrs_dev) return set_error(EVDR_ERR_STATE, "evdr_peer_sharded_topk before evdr_peer_connect");
  EVDR_CHECK_ARG(nq >= 1 && nq <= PEER_MAXQ && k >= 1 && (long long)nq * k <= h->peer_cap,
                 "evdr_peer_sharded_topk: nq=%d x k=%d exceeds the exchange region (%d candidates, %d queries)",
                 nq, k, h->peer_cap, PEER_MAXQ);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t per = (size_t)nq * k;
  unsigned char* ws = nullptr;
  EVDR_CUDA(ws_alloc(h, &ws, per * 12 + 64, st));
  int64_t* loc_i = reinterpret_cast<int64_t*>(ws);
  float* loc_s = reinterpret_cast<float*>(ws + per * 8);
  int rc = evdr_topk(h, local_scores, ld, nq, np_local, k, nullptr, id_offset, loc_s, loc_i, stream);
  if (rc == EVDR_OK)
    rc = peer_exchange(h, loc_s, reinterpret_cast<const long long*>(loc_i), nq, k, k, top_scores,
                       reinterpret_cast<long long*>(top_ids), st);
  cudaFreeAsync(ws, st);
  return rc;
}

int evdr_allgather_score_rows(evdr_handle* h, const float* local_scores, int64_t ld, int B, int np_total,
                              float* out, int64_t out_ld, void* stream) {
  EVDR_REQUIRE_HANDLE(h);
  EVDR_CHECK_ARG(local_scores && out && B >= 1 && np_total >= 1 && out_ld >= np_total,
                 "evdr_allgather_score_rows: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int world = h->nccl_comm ? h->world : 1, rank = h->nccl_comm ? h->rank : 0;
  const int cnt = part_count(np_total, world, rank);
  EVDR_CHECK_ARG(ld >= cnt, "evdr_allgather_score_rows: ld=%lld < this rank's %d pages", (long long)ld, cnt);
  if (world == 1) {
    EVDR_CUDA(cudaMemcpy2DAsync(out, out_ld * sizeof(float), local_scores, ld * sizeof(float),
                                (size_t)np_total * sizeof(float), B, cudaMemcpyDeviceToDevice, st));
    return EVDR_OK;
  }
  NcclApi* api = nccl_api();
  if (!api) return set_error(EVDR_ERR_NCCL, "libnccl.so.2 not loadable");
  const int cmax = part_count(np_total, world, 0);
  const size_t blk = (size_t)B * cmax;
  float* ws = nullptr;
  EVDR_CUDA(evdr::ws_alloc(h, reinterpret_cast<void**>(&ws), sizeof(float) * blk * ((size_t)world + 1), st));
  float* send = ws;
  float* recv = ws + blk;
  pad_block_kernel<<<(unsigned)((blk + 255) / 256), 256, 0, st>>>(local_scores, ld, B, cnt, cmax, send);
  nccl_result_t r = api->AllGather(send, recv, blk * sizeof(float), kNcclChar,
                                   static_cast<nccl_comm_t>(h->nccl_comm), st);
  if (r != 0) {
    cudaFreeAsync(ws, st);
    return set_error(EVDR_ERR_NCCL, "ncclAllGather failed: %s", api->GetErrorString(r));
  }
  place_blocks_kernel<<<(unsigned)((blk * world + 255) / 256), 256, 0, st>>>(recv, world, B, np_total, cmax, out,
                                                                             out_ld);
  cudaError_t le = cudaGetLastError();
  cudaFreeAsync(ws, st);
  if (le != cudaSuccess)
    return set_error(EVDR_ERR_CUDA, "launch of place_blocks_kernel failed: %s", cudaGetErrorString(le));
  return EVDR_OK;
}

int evdr_destroy(evdr_handle* h) {
  if (!h) return EVDR_OK;
  for (int r = 0; r < PEER_MAX_WORLD; ++r)
    if (h->peer_opened[r]) cudaIpcCloseMemHandle(h->peer_opened[r]);
  if (h->peer_ptrs_dev) cudaFree(h->peer_ptrs_dev);
  if (h->peer_local) cudaFree(h->peer_local);
  if (h->peer_err_host) cudaFreeHost(h->peer_err_host);
  if (h->pool) cudaMemPoolDestroy(h->pool);
  if (h->nccl_comm) {
    NcclApi* api = nccl_api();
    if (api) api->CommDestroy(static_cast<nccl_comm_t>(h->nccl_comm));
    h->nccl_comm = nullptr;
  }
  if (h->sched_ring) cudaFree(h->sched_ring);
  delete h;
  return EVDR_OK;
}

int evdr_comm_unique_id(void* id_out) {
  EVDR_CHECK_ARG(id_out, "evdr_comm_unique_id: null pointer");
  NcclApi* api = nccl_api();
  if (!api) return set_error(EVDR_ERR_NCCL, "libnccl.so.2 not loadable: %s", dlerror());
  nccl_uid_t uid;
  EVDR_NCCL(api, api->GetUniqueId(&uid));
  memcpy(id_out, &uid, sizeof(uid));
  return EVDR_OK;
}

int evdr_comm_init(evdr_handle* h, const void* id, int rank, int world) {
  EVDR_REQUIRE_HANDLE(h);
  EVDR_CHECK_ARG(id && world >= 1 && rank >= 0 && rank < world, "evdr_comm_init: bad arguments");
  if (h->nccl_comm) return set_error(EVDR_ERR_STATE, "evdr_comm_init: communicator already initialised");
  NcclApi* api = nccl_api();
  if (!api) return set_error(EVDR_ERR_NCCL, "libnccl.so.2 not loadable: %s", dlerror());
  nccl_uid_t uid;
  memcpy(&uid, id, sizeof(uid));
  nccl_comm_t comm = nullptr;
  EVDR_NCCL(api, api->CommInitRank(&comm, world, uid, rank));
  h->nccl_comm = comm;
  h->rank = rank;
  h->world = world;
  return EVDR_OK;
}

int evdr_comm_world(const evdr_handle* h) { return h ? h->world : 0; }

int evdr_sharded_topk(evdr_handle* h, const float* local_scores, int64_t ld, int nq, int np_local,
                      int k, int64_t id_offset, float* top_scores, int64_t* top_ids,
                      void* stream) {
  EVDR_REQUIRE_HANDLE(h);
  EVDR_CHECK_ARG(local_scores && top_scores && top_ids, "evdr_sharded_topk: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (h->world == 1 && !h->nccl_comm)
    return evdr_topk(h, local_scores, ld, nq, np_local, k, nullptr, id_offset, top_scores, top_ids, stream);
  return nccl_exchange(h, nq, k, top_scores, reinterpret_cast<long long*>(top_ids), st,
                       [&](float* loc_s, long long* loc_i) {
                         return evdr_topk(h, local_scores, ld, nq, np_local, k, nullptr, id_offset, loc_s,
                                          reinterpret_cast<int64_t*>(loc_i), stream);
                       });
}

/* Fused search: packed queries -> local scan with the top-k folded into the scan's epilogue (when the launch
 * has one query group; otherwise scores + evdr_topk on a stream-ordered scratch row) -> candidate exchange
 * (exchange: 0 = none / single GPU, 1 = ncclAllGather + merge, 2 = peer-memory kernel) -> global top-k. */
int evdr_sharded_search(evdr_handle* h, const void* Qb, const uint8_t* qmask32, int nq, const void* tokens,
                        int64_t n_tokens, const int32_t* chunk_off, int np, int64_t total_chunks, int k,
                        int64_t id_offset, int exchange, float* top_scores, int64_t* top_ids, void* stream) {
  EVDR_REQUIRE_HANDLE(h);
  EVDR_CHECK_ARG(top_scores && top_ids, "evdr_sharded_search: null output pointer");
  EVDR_CHECK_ARG(exchange >= 0 && exchange <= 2, "evdr_sharded_search: exchange=%d (0 none, 1 nccl, 2 peer)", exchange);
  EVDR_CHECK_ARG(k >= 1 && k <= 128, "evdr_sharded_search: k=%d outside [1,128]", k);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool solo = exchange == 0 || (exchange == 1 && h->world == 1 && !h->nccl_comm);
  // per-CTA candidate lists of the fused scan (one query group: at most 8 queries, k <= 32) ...
  const int max_cta = h->sm_count;
  unsigned char* ws = nullptr;
  const bool try_fused = nq <= 8 && k <= 32;
  const size_t cand_n = try_fused ? (size_t)nq * max_cta * k : 0;
  const size_t row_bytes = sizeof(float) * (size_t)nq * np;     // ... or the score rows of the unfused route
  EVDR_CUDA(ws_alloc(h, &ws, cand_n * 12 + 64, st));
  long long* cand_i = reinterpret_cast<long long*>(ws);
  float* cand_s = reinterpret_cast<float*>(ws + cand_n * 8);
  int n_cta = 0;
  int rc = try_fused ? scan_topk(h, Qb, qmask32, nq, tokens, n_tokens, chunk_off, np, total_chunks, k, id_offset,
                                 cand_s, cand_i, &n_cta, st)
                     : EVDR_ERR_UNSUPPORTED;
  if (rc == EVDR_OK) {
    const int ncand = n_cta * k;
    if (solo) rc = select_candidates(h, cand_s, cand_i, nq, ncand, k, top_scores, reinterpret_cast<long long*>(top_ids), st);
    else if (exchange == 2) rc = peer_exchange(h, cand_s, cand_i, nq, ncand, k, top_scores, reinterpret_cast<long long*>(top_ids), st);
    else rc = nccl_exchange(h, nq, k, top_scores, reinterpret_cast<long long*>(top_ids), st,
                            [&](float* loc_s, long long* loc_i) {
                              return select_candidates(h, cand_s, cand_i, nq, ncand, k, loc_s, loc_i, st);
                            });
    cudaFreeAsync(ws, st);
    return rc;
  }
  cudaFreeAsync(ws, st);
  if (rc != EVDR_ERR_UNSUPPORTED) return rc;
  // several query groups (more than 8 queries), short pages or k > 32: scores, then the two-stage top-k
  float* rows = nullptr;
  EVDR_CUDA(ws_alloc(h, &rows, row_bytes, st));
  rc = evdr_maxsim_packed_fwd(h, Qb, qmask32, nq, tokens, n_tokens, chunk_off, np, total_chunks, 0, 1, rows, np, stream);
  if (rc == EVDR_OK) {
    if (solo) rc = evdr_topk(h, rows, np, nq, np, k, nullptr, id_offset, top_scores, top_ids, stream);
    else if (exchange == 2) rc = evdr_peer_sharded_topk(h, rows, np, nq, np, k, id_offset, top_scores, top_ids, stream);
    else rc = evdr_sharded_topk(h, rows, np, nq, np, k, id_offset, top_scores, top_ids, stream);
  }
  cudaFreeAsync(rows, st);
  return rc;
}

}  // extern "C"
