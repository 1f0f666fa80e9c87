// sharded.cu -- multi-GPU tail: local top-k -> ncclAllGather of (score, id) candidates over
// NVLink/NVSwitch -> merge on every rank.  Pages are sharded by contiguous range, queries
// replicated.  NCCL is resolved at run time with dlopen/dlsym so that the
// library carries no link-time NCCL dependency and shares torch's libnccl.so.2.
#include <dlfcn.h>

#include "common.cuh"

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
  static bool tried = false;
  if (tried) return api.ok ? &api : nullptr;
  tried = true;
  void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) return nullptr;
  api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(lib, "ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(lib, "ncclCommInitRank"));
  api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(lib, "ncclCommDestroy"));
  api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(lib, "ncclAllGather"));
  api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(lib, "ncclGetErrorString"));
  api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather && api.GetErrorString;
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

}  // namespace evdr

using namespace evdr;

extern "C" {

int evdr_destroy(evdr_handle* h) {
  if (!h) return EVDR_OK;
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
  if (!h->nccl_comm) return set_error(EVDR_ERR_STATE, "evdr_sharded_topk before evdr_comm_init");
  NcclApi* api = nccl_api();
  if (!api) return set_error(EVDR_ERR_NCCL, "libnccl.so.2 not loadable");
  const int world = h->world;
  const size_t per = (size_t)nq * k;
  const size_t rank_bytes = blob_bytes(per);
  // stream-ordered scratch: [send blob][recv blobs x world][cand ids world*per][cand scores world*per]
  unsigned char* ws = nullptr;
  EVDR_CUDA(cudaMallocAsync(&ws, rank_bytes * (2 * (size_t)world + 1) + 64, st));
  unsigned char* send = ws;
  unsigned char* recv = ws + rank_bytes;
  long long* cand_i = reinterpret_cast<long long*>(recv + rank_bytes * world);
  float* cand_s = reinterpret_cast<float*>(cand_i + (size_t)world * per);
  int64_t* loc_i = reinterpret_cast<int64_t*>(send);
  float* loc_s = reinterpret_cast<float*>(send + per * sizeof(long long));
  int rc = evdr_topk(h, local_scores, ld, nq, np_local, k, nullptr, id_offset, loc_s, loc_i, stream);
  if (rc != EVDR_OK) { cudaFreeAsync(ws, st); return rc; }
  nccl_result_t r = api->AllGather(send, recv, rank_bytes, kNcclChar,
                                   static_cast<nccl_comm_t>(h->nccl_comm), st);
  if (r != 0) {
    cudaFreeAsync(ws, st);
    return set_error(EVDR_ERR_NCCL, "ncclAllGather failed: %s", api->GetErrorString(r));
  }
  const int total = world * (int)per;
  unpack_gathered_kernel<<<(total + 255) / 256, 256, 0, st>>>(recv, world, nq, k, cand_s, cand_i);
  EVDR_LAUNCH_CHECK("unpack_gathered_kernel");
  rc = evdr_topk_merge(h, cand_s, reinterpret_cast<const int64_t*>(cand_i), nq, world * k, k,
                       top_scores, top_ids, stream);
  cudaFreeAsync(ws, st);
  return rc;
}

}  // extern "C"
