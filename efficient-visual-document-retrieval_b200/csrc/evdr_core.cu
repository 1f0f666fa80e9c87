// evdr_core.cu -- handle, error string, version.  No compute here.
#include "common.cuh"

namespace evdr {

char* tls_error_buf() {
  static thread_local char buf[1024] = {0};
  return buf;
}

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(tls_error_buf(), 1024, fmt, ap);
  va_end(ap);
  return code;
}

char* tls_kernel_buf() {
  static thread_local char buf[160] = {0};
  return buf;
}

void note_kernel(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(tls_kernel_buf(), 160, fmt, ap);
  va_end(ap);
}

}  // namespace evdr

extern "C" {

int evdr_version(void) { return EVDR_VERSION_INT; }

const char* evdr_last_kernel(void) { return evdr::tls_kernel_buf(); }

const char* evdr_last_error(void) { return evdr::tls_error_buf(); }

int evdr_create(int device, evdr_handle** out) {
  EVDR_CHECK_ARG(out != nullptr, "evdr_create: out is null");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return evdr::set_error(EVDR_ERR_CUDA,
                           "no CUDA device visible (%s); libevdr has no CPU fallback",
                           e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  EVDR_CHECK_ARG(device >= 0 && device < ndev, "evdr_create: device %d out of range [0,%d)", device, ndev);
  EVDR_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  EVDR_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return evdr::set_error(EVDR_ERR_CUDA,
                           "device %d is sm_%d%d; libevdr is built for sm_100a (B200) only",
                           device, prop.major, prop.minor);
  evdr_handle* h = new evdr_handle();   // value-initialised: all members zero
  h->device = device;
  h->sm_count = prop.multiProcessorCount;
  h->cc_major = prop.major;
  h->cc_minor = prop.minor;
  h->smem_optin = prop.sharedMemPerBlockOptin;
  h->world = 1;
  h->rank = 0;
  // cuTensorMapEncodeTiled through the runtime, so the library needs no link-time libcuda.
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || fn == nullptr) {
    delete h;
    return evdr::set_error(EVDR_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  }
  h->encode_tiled = reinterpret_cast<evdr_encode_tiled_fn>(fn);
  if (cudaMalloc(reinterpret_cast<void**>(&h->sched_ring), sizeof(int) * EVDR_SCHED_RING) != cudaSuccess) {
    delete h;
    return evdr::set_error(EVDR_ERR_CUDA, "cudaMalloc of the scheduler ring failed");
  }
  cudaMemset(h->sched_ring, 0, sizeof(int) * EVDR_SCHED_RING);   // every user hands its slot back zeroed
  // Stream-ordered scratch for small temporaries: a PRIVATE pool whose memory stays cached up to a bounded
  // threshold (a pool at the default threshold 0 gives everything back to the driver at every
  // synchronisation, which turns each later allocation into a millisecond-scale driver call -- a per-step
  // loss.item() made every scorer call 20+ ms).  EVDR_POOL_KEEP_MB overrides the 2 GiB default.
  {
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = device;
    if (cudaMemPoolCreate(&h->pool, &props) != cudaSuccess) {
      cudaFree(h->sched_ring);
      delete h;
      return evdr::set_error(EVDR_ERR_CUDA, "cudaMemPoolCreate failed");
    }
    uint64_t keep = 2048ull << 20;
    if (const char* e = getenv("EVDR_POOL_KEEP_MB")) keep = strtoull(e, nullptr, 10) << 20;
    cudaMemPoolSetAttribute(h->pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  *out = h;
  return EVDR_OK;
}

int evdr_sm_count(const evdr_handle* h) { return h ? h->sm_count : 0; }

}  // extern "C"
