// train.cu -- the forward of a gradient call as ONE library call (evdr_maxsim_train_fwd): bf16x3
// packing of both operands and ONE tcgen05 pass that yields the scores and the winning page token of
// every (query token, page) -- optionally followed by exact fp32 rescoring of the winners -- on a
// single stream-ordered workspace.  A training step of the reference calls the
// scorer on a new student tensor every iteration (mainv2_iter_listwise.py:280,287), so host-side
// overhead (allocations, Python <-> C transitions) is paid 23 460 times per run; composed from Python
// the same six launches cost ~0.16 ms of issue time against ~0.25 ms of GPU time.
#include "common.cuh"

using namespace evdr;

namespace {
inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
}

extern "C" int evdr_maxsim_train_fwd(evdr_handle* h, const void* Q, int q_dtype, const void* P, int p_dtype,
                                     const uint8_t* qmask, const uint8_t* pmask, int nq, int np, int lq,
                                     int lp, int d, int rescore, int32_t* argmax, float* out, void* stream) {
  EVDR_REQUIRE_HANDLE(h);
  EVDR_CHECK_ARG(Q && P && qmask && pmask && argmax && out, "evdr_maxsim_train_fwd: null pointer");
  EVDR_CHECK_ARG(nq >= 1 && np >= 1 && lq >= 1 && lp >= 1, "evdr_maxsim_train_fwd: bad shape");
  if (d != EVDR_TC_DIM)
    return set_error(EVDR_ERR_UNSUPPORTED, "evdr_maxsim_train_fwd: d=%d (tensor-core path needs 128)", d);
  int groups = 1;
  if (lq > EVDR_TC_LQ) {
    if (lq % EVDR_TC_LQ != 0 || lq > 4 * EVDR_TC_LQ)
      return set_error(EVDR_ERR_UNSUPPORTED,
                       "evdr_maxsim_train_fwd: lq=%d (longer queries must be padded to 64, 96 or 128 tokens)", lq);
    groups = lq / EVDR_TC_LQ;
  }
  const int lq_g = lq / groups, nq_g = nq * groups;
  const size_t R = (size_t)nq * lq;
  const size_t cpp = (size_t)(lp + EVDR_CHUNK - 1) / EVDR_CHUNK;        // rectangular plan: no host sync
  const size_t total_chunks = (size_t)np * cpp;
  // rescore != 0: workspace for the per-(page, row) fp32 dot products; else: per-group score rows (lq > 32)
  const size_t rowmax_bytes = rescore ? (size_t)np * R * sizeof(float)
                                      : (groups > 1 ? (size_t)nq_g * np * sizeof(float) : 0);
  if ((size_t)np * R * sizeof(int32_t) > ((size_t)1 << 30) || np > 65535)
    return set_error(EVDR_ERR_UNSUPPORTED, "evdr_maxsim_train_fwd: np*nq*lq too large for one call");
  const size_t b_q = align256((size_t)nq_g * EVDR_TC_LQ * EVDR_TC_DIM * 2 * 2);   // [hi | lo] rows
  const size_t b_m = align256((size_t)nq_g * EVDR_TC_LQ);
  const size_t b_o = align256(((size_t)np + 1) * sizeof(int32_t));
  const size_t b_t = align256(total_chunks * EVDR_CHUNK * EVDR_TC_DIM * 2 * 2);
  const size_t b_r = align256(rowmax_bytes);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  unsigned char* ws = nullptr;
  EVDR_CUDA(evdr::ws_alloc(h, &ws, b_q + b_m + b_o + b_t + b_r, st));
  unsigned char* Qb = ws;
  uint8_t* qm32 = ws + b_q;
  int32_t* chunk_off = reinterpret_cast<int32_t*>(ws + b_q + b_m);
  unsigned char* tokens = ws + b_q + b_m + b_o;
  float* rowmax = reinterpret_cast<float*>(ws + b_q + b_m + b_o + b_t);
  int rc = evdr_pack_queries(h, Q, q_dtype, qmask, nq_g, lq_g, d, 3, Qb, qm32, stream);
  if (rc == EVDR_OK) rc = evdr_pack_plan(h, pmask, np, lp, 0, chunk_off, stream);
  if (rc == EVDR_OK) rc = evdr_pack_pages(h, P, p_dtype, pmask, chunk_off, np, lp, d, 0, 3, tokens, stream);
  if (rescore) {
    if (rc == EVDR_OK)
      rc = evdr_maxsim_packed_argmax(h, Qb, qm32, nq_g, lq_g, tokens, (int64_t)(total_chunks * EVDR_CHUNK),
                                     chunk_off, np, (int64_t)total_chunks, 3, argmax, stream);
    if (rc == EVDR_OK)
      rc = evdr_maxsim_rescore(h, Q, q_dtype, P, p_dtype, qmask, pmask, argmax, nq, np, lq, lp, d, rowmax, out,
                               stream);
  } else {
    // search AND score in the one tcgen05 pass (bf16x3: <Q,P> to ~2^-18 relative), then name the tokens
    float* rows = groups > 1 ? rowmax : out;
    if (rc == EVDR_OK)
      rc = evdr_maxsim_packed_argmax_scores(h, Qb, qm32, nq_g, lq_g, tokens, (int64_t)(total_chunks * EVDR_CHUNK),
                                            chunk_off, np, (int64_t)total_chunks, 3, argmax, rows, np, stream);
    if (rc == EVDR_OK && groups > 1) rc = evdr_group_sum(h, rows, np, nq, groups, np, out, np, stream);
    if (rc == EVDR_OK) rc = evdr_slots_to_tokens(h, pmask, argmax, np, (int)R, lp, stream);
  }
  cudaFreeAsync(ws, st);
  return rc;
}
