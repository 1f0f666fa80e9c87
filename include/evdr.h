/* evdr.h -- C ABI of libevdr.so: B200-native late-interaction (MaxSim) scoring.
 *
 * Drop-in boundary for ONE path of kimjy-st/Efficient-Visual-Document-Retrieval:
 * the ColPali-style MaxSim scorer, its backward, and the top-k / nDCG@k tail of
 * evaluation.  The reference has no FFI of its own; its "operator API" is the
 * Python signature `score_multi_vector_masked(Q, P, qmask, pmask, chunk_p)`
 * (evaluator/retrieval.py:166-213) imported by name by all 21 main*.py scripts.
 * The Python mirror of that signature (package `efficient-visual-document-
 * retrieval_b200`, module `retrieval`) binds exactly the entry points below via
 * ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions
 *  - every pointer named `*_dev` / documented "device" is a CUDA device pointer
 *    owned by the caller (a torch allocation); the library allocates nothing
 *    persistent except the opaque handle;
 *  - all launches go to the caller's stream (`stream` is a cudaStream_t passed as
 *    void*; NULL = legacy default stream); no hidden synchronisation unless the
 *    function says so;
 *  - every function returns 0 on success, <0 on error (enum evdr_status) and
 *    records a thread-local message retrievable with evdr_last_error();
 *  - no C++ exceptions and no torch types cross this boundary;
 *  - there is NO CPU fallback: every compute entry point fails with
 *    EVDR_ERR_CUDA when no sm_100 device is present.
 */
#ifndef EVDR_H_
#define EVDR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define EVDR_API
#else
#define EVDR_API __attribute__((visibility("default")))
#endif

typedef struct evdr_handle evdr_handle;

enum evdr_status {
  EVDR_OK = 0,
  EVDR_ERR_INVALID = -1,     /* bad argument / shape */
  EVDR_ERR_CUDA = -2,        /* CUDA runtime / driver error, or no sm_100 device */
  EVDR_ERR_UNSUPPORTED = -3, /* shape outside what the kernel family covers */
  EVDR_ERR_NCCL = -4,        /* NCCL not loadable or collective failed */
  EVDR_ERR_STATE = -5        /* call order (e.g. sharded op before comm_init) */
};

enum evdr_dtype { EVDR_F32 = 0, EVDR_BF16 = 1 };

/* Packed page stream granularity: every page occupies a whole number of
 * EVDR_CHUNK-token slots in the bf16 token stream. */
#define EVDR_CHUNK 16
/* Tensor-core path operand geometry (ColPali): 128-d tokens, queries padded to
 * 32 token slots (one warp of epilogue lanes per query). */
#define EVDR_TC_DIM 128
#define EVDR_TC_LQ 32

EVDR_API int evdr_version(void);
EVDR_API const char* evdr_last_error(void);
/* Name and template arguments of the MaxSim scoring kernel the calling thread launched last (thread-local, ""
 * before the first launch): which flavour the dispatch picked, for tests, benchmarks and bug reports. */
EVDR_API const char* evdr_last_kernel(void);

/* Handle: per (thread, device); holds the device ordinal, SM count, the driver
 * entry point used to encode TMA descriptors and, after evdr_comm_init, the NCCL
 * communicator.  A handle is single-threaded: one host thread (and, for the peer exchange, one
 * stream) at a time.  Stream-ordered scratch comes from a memory pool owned by the handle
 * (release threshold 2 GiB, EVDR_POOL_KEEP_MB); the device's default pool is not touched. */
EVDR_API int evdr_create(int device, evdr_handle** out);
EVDR_API int evdr_destroy(evdr_handle* h);
EVDR_API int evdr_sm_count(const evdr_handle* h);

/* ------------------------------------------------------------------------
 * A1  score_multi_vector_masked -- general exact path (evaluator/retrieval.py:166-213)
 *
 *   out[q,c] = has[c] * sum_n qmask[q,n] * max_m ( pmask[c,m] ? <Q[q,n,:],P[c,m,:]> : -1e4 )
 *
 * fp32 FMA accumulation on the operands as given (bf16 inputs are up-cast like the
 * reference's `.to(float32)`, :176-177).  Any nq, np, lq, lp >= 1 and d % 4 == 0,
 * arbitrary mask patterns (uint8, non-zero = valid).  `argmax_dev` (optional,
 * int32, PAGE-MAJOR [np, nq*lq]) receives the lowest page-token index attaining
 * each max (torch.max's tie rule), for the backward.  `rowmax_ws_dev` is
 * caller-provided scratch of np*nq*lq floats (same layout).  np <= 65535 per call.
 * ------------------------------------------------------------------------ */
EVDR_API int evdr_maxsim_fwd(evdr_handle* h,
                             const void* Q_dev, int q_dtype,
                             const void* P_dev, int p_dtype,
                             const uint8_t* qmask_dev, const uint8_t* pmask_dev,
                             int nq, int np, int lq, int lp, int d,
                             float* rowmax_ws_dev, int32_t* argmax_dev,
                             float* out_dev, void* stream);

/* A5  backward of A1 (autograd of evaluator/retrieval.py:166-213):
 *   dP[c,m,:] = sum_{q,n} g[q,c] qmask[q,n] has[c] [m == argmax[q,c,n]] Q[q,n,:]
 *   dQ[q,n,:] = sum_c     g[q,c] qmask[q,n] has[c] P[c, argmax[q,c,n], :]
 * Deterministic (no floating-point atomics): one CTA owns one page's dP tile and
 * accumulates query tokens in index order.  dQ_dev may be NULL. */
EVDR_API int evdr_maxsim_bwd(evdr_handle* h,
                             const float* gout_dev,
                             const void* Q_dev, int q_dtype,
                             const void* P_dev, int p_dtype,
                             const uint8_t* qmask_dev, const uint8_t* pmask_dev,
                             const int32_t* argmax_dev,
                             int nq, int np, int lq, int lp, int d,
                             float* dP_dev, float* dQ_dev, void* stream);

/* ------------------------------------------------------------------------
 * Packed page stream (the HBM layout of the tensor-core path)
 *
 * evdr_pack_plan: chunk_off[c] = first 16-token chunk of page c in the stream,
 *   chunk_off[np] = total chunks.  varlen == 0: every page gets ceil(lp/16)
 *   chunks (offsets are arithmetic, no host sync needed to size the stream);
 *   varlen != 0: a page gets ceil((last valid index + 1)/16) chunks, at least 1.
 * evdr_pack_pages: writes the bf16 stream.  Masked / padded slots inside a page's
 *   chunks are filled with a copy of a valid token of the same page (nearest
 *   preceding valid token, else the first valid one), so they can never change a
 *   max; a page with no valid token becomes all-zero tokens (score exactly 0,
 *   as evaluator/retrieval.py:192,204).  normalize != 0 fuses the reference's
 *   `l2_normalize(P * pmask[...,None])` (utils/preprocess_data.py:8-9,
 *   mainv2_iter_listwise.py:280) in fp32 before the bf16 rounding.
 * k_blocks: 1 = plain bf16 rows of 128; 3 = the bf16x3 split for near-fp32 fidelity: page AND query rows are
 *   [hi | lo] of 256 bf16 (hi = bf16(x), lo = bf16(x - hi)); the kernel runs the three K = 128 products
 *   Qhi.Phi + Qhi.Plo + Qlo.Phi into one accumulator = <Q,P> up to a 2^-18 relative term.
 * ------------------------------------------------------------------------ */
EVDR_API int evdr_pack_plan(evdr_handle* h, const uint8_t* pmask_dev, int np, int lp,
                            int varlen, int32_t* chunk_off_dev, void* stream);
EVDR_API int evdr_pack_pages(evdr_handle* h, const void* P_dev, int p_dtype,
                             const uint8_t* pmask_dev, const int32_t* chunk_off_dev,
                             int np, int lp, int d, int normalize, int k_blocks,
                             void* tokens_bf16_dev, void* stream);
/* Ragged input (the on-disk layout: object arrays of (Li, 128) tokens with attention / image masks,
 * utils/preprocess_data.py:37-105, schema preprocess/split_data.py:29-36): one flat token array
 * [T, 128] + flat uint8 mask [T] (valid & attn & img) + int64 tok_off[np+1]; no padded (N, Lmax, D)
 * intermediate is ever built.  The plan is always varlen; max_len = longest page (sizes scratch). */
EVDR_API int evdr_pack_plan_ragged(evdr_handle* h, const uint8_t* mask_flat_dev, const int64_t* tok_off_dev,
                                   int np, int32_t* chunk_off_dev, void* stream);
EVDR_API int evdr_pack_pages_ragged(evdr_handle* h, const void* tok_flat_dev, int dtype,
                                    const uint8_t* mask_flat_dev, const int64_t* tok_off_dev,
                                    const int32_t* chunk_off_dev, int np, int max_len, int d,
                                    int normalize, int k_blocks, void* tokens_bf16_dev, void* stream);
/* Queries for the tensor-core path: bf16 [nq, 32, 128], masked / padded token
 * slots zeroed; qmask32 [nq, 32] uint8.  Requires lq <= 32, d == 128. */
EVDR_API int evdr_pack_queries(evdr_handle* h, const void* Q_dev, int q_dtype,
                               const uint8_t* qmask_dev, int nq, int lq, int d, int k_blocks,
                               void* Qb_bf16_dev, uint8_t* qmask32_dev, void* stream);

/* ------------------------------------------------------------------------
 * A1 on tensor cores: TMA-fed bf16 tcgen05.mma with fp32 TMEM accumulators; the
 * row-max over page tokens and the query-token sum are fused into the epilogue,
 * the similarity matrix never leaves the SM.  Persistent kernel, one CTA per SM.
 *
 *   tokens_dev : bf16 [n_tokens, 128] (k_blocks = 1) or [n_tokens, 256] (k_blocks = 3), n_tokens >= 16*chunk_off[np];
 *   Qb_dev     : bf16 [nq, 32, 128 | 256] from evdr_pack_queries with the same k_blocks;
 *   out_dev    : fp32 [nq, out_ld], column c = page c.
 * The base addresses must be 16-byte aligned (TMA).  `pages_per_item` <= 0 lets
 * the library pick the work-item size.
 * ------------------------------------------------------------------------ */
EVDR_API int evdr_maxsim_packed_fwd(evdr_handle* h,
                                    const void* Qb_dev, const uint8_t* qmask32_dev, int nq,
                                    const void* tokens_dev, int64_t n_tokens,
                                    const int32_t* chunk_off_dev, int np,
                                    int64_t total_chunks, int pages_per_item, int k_blocks,
                                    float* out_dev, int64_t out_ld, void* stream);

/* Queries longer than 32 tokens on the tensor-core path: the host splits a query into `groups`
 * 32-token groups scored as separate rows (exact: the max over page tokens is per query token);
 * evdr_group_sum adds the group rows back: out[q, c] = sum_g in[q*groups + g, c]. */
EVDR_API int evdr_group_sum(evdr_handle* h, const float* in_dev, int64_t in_ld, int nq, int groups,
                            int np, float* out_dev, int64_t out_ld, void* stream);

/* Gradient calls: tensor-core search + fp32 rescoring.
 * evdr_maxsim_packed_argmax runs the same tcgen05 kernel but emits, instead of scores, the packed
 *   slot attaining each max: argmax_slots_dev int32, PAGE-MAJOR [np, nq*lq] (position inside the
 *   page's chunks; chosen on bf16 operands with a 2^-19 selection tolerance).
 * evdr_maxsim_rescore maps every slot back to the page token it holds (fillers -> the valid token
 *   they duplicate), overwrites argmax_dev with token indices, recomputes exactly those
 *   <Q[q,n],P[c,m*]> in fp32 on the ORIGINAL operands and reduces them like evdr_maxsim_fwd.  The
 *   result equals the exact path wherever bf16 and fp32 agree on the winner (elsewhere it is lower
 *   by at most the bf16 near-tie gap); evdr_maxsim_bwd consumes the same argmax_dev. */
EVDR_API int evdr_maxsim_packed_argmax(evdr_handle* h,
                                       const void* Qb_dev, const uint8_t* qmask32_dev, int nq, int lq,
                                       const void* tokens_dev, int64_t n_tokens,
                                       const int32_t* chunk_off_dev, int np, int64_t total_chunks,
                                       int k_blocks, int32_t* argmax_slots_dev, void* stream);
EVDR_API int evdr_maxsim_rescore(evdr_handle* h,
                                 const void* Q_dev, int q_dtype, const void* P_dev, int p_dtype,
                                 const uint8_t* qmask_dev, const uint8_t* pmask_dev,
                                 int32_t* argmax_dev, int nq, int np, int lq, int lp, int d,
                                 float* rowmax_ws_dev, float* out_dev, void* stream);

/* The forward of a gradient call in ONE library call: evdr_pack_queries(k_blocks=3) + evdr_pack_plan +
 * evdr_pack_pages(k_blocks=3) + one tcgen05 pass on a stream-ordered
 * workspace.  rescore == 0: the pass yields scores AND winners (bf16x3 contraction: <Q,P> to ~2^-18
 * relative, measured <= 5e-6 on the scores); rescore != 0: winners only, their dot products are recomputed
 * exactly in fp32 on the original operands (evdr_maxsim_rescore) (the student tensor changes every step, mainv2_iter_listwise.py:280,287, so nothing can be
 * cached and the host-side issue time is paid per step).  Q (nq, lq, 128) with lq <= 32 or lq in
 * {64, 96, 128} (zero-padded, masked); argmax_dev int32 [np, nq*lq] receives the winning page-token
 * indices for evdr_maxsim_bwd; out_dev fp32 [nq, np]. */
EVDR_API int evdr_maxsim_train_fwd(evdr_handle* h, const void* Q_dev, int q_dtype, const void* P_dev,
                                   int p_dtype, const uint8_t* qmask_dev, const uint8_t* pmask_dev,
                                   int nq, int np, int lq, int lp, int d, int rescore, int32_t* argmax_dev,
                                   float* out_dev, void* stream);
/* The pieces of the rescore == 0 route: evdr_maxsim_packed_argmax_scores is evdr_maxsim_packed_argmax that
 * ALSO writes the scores of the same pass (out_dev fp32 [nq, out_ld]); evdr_slots_to_tokens turns packed slots
 * into page-token indices in place (fillers -> the valid token they duplicate; empty pages -> 0). */
EVDR_API int evdr_maxsim_packed_argmax_scores(evdr_handle* h, const void* Qb_dev, const uint8_t* qmask32_dev,
                                              int nq, int lq, const void* tokens_dev, int64_t n_tokens,
                                              const int32_t* chunk_off_dev, int np, int64_t total_chunks,
                                              int k_blocks, int32_t* argmax_slots_dev, float* out_dev,
                                              int64_t out_ld, void* stream);
EVDR_API int evdr_slots_to_tokens(evdr_handle* h, const uint8_t* pmask_dev, int32_t* argmax_dev, int np,
                                  int n_rows, int lp, void* stream);

/* A8  hard-token mining (mainv3_iter_liscore_QA_hardtoken.py:398-425): for every (query q, page
 * doc_idx[q, a]) the index of the VALID page token m maximising max_{valid n} <Q[q,n], P[c,m]> (the
 * column-max variant of MaxSim; ties -> lowest index).  Replaces a Python double loop with a host
 * sync per pair.  doc_idx / out_idx: int64 [nq, n_docs_per_query]. */
EVDR_API int evdr_hard_token_argmax(evdr_handle* h, const void* Q_dev, int q_dtype, const void* P_dev,
                                    int p_dtype, const uint8_t* qmask_dev, const uint8_t* pmask_dev,
                                    const int64_t* doc_idx_dev, int nq, int n_docs_per_query, int lq,
                                    int lp, int d, int64_t* out_idx_dev, void* stream);

/* ------------------------------------------------------------------------
 * A7  evaluation tail: top-k and trec_eval-style measures on device
 * (replaces the Nq*Np `.item()` dict build of mainv2_iter_listwise.py:312-318 and
 * mteb/pytrec_eval behind evaluator/retrieval.py:239-246).
 *
 * evdr_topk: per query the k best columns of scores[nq, ld] (first np columns),
 *   ordered by score descending, ties by tie_rank[c] ascending when given (host
 *   passes the rank of the docid string in DESCENDING order to reproduce
 *   trec_eval), else by column index ascending.  ids are column + id_offset.
 *   Slots beyond np are (-inf, -1).  1 <= k <= 128.
 * evdr_topk_merge: same selection over explicit (score, id) candidate lists
 *   (ties by id ascending) -- the merge step after the all-gather.
 * evdr_retrieval_metrics: out[m][j][q] for m in {ndcg, map, recall, precision,
 *   mrr}, j over k_values, from ranked ids [nq, kmax] and CSR qrels (ids within a
 *   query need not be sorted).  Linear gain, log2 discount, double accumulators
 *   like trec_eval; the host takes the mean and round(., 5) like mteb.
 * ------------------------------------------------------------------------ */
EVDR_API int evdr_topk(evdr_handle* h, const float* scores_dev, int64_t ld, int nq, int np,
                       int k, const int32_t* tie_rank_dev, int64_t id_offset,
                       float* top_scores_dev, int64_t* top_ids_dev, void* stream);
EVDR_API int evdr_topk_merge(evdr_handle* h, const float* cand_scores_dev,
                             const int64_t* cand_ids_dev, int nq, int ncand, int k,
                             float* top_scores_dev, int64_t* top_ids_dev, void* stream);
EVDR_API int evdr_retrieval_metrics(evdr_handle* h, const int64_t* top_ids_dev, int kmax, int nq,
                                    const int32_t* qrel_off_dev, const int64_t* qrel_ids_dev,
                                    const int32_t* qrel_rels_dev,
                                    const int32_t* k_values_dev, int nk,
                                    double* out_dev, void* stream);

/* Fused form: top-kmax selection and the measures in ONE pass -- the CTA that produces a query's
 * final ranking also scores it (qrel ids are column + id_offset). */
EVDR_API int evdr_topk_metrics(evdr_handle* h, const float* scores_dev, int64_t ld, int nq, int np,
                               int kmax, const int32_t* tie_rank_dev, int64_t id_offset,
                               const int32_t* qrel_off_dev, const int64_t* qrel_ids_dev,
                               const int32_t* qrel_rels_dev, const int32_t* k_values_dev, int nk,
                               float* top_scores_dev, int64_t* top_ids_dev, double* out_dev,
                               void* stream);

/* ------------------------------------------------------------------------
 * Multi-GPU: pages sharded by contiguous range, queries replicated.  Each rank
 * runs its local scan + evdr_topk, then one ncclAllGather of nq*k (score, id)
 * pairs and a merge on every rank (identical result everywhere).
 * NCCL is resolved at run time from the process (torch's libnccl.so.2).
 * ------------------------------------------------------------------------ */
#define EVDR_NCCL_UNIQUE_ID_BYTES 128
EVDR_API int evdr_comm_unique_id(void* id_out_128B);
EVDR_API int evdr_comm_init(evdr_handle* h, const void* id_128B, int rank, int world);
EVDR_API int evdr_comm_world(const evdr_handle* h);
/* Scratch for the exchange is stream-ordered (cudaMallocAsync on `stream`). */
EVDR_API int evdr_sharded_topk(evdr_handle* h, const float* local_scores_dev, int64_t ld,
                               int nq, int np_local, int k, int64_t id_offset,
                               float* top_scores_dev, int64_t* top_ids_dev, void* stream);

/* Training collective: every rank scores the batch against ITS pages; the loss needs full
 * (B, Np) rows on every rank.  Pages are partitioned contiguously, the first np_total % world ranks
 * own one page more (rank r: np_total / world (+1) columns starting at r * (np_total / world) +
 * min(r, np_total % world)).  One ncclAllGather of the zero-padded column blocks, placed by page offset.
 * The backward of this step is the local column slice of d(loss)/d(scores): no collective. */
EVDR_API int evdr_allgather_score_rows(evdr_handle* h, const float* local_scores_dev, int64_t ld, int B,
                                       int np_total, float* out_dev, int64_t out_ld, void* stream);

/* Peer-memory exchange (no NCCL on the data path): every rank owns an exchange region shared through
 * CUDA IPC; ONE kernel reduces the local candidates of each query to the local top-k, copies them into
 * all peers' regions with remote stores over NVLink, releases per-query flags, acquires the peers' flags
 * and merges the world * k candidates into the result (same order as evdr_topk_merge: score descending,
 * id ascending).
 *   evdr_peer_region_init : allocate this rank's region for up to `max_candidates` = nq * k entries per
 *                           call and return its 64-byte cudaIpcMemHandle_t;
 *   evdr_peer_connect     : `all_ipc_handles` = world x 64 bytes in rank order (the host plumbing --
 *                           torch.distributed -- gathers them); opens every peer's region;
 *   evdr_peer_sharded_topk: collective, same arguments and result as evdr_sharded_topk.  nq <= 4096.
 * Contract: every rank makes the same calls with the same nq and k (checked on the device), on one
 * stream per handle.  The call sequence number lives on the device, so the kernel may be captured in a
 * CUDA graph.  Waits are bounded by EVDR_PEER_TIMEOUT_MS (default 30000; 0 = wait forever like NCCL): when
 * a peer never arrives (or arrives with another nq / k) the kernel does not trap -- the affected queries
 * get (-inf, -1) results, the launch completes, and every later evdr_peer_* / peer-mode search call on
 * the handle returns EVDR_ERR_STATE (the process stays alive and can fall back to the NCCL exchange). */
#define EVDR_IPC_HANDLE_BYTES 64
EVDR_API int evdr_peer_region_init(evdr_handle* h, int rank, int world, int max_candidates,
                                   void* ipc_handle_out_64B);
EVDR_API int evdr_peer_connect(evdr_handle* h, const void* all_ipc_handles);
EVDR_API int evdr_peer_world(const evdr_handle* h);
EVDR_API int evdr_peer_sharded_topk(evdr_handle* h, const float* local_scores_dev, int64_t ld,
                                    int nq, int np_local, int k, int64_t id_offset,
                                    float* top_scores_dev, int64_t* top_ids_dev, void* stream);

/* Fused search (A1 + A7 + the exchange in one library call; mainv2_iter_listwise.py:309-320 at retrieval
 * scale): packed queries -> local scan -> local top-k -> candidate exchange -> global top-k on every rank.
 * With at most 8 queries on long pages (one query group) and k <= 32 the top-k is folded into the scan's
 * epilogue: every epilogue warp keeps the k best (score, page) of its query in registers and only
 * k candidates per (query, CTA) leave the SM -- the (nq, np) score row is never written; other shapes
 * score into a stream-ordered scratch row and run evdr_topk.  Results are identical either way.
 *   exchange: 0 = none (single GPU), 1 = ncclAllGather + merge (needs evdr_comm_init),
 *             2 = peer-memory kernel (needs evdr_peer_connect).
 * ids are page + id_offset; slots beyond np are (-inf, -1).  All launches go to `stream`; the sequence is
 * CUDA-graph capturable. */
EVDR_API int evdr_sharded_search(evdr_handle* h, const void* Qb_dev, const uint8_t* qmask32_dev, int nq,
                                 const void* tokens_dev, int64_t n_tokens, const int32_t* chunk_off_dev,
                                 int np, int64_t total_chunks, int k, int64_t id_offset, int exchange,
                                 float* top_scores_dev, int64_t* top_ids_dev, void* stream);

/* ------------------------------------------------------------------------
 * A4  fused mask * l2_normalize (utils/preprocess_data.py:8-9 applied as
 * mainv2_iter_listwise.py:280) and its backward, fp32:
 *   y = (x*mask) / (||x*mask||_2 + 1e-12)
 * ------------------------------------------------------------------------ */
EVDR_API int evdr_l2norm_fwd(evdr_handle* h, const float* X_dev, const uint8_t* mask_dev,
                             int64_t n_tok, int d, float* Y_dev, void* stream);
EVDR_API int evdr_l2norm_bwd(evdr_handle* h, const float* X_dev, const uint8_t* mask_dev,
                             const float* gY_dev, int64_t n_tok, int d, float* gX_dev,
                             void* stream);

/* ------------------------------------------------------------------------
 * The (B, N) score-matrix consumers directly behind the scorer in a training step.
 * Scores are fp32 [B, ld] rows (first n columns used).  Everything is deterministic.
 *
 * evdr_rank_gap -- mainv3_iter_liscore_QA_hardtoken.py:368-385 in one launch:
 *   rank_x[b,i] = position of page i in the descending order of row b (the reference's
 *                 argsort(argsort(x, descending=True)); equal scores keep index order),
 *   gap = rank_t - rank_s (fp32, :371), G[i] = sum_b |gap[b,i]| (:372),
 *   topk_idx[b, 0..kk) = the teacher's kk best pages, best first (:377-378),
 *   aux_idx[b, 0..a)   = those of them with the largest |gap|, largest first (:379-382; equal gaps
 *                        keep teacher-rank order).
 *   Any output pointer may be NULL.  0 <= a <= kk <= min(n, 1024), B*n < 2^24.
 * evdr_liscore_loss -- criterion.py:123-142 (listwise_distillation_loss: top-k partial cross entropy
 *   between temperature softmaxes, times T^2) and criterion.py:81-83 (score_preserving_loss = MSE),
 *   forward AND d/d(sc_s) of each, one launch:  loss2[0] = listwise, loss2[1] = mse,
 *   grad_listwise / grad_mse fp32 [B, n].  row_loss_ws: scratch of 2*B floats.  k > n is an error like
 *   torch.topk's.
 * evdr_pairwise_loss -- criterion.py:96-110 (pairwise_distillation_loss = mean BCE-with-logits of all
 *   score differences against sigmoid(teacher differences)) and its gradient without the three
 *   (B, N, N) intermediates.  part_ws: scratch of B * ceil(n / 256) floats.
 * ------------------------------------------------------------------------ */
EVDR_API int evdr_rank_gap(evdr_handle* h, const float* sc_t_dev, int64_t ld_t, const float* sc_s_dev,
                           int64_t ld_s, int B, int n, int kk, int a, float* gap_dev, float* G_dev,
                           int32_t* rank_t_dev, int32_t* rank_s_dev, int64_t* topk_idx_dev,
                           int64_t* aux_idx_dev, void* stream);
EVDR_API int evdr_liscore_loss(evdr_handle* h, const float* sc_s_dev, int64_t ld_s, const float* sc_t_dev,
                               int64_t ld_t, int B, int n, int k, float temperature,
                               float* grad_listwise_dev, float* grad_mse_dev, float* row_loss_ws_dev,
                               float* loss2_dev, void* stream);
EVDR_API int evdr_pairwise_loss(evdr_handle* h, const float* sc_s_dev, int64_t ld_s, const float* sc_t_dev,
                                int64_t ld_t, int B, int n, float* grad_dev, float* part_ws_dev,
                                float* loss1_dev, void* stream);
/* evdr_lambda_loss -- criterion.py:148-189 (lambda_loss): pairs ordered by the TEACHER ranking, weight
 *   10 * |sigmoid(t_i) - sigmoid(t_j)| * |1/log2(rank_i + 2) - 1/log2(rank_j + 2)|, term
 *   softplus(-alpha * (s_i - s_j)) for the better-ranked i, summed over the batch and divided by
 *   N(N-1)/2 + eps; loss1[0] and d/d(sc_s).  part_ws: B * ceil(n / 256) floats.
 * evdr_ranknce_loss -- criterion.py:192-226 (ranknce_loss): InfoNCE with the teacher's best page as the
 *   positive (cross entropy of sc_s / T) + lambda_weight * mean over adjacent pairs of the teacher order of
 *   sigmoid(t_k - t_k+1) * softplus(-(s_k - s_k+1)); loss3 = {total, infonce, ranking}, d/d(sc_s) of the
 *   total.  row_loss_ws: 2*B floats.  (n == 1: the ranking term is 0 where the reference's empty mean is NaN.) */
EVDR_API int evdr_lambda_loss(evdr_handle* h, const float* sc_s_dev, int64_t ld_s, const float* sc_t_dev,
                              int64_t ld_t, int B, int n, float alpha, float eps, float* grad_dev,
                              float* part_ws_dev, float* loss1_dev, void* stream);
EVDR_API int evdr_ranknce_loss(evdr_handle* h, const float* sc_s_dev, int64_t ld_s, const float* sc_t_dev,
                               int64_t ld_t, int B, int n, float temperature, float lambda_weight,
                               float* grad_dev, float* row_loss_ws_dev, float* loss3_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EVDR_H_ */
