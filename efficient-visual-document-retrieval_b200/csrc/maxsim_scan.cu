// maxsim_scan.cu -- K1s: the streaming single-query scan (1 or 2 queries) with the MMA operands SWAPPED.
//
//   out[q,c] = sum_n qmask[q,n] * max_{m in page c} <Q[q,n,:], P[c,m,:]>      (evaluator/retrieval.py:166-213)
//
// maxsim_tc.cu puts the query tokens on the M side of tcgen05.mma (M = 128 rows = 4 queries x 32 tokens,
// N = 256 page tokens): with ONE query 96 of the 128 rows are zero padding, i.e. 4x the tensor work the
// problem needs.  The scan is HBM-bound, but on B200 it is also POWER-bound: back to back it pulls ~1000 W,
// the SM clock is throttled to ~0.75 GHz and at that clock 1024 tensor cycles per 64 KB tile become the
// limit (6.3-6.5 TB/s sustained vs 7.2 TB/s in short bursts, profiles/r2_power_trace.log).  Here the PAGE
// tokens are the M side (128 per MMA, straight from the TMA ring) and the query tokens the N side
// (N = 32 per query): 4x (one query) / 2x (two queries) fewer tensor cycles and TMEM columns per byte.
//
//   D[page token (TMEM lane), query token (column)]  ->  the max over page tokens is a max ACROSS lanes:
//   each epilogue thread keeps a running elementwise max of its lane's 32 columns over the tiles of a
//   page (32 FMNMX per tile, no cross-lane work); at a page end 32 redux.sync.max.f32 fold the warp's
//   lanes, the four warps meet through 1 KB of shared memory, and one warp per query applies qmask, sums
//   the 32 tokens, stores the score and feeds the fused top-k list.
//
// Same packed page stream, scheduler (dynamic items, self-resetting counter), TMA ring and fused top-k
// candidate format as maxsim_tc.cu.  Warp roles: 0 = TMA producer + scheduler, 1 = MMA issuer, 2..5 = epilogue.
#include "common.cuh"
#include "ptx.cuh"

namespace evdr {

using namespace ptx;

constexpr int SCAN_TM = 128;                         // page tokens per MMA / per ring stage
constexpr int SCAN_STAGES = 6;                       // 6 x 32 KB ring
constexpr int SCAN_STAGE_BYTES = SCAN_TM * EVDR_TC_DIM * 2;
constexpr int SCAN_HALF_BYTES = SCAN_TM * 64 * 2;    // one 64-wide K sub-tile of a stage
constexpr int SCAN_NBUF = 4;                         // TMEM accumulator buffers of 128 columns each
constexpr int SCAN_NSLOT = 4;
constexpr int SCAN_CH = SCAN_TM / EVDR_CHUNK;        // 8 chunks per tile

struct ScanParams {
  const uint8_t* qmask32;     // [nq, 32]
  const int32_t* chunk_off;   // [np + 1]
  float* out;                 // [nq, out_ld] or nullptr
  long long out_ld;
  int nq, np;
  int ppi, n_items;
  int* sched;
  float* cand_s;              // fused top-k (tk_k > 0): [nq][gridDim.x][tk_k]
  long long* cand_i;
  int tk_k;
  long long id_offset;
};

__device__ __forceinline__ float redux_max_f32(float v) {
  float r;
  asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ void epi_bar_sync() {      // the 128 epilogue threads only
  asm volatile("bar.sync 1, 128;" ::: "memory");
}

template <int NQ>
struct ScanCfg {
  static constexpr int N = 32 * NQ;                              // MMA N = query tokens
  static constexpr int Q_HALF_BYTES = N * 64 * 2;
  static constexpr int Q_BYTES = 2 * Q_HALF_BYTES;               // 8 KB per query
  static constexpr int Q_ALLOC = ((Q_BYTES + 1023) / 1024) * 1024;
  static constexpr int P_BYTES = SCAN_STAGES * SCAN_STAGE_BYTES;
  static constexpr int BAR_OFF = Q_ALLOC + P_BYTES;
  // full[S], empty[S], q_full, t_full[NBUF], t_empty[NBUF], tmem ptr, slot_full[NSLOT], slot_empty[NSLOT], items
  static constexpr int BAR_BYTES = 8 * (2 * SCAN_STAGES + 1 + 2 * SCAN_NBUF) + 8 + 8 * 2 * SCAN_NSLOT + 4 * SCAN_NSLOT;
  static constexpr int XW_OFF = BAR_OFF + ((BAR_BYTES + 15) / 16) * 16;
  static constexpr int XW_BYTES = 2 * 4 * NQ * 32 * 4;           // [parity][warp][query][token]
  static constexpr int TOTAL = XW_OFF + XW_BYTES + 1024;
};

template <int NQ>
__global__ void __launch_bounds__(192, 1)
maxsim_scan_kernel(const __grid_constant__ CUtensorMap tmapQ, const __grid_constant__ CUtensorMap tmapP,
                   const ScanParams prm) {
  using S = ScanCfg<NQ>;
  constexpr uint32_t IDESC = make_idesc_bf16(SCAN_TM, S::N);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t sQ = smem_base;
  const uint32_t sP = smem_base + S::Q_ALLOC;
  const uint32_t bars = smem_base + S::BAR_OFF;
  auto bar_full = [&](int s) { return bars + 8u * s; };
  auto bar_empty = [&](int s) { return bars + 8u * (SCAN_STAGES + s); };
  const uint32_t bar_q_full = bars + 8u * (2 * SCAN_STAGES);
  auto bar_t_full = [&](int b) { return bars + 8u * (2 * SCAN_STAGES + 1 + b); };
  auto bar_t_empty = [&](int b) { return bars + 8u * (2 * SCAN_STAGES + 1 + SCAN_NBUF + b); };
  const uint32_t tmem_slot = bars + 8u * (2 * SCAN_STAGES + 1 + 2 * SCAN_NBUF);
  const uint32_t sched_base = tmem_slot + 8u;
  auto bar_s_full = [&](int s) { return sched_base + 8u * s; };
  auto bar_s_empty = [&](int s) { return sched_base + 8u * (SCAN_NSLOT + s); };
  volatile int* item_ring = reinterpret_cast<volatile int*>(smem_al + S::BAR_OFF + (8 * (2 * SCAN_STAGES + 1 + 2 * SCAN_NBUF) + 8 + 8 * 2 * SCAN_NSLOT));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_al + S::BAR_OFF + 8 * (2 * SCAN_STAGES + 1 + 2 * SCAN_NBUF));
  float* xw = reinterpret_cast<float*>(smem_al + S::XW_OFF);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_items = prm.n_items;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmapQ);
    prefetch_tmap(&tmapP);
    for (int s = 0; s < SCAN_STAGES; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), 1); }
    mbar_init(bar_q_full, 1);
    for (int b = 0; b < SCAN_NBUF; ++b) { mbar_init(bar_t_full(b), 1); mbar_init(bar_t_empty(b), 4); }
    for (int s = 0; s < SCAN_NSLOT; ++s) { mbar_init(bar_s_full(s), 1); mbar_init(bar_s_empty(s), 1 + 4); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================================================== TMA producer + tile scheduler
    if (lane == 0) {
      const uint64_t pol = policy_evict_first();                  // every page tile is read exactly once
      // the query tile (N rows; rows beyond nq * 32 are zero-filled by TMA) stays resident
      mbar_expect_tx(bar_q_full, S::Q_BYTES);
      tma_load_2d(sQ, &tmapQ, bar_q_full, 0, 0);
      tma_load_2d(sQ + S::Q_HALF_BYTES, &tmapQ, bar_q_full, 64, 0);
      int stage = 0, phase = 0, sslot = 0, sphase = 0;
      for (;;) {
        mbar_wait(bar_s_empty(sslot), sphase ^ 1);
        const int it = atomicAdd(prm.sched, 1);
        if (it == n_items + (int)gridDim.x - 1) atomicExch(prm.sched, 0);   // last terminal draw: slot back to 0
        item_ring[sslot] = it;
        mbar_arrive(bar_s_full(sslot));
        if (++sslot == SCAN_NSLOT) { sslot = 0; sphase ^= 1; }
        if (it >= n_items) break;
        const int p0 = it * prm.ppi, p1 = min(p0 + prm.ppi, prm.np);
        const int c0 = prm.chunk_off[p0], c1 = prm.chunk_off[p1];
        const int ntile = (c1 - c0 + SCAN_CH - 1) / SCAN_CH;
        for (int t = 0; t < ntile; ++t) {
          mbar_wait(bar_empty(stage), phase ^ 1);
          mbar_expect_tx(bar_full(stage), SCAN_STAGE_BYTES);
          const int row = (c0 + t * SCAN_CH) * EVDR_CHUNK;
          const uint32_t dst = sP + stage * SCAN_STAGE_BYTES;
          tma_load_2d_hint(dst, &tmapP, bar_full(stage), 0, row, pol);
          tma_load_2d_hint(dst + SCAN_HALF_BYTES, &tmapP, bar_full(stage), 64, row, pol);
          if (++stage == SCAN_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (one thread)
    if (lane == 0) {
      int stage = 0, phase = 0, sslot = 0, sphase = 0;
      uint32_t tile_it = 0;
      mbar_wait(bar_q_full, 0);
      tc_fence_after();
      for (;;) {
        mbar_wait(bar_s_full(sslot), sphase);
        const int it = item_ring[sslot];
        mbar_arrive(bar_s_empty(sslot));
        if (++sslot == SCAN_NSLOT) { sslot = 0; sphase ^= 1; }
        if (it >= n_items) break;
        const int p0 = it * prm.ppi, p1 = min(p0 + prm.ppi, prm.np);
        const int ntile = (prm.chunk_off[p1] - prm.chunk_off[p0] + SCAN_CH - 1) / SCAN_CH;
        for (int t = 0; t < ntile; ++t, ++tile_it) {
          const int buf = tile_it % SCAN_NBUF;
          mbar_wait(bar_full(stage), phase);
          mbar_wait(bar_t_empty(buf), ((tile_it / SCAN_NBUF) & 1) ^ 1);
          tc_fence_after();
          const uint32_t a_base = sP + stage * SCAN_STAGE_BYTES;
          const uint32_t d_tmem = tmem_base + buf * 128;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint64_t da = make_sw128_desc(a_base + (k >> 2) * SCAN_HALF_BYTES + (k & 3) * 32);
            const uint64_t db = make_sw128_desc(sQ + (k >> 2) * S::Q_HALF_BYTES + (k & 3) * 32);
            mma_bf16_ss(d_tmem, da, db, IDESC, k > 0);
          }
          mma_commit(bar_t_full(buf));
          mma_commit(bar_empty(stage));
          if (++stage == SCAN_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ===================================================== epilogue warps: lane = page token
    const int quad = warp & 3;                              // TMEM lane quadrant of this warp
    const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
    const int my_ch = 2 * quad + (lane >> 4);               // chunk (of the tile's 8) holding this lane's token
    float qm[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) qm[q] = (q < prm.nq) ? (float)prm.qmask32[(size_t)q * EVDR_TC_LQ + lane] : 0.f;
    // fused top-k: warp q keeps query q's list (entry `lane`) -- NQ <= 4 queries, 4 epilogue warps
    float tk_s = -INFINITY, tk_thr = -INFINITY;
    int tk_p = -1;
    uint32_t tile_it = 0, fin_it = 0;
    int sslot = 0, sphase = 0;
    for (;;) {
      mbar_wait(bar_s_full(sslot), sphase);
      const int it = item_ring[sslot];
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_s_empty(sslot));
      if (++sslot == SCAN_NSLOT) { sslot = 0; sphase ^= 1; }
      if (it >= n_items) break;
      const int p0 = it * prm.ppi, p1 = min(p0 + prm.ppi, prm.np);
      const int c0 = prm.chunk_off[p0], c1 = prm.chunk_off[p1];
      const int ntile = (c1 - c0 + SCAN_CH - 1) / SCAN_CH;
      // page-end cache: lane l holds the item-relative end chunk of page p0 + 32 * cb + l
      int cb = 0;
      auto load_pe = [&](int b) {
        const int pp = p0 + 32 * b + lane;
        return (pp < p1) ? (prm.chunk_off[pp + 1] - c0) : 0x7fffffff;
      };
      int pe_cache = load_pe(0);
      int p = p0;
      int pe = __shfl_sync(0xffffffffu, pe_cache, 0);
      float rm[NQ][32];
#pragma unroll
      for (int q = 0; q < NQ; ++q)
#pragma unroll
        for (int n = 0; n < 32; ++n) rm[q][n] = -INFINITY;

      for (int t = 0; t < ntile; ++t, ++tile_it) {
        const int buf = tile_it % SCAN_NBUF;
        const int ch0 = t * SCAN_CH;
        mbar_wait(bar_t_full(buf), (tile_it / SCAN_NBUF) & 1);
        tc_fence_after();
        float v[NQ][32];
#pragma unroll
        for (int q = 0; q < NQ; ++q) tmem_ld32_issue(t_lane + buf * 128 + q * 32, v[q]);
#pragma unroll
        for (int q = 0; q < NQ; ++q) tmem_wait_ld_tied(v[q]);
        // the accumulator buffer goes back to the MMA warp before any page logic
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_t_empty(buf));

        int lo = 0;                                         // chunks of this tile below `lo` are consumed
        while (pe - ch0 <= SCAN_CH) {                       // page p ends inside (or at the end of) this tile
          const int e = pe - ch0;
          const bool in = (my_ch >= lo) && (my_ch < e);
#pragma unroll
          for (int q = 0; q < NQ; ++q)
#pragma unroll
            for (int n = 0; n < 32; ++n) rm[q][n] = in ? fmaxf(rm[q][n], v[q][n]) : rm[q][n];
          // ---- page complete: max over the warp's 32 page-token lanes, then over the 4 warps
          float* xbuf = xw + (fin_it & 1) * (4 * NQ * 32);
#pragma unroll
          for (int q = 0; q < NQ; ++q) {
            float mine = -INFINITY;
#pragma unroll
            for (int n = 0; n < 32; ++n) {
              const float m = redux_max_f32(rm[q][n]);
              mine = (lane == n) ? m : mine;
              rm[q][n] = -INFINITY;
            }
            xbuf[(quad * NQ + q) * 32 + lane] = mine;
          }
          epi_bar_sync();
          if (quad < NQ && quad < prm.nq) {                 // warp q finishes query q
            const int q = quad;
            float mx = xbuf[(0 * NQ + q) * 32 + lane];
#pragma unroll
            for (int w = 1; w < 4; ++w) mx = fmaxf(mx, xbuf[(w * NQ + q) * 32 + lane]);
            float w = qm[0];
#pragma unroll
            for (int j = 1; j < NQ; ++j) w = (q == j) ? qm[j] : w;
            const float s = warp_sum(mx * w);
            if (prm.out != nullptr && lane == 0) prm.out[(size_t)q * prm.out_ld + p] = s;
            if (prm.tk_k > 0 && s > tk_thr) {
              // pages arrive in ascending order within a CTA: strict '>' keeps (score desc, page asc)
              const int pos = __popc(__ballot_sync(0xffffffffu, tk_s >= s));
              const float us = __shfl_up_sync(0xffffffffu, tk_s, 1);
              const int up = __shfl_up_sync(0xffffffffu, tk_p, 1);
              if (lane == pos) { tk_s = s; tk_p = p; }
              else if (lane > pos) { tk_s = us; tk_p = up; }
              tk_thr = __shfl_sync(0xffffffffu, tk_s, prm.tk_k - 1);
            }
          }
          ++fin_it;
          lo = e;
          ++p;
          const int b = (p - p0) >> 5;
          if (b != cb) { pe_cache = load_pe(b); cb = b; }
          pe = __shfl_sync(0xffffffffu, pe_cache, (p - p0) & 31);
        }
        if (lo == 0) {                                      // whole tile inside one page: no predicate
#pragma unroll
          for (int q = 0; q < NQ; ++q)
#pragma unroll
            for (int n = 0; n < 32; ++n) rm[q][n] = fmaxf(rm[q][n], v[q][n]);
        } else if (lo < SCAN_CH) {
          const bool in = my_ch >= lo;
#pragma unroll
          for (int q = 0; q < NQ; ++q)
#pragma unroll
            for (int n = 0; n < 32; ++n) rm[q][n] = in ? fmaxf(rm[q][n], v[q][n]) : rm[q][n];
        }
      }
    }
    if (prm.tk_k > 0 && quad < NQ && quad < prm.nq && lane < prm.tk_k) {
      const size_t o = ((size_t)quad * gridDim.x + blockIdx.x) * prm.tk_k + lane;
      prm.cand_s[o] = tk_s;
      prm.cand_i[o] = tk_p >= 0 ? (long long)tk_p + prm.id_offset : -1LL;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    __syncwarp();
    tmem_dealloc<512>(tmem_base);
  }
}

static int encode_2d(evdr_handle* h, CUtensorMap* tm, const void* base, uint64_t rows, uint32_t box_rows,
                     const char* what) {
  cuuint64_t dims[2] = {(cuuint64_t)EVDR_TC_DIM, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)EVDR_TC_DIM * 2};
  cuuint32_t box[2] = {64u, box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = h->encode_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box,
                               estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(EVDR_ERR_CUDA, "cuTensorMapEncodeTiled(%s) failed with CUresult %d", what, (int)r);
  return EVDR_OK;
}

template <int NQ>
static int launch_scan(evdr_handle* h, const void* Qb, const void* tokens, int64_t n_tokens, ScanParams prm,
                       int grid, cudaStream_t st) {
  using S = ScanCfg<NQ>;
  CUtensorMap tmQ, tmP;
  int rc = encode_2d(h, &tmQ, Qb, (uint64_t)prm.nq * EVDR_TC_LQ, S::N, "queries");
  if (rc != EVDR_OK) return rc;
  rc = encode_2d(h, &tmP, tokens, (uint64_t)n_tokens, SCAN_TM, "page tokens");
  if (rc != EVDR_OK) return rc;
  auto kern = maxsim_scan_kernel<NQ>;
  if ((size_t)S::TOTAL > h->smem_optin)
    return set_error(EVDR_ERR_UNSUPPORTED, "scan kernel needs %d B of shared memory", S::TOTAL);
  EVDR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
  prm.sched = h->sched_ring + (h->sched_next.fetch_add(1, std::memory_order_relaxed) % EVDR_SCHED_RING);
  note_kernel("maxsim_scan_kernel<NQ=%d> grid=%d topk=%d", NQ, grid, prm.tk_k);
  kern<<<grid, 192, S::TOTAL, st>>>(tmQ, tmP, prm);
  EVDR_LAUNCH_CHECK("maxsim_scan_kernel");
  return EVDR_OK;
}

// 1 or 2 packed queries against the packed stream: scores (out may be null) and / or the per-CTA top-k
// candidate lists (tk_k > 0).  *n_cta receives the grid size (the candidate buffers hold nq * n_cta * k).
int scan_swapped(evdr_handle* h, const void* Qb, const uint8_t* qmask32, int nq, const void* tokens, int64_t n_tokens,
                 const int32_t* chunk_off, int np, int64_t total_chunks, int pages_per_item, float* out,
                 int64_t out_ld, int tk_k, int64_t id_offset, float* cand_s, long long* cand_i, int* n_cta,
                 cudaStream_t st) {
  EVDR_CHECK_ARG(nq >= 1 && nq <= 2, "scan_swapped: nq=%d outside [1,2]", nq);
  EVDR_CHECK_ARG(tk_k >= 0 && tk_k <= 32, "scan_swapped: k=%d outside [0,32]", tk_k);
  ScanParams prm = {};
  prm.qmask32 = qmask32;
  prm.chunk_off = chunk_off;
  prm.out = out;
  prm.out_ld = out_ld;
  prm.nq = nq;
  prm.np = np;
  prm.cand_s = cand_s;
  prm.cand_i = cand_i;
  prm.tk_k = tk_k;
  prm.id_offset = id_offset;
  // work items: ~64K tokens (amortises the per-item page-end cache refill), never fewer items than CTAs,
  // and for small corpora an item size that balances the tail (see maxsim_tc.cu)
  int ppi = pages_per_item;
  if (ppi <= 0) {
    const double avg_tok = (double)total_chunks * EVDR_CHUNK / (double)np;
    ppi = (int)(65536.0 / avg_tok + 0.5);
    if (ppi < 1) ppi = 1;
    const double tiles_per_cta = (double)total_chunks / SCAN_CH / (double)h->sm_count;
    int ppi_bal = (int)(sqrt(4.0 * tiles_per_cta) * SCAN_TM / avg_tok + 0.5);
    if (ppi_bal < 1) ppi_bal = 1;
    if (ppi > ppi_bal) ppi = ppi_bal;
    while (ppi > 1 && (long long)((np + ppi - 1) / ppi) < (long long)h->sm_count) ppi = (ppi + 1) / 2;
  }
  if (ppi > np) ppi = np;
  prm.ppi = ppi;
  prm.n_items = (np + ppi - 1) / ppi;
  const int grid = prm.n_items < h->sm_count ? prm.n_items : h->sm_count;
  if (n_cta) *n_cta = grid;
  if (nq == 1) return launch_scan<1>(h, Qb, tokens, n_tokens, prm, grid, st);
  return launch_scan<2>(h, Qb, tokens, n_tokens, prm, grid, st);
}

}  // namespace evdr
