// maxsim_tc.cu -- K1: fused MaxSim on 5th-gen tensor cores (sm_100a).
//
//   out[q,c] = sum_n qmask[q,n] * max_{m in page c} <Q[q,n,:], P[c,m,:]>
//
// (evaluator/retrieval.py:166-213 on the packed page stream, where masked slots hold
//  duplicates of valid tokens and empty pages are zero tokens -- see pack.cu.)
//
// Structure: persistent, warp-specialised, one CTA per SM, 192 threads:
//   warp 0      TMA producer  : cp.async.bulk.tensor of query tiles (A, resident) and page-token
//                               tiles (B, STAGES-deep ring) into 128B-swizzled shared memory
//   warp 1      MMA issuer    : one thread issues tcgen05.mma.kind::f16 (M=128, N=BN, K=16 x 8),
//                               fp32 accumulators double-buffered in TMEM (2 x BN columns)
//   warps 2..5  epilogue      : tcgen05.ld 32 columns at a time; lane = query token, so the max over
//                               page tokens is a per-thread running FMNMX3 chain; at a page's last
//                               chunk: * qmask, warp-shuffle sum over the 32 tokens of the query,
//                               one fp32 store.  The similarity matrix never leaves the SM.
//
// Work item = (page block pb, query group g): R resident A tiles (4 queries each) x all B tiles of
// the block.  Items are ordered g-fastest and dealt round-robin to CTAs so that concurrently
// running CTAs read the same page block (L2 hit) and HBM sees every page once.
#include "common.cuh"
#include "ptx.cuh"

namespace evdr {

using namespace ptx;

constexpr int TC_THREADS = 192;
constexpr int A_TILE_BYTES = 128 * 128 * 2;      // 128 query-token rows x 128 dims bf16
constexpr int A_HALF_BYTES = 128 * 64 * 2;       // one 64-wide K half (128-B rows)

struct TcParams {
  const uint8_t* qmask32;     // [nq, 32]
  const int32_t* chunk_off;   // [np + 1]
  float* out;                 // [nq, out_ld]
  long long out_ld;
  int nq, np;
  int nqt;                    // ceil(nq / 4) A tiles
  int ppi;                    // pages per item
  int n_pblocks, n_qgroups;
  int stream_hint;            // 1: page tiles are read once (evict-first)
};

template <int R, int BN, int STAGES>
struct TcSmem {
  static constexpr int B_STAGE_BYTES = BN * 128 * 2;
  static constexpr int B_HALF_BYTES = BN * 64 * 2;
  static constexpr int A_BYTES = R * A_TILE_BYTES;
  static constexpr int B_BYTES = STAGES * B_STAGE_BYTES;
  static constexpr int BAR_OFF = A_BYTES + B_BYTES;
  // barriers: full[STAGES], empty[STAGES], a_full, a_empty, tmem_full[2], tmem_empty[2], tmem ptr
  static constexpr int BAR_BYTES = 8 * (2 * STAGES + 6) + 8;
  static constexpr int TOTAL = BAR_OFF + BAR_BYTES + 1024;  // + alignment slack
};

template <int R, int BN, int STAGES>
__global__ void __launch_bounds__(TC_THREADS, 1)
maxsim_tc_kernel(const __grid_constant__ CUtensorMap tmapQ, const __grid_constant__ CUtensorMap tmapP,
                 const TcParams prm) {
  using S = TcSmem<R, BN, STAGES>;
  constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128
                            : (2 * BN <= 256) ? 256 : 512;
  static_assert(2 * BN <= 512, "two accumulator buffers must fit TMEM");
  static_assert(BN % 32 == 0 && BN >= 32 && BN <= 256, "BN");
  constexpr uint32_t IDESC = make_idesc_bf16(128, BN);
  constexpr int CH_PER_TILE = BN / EVDR_CHUNK;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = smem_base;
  const uint32_t sB = smem_base + S::A_BYTES;
  const uint32_t bars = smem_base + S::BAR_OFF;
  auto bar_full = [&](int s) { return bars + 8u * s; };
  auto bar_empty = [&](int s) { return bars + 8u * (STAGES + s); };
  const uint32_t bar_a_full = bars + 8u * (2 * STAGES);
  const uint32_t bar_a_empty = bars + 8u * (2 * STAGES + 1);
  auto bar_t_full = [&](int b) { return bars + 8u * (2 * STAGES + 2 + b); };
  auto bar_t_empty = [&](int b) { return bars + 8u * (2 * STAGES + 4 + b); };
  const uint32_t tmem_slot = bars + 8u * (2 * STAGES + 6);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmapQ);
    prefetch_tmap(&tmapP);
    for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), 1); }
    mbar_init(bar_a_full, 1);
    mbar_init(bar_a_empty, 1);
    for (int b = 0; b < 2; ++b) { mbar_init(bar_t_full(b), 1); mbar_init(bar_t_empty(b), 4); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int n_items = prm.n_pblocks * prm.n_qgroups;

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      const uint64_t pol = prm.stream_hint ? policy_evict_first() : policy_evict_last();
      int stage = 0, phase = 0, a_loads = 0, cur_g = -1;
      for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const int g = it % prm.n_qgroups, pb = it / prm.n_qgroups;
        const int p0 = pb * prm.ppi, p1 = min(p0 + prm.ppi, prm.np);
        const int c0 = prm.chunk_off[p0], c1 = prm.chunk_off[p1];
        const int ntile = ((c1 - c0) * EVDR_CHUNK + BN - 1) / BN;
        const int qt0 = g * R;
        const int rc = min(R, prm.nqt - qt0);
        if (g != cur_g) {
          if (a_loads > 0) mbar_wait(bar_a_empty, (a_loads - 1) & 1);
          mbar_expect_tx(bar_a_full, rc * A_TILE_BYTES);
          for (int r = 0; r < rc; ++r) {
            tma_load_2d(sA + r * A_TILE_BYTES, &tmapQ, bar_a_full, 0, (qt0 + r) * 128);
            tma_load_2d(sA + r * A_TILE_BYTES + A_HALF_BYTES, &tmapQ, bar_a_full, 64, (qt0 + r) * 128);
          }
          ++a_loads;
          cur_g = g;
        }
        for (int t = 0; t < ntile; ++t) {
          mbar_wait(bar_empty(stage), phase ^ 1);
          mbar_expect_tx(bar_full(stage), S::B_STAGE_BYTES);
          const int row = c0 * EVDR_CHUNK + t * BN;
          tma_load_2d_hint(sB + stage * S::B_STAGE_BYTES, &tmapP, bar_full(stage), 0, row, pol);
          tma_load_2d_hint(sB + stage * S::B_STAGE_BYTES + S::B_HALF_BYTES, &tmapP, bar_full(stage), 64, row, pol);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (one thread)
    if (lane == 0) {
      int stage = 0, phase = 0, a_phase = 0, cur_g = -1;
      uint32_t acc_it = 0;
      for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const int g = it % prm.n_qgroups, pb = it / prm.n_qgroups;
        const int p0 = pb * prm.ppi, p1 = min(p0 + prm.ppi, prm.np);
        const int c0 = prm.chunk_off[p0], c1 = prm.chunk_off[p1];
        const int ntile = ((c1 - c0) * EVDR_CHUNK + BN - 1) / BN;
        const int rc = min(R, prm.nqt - g * R);
        if (g != cur_g) {
          if (cur_g != -1) mma_commit(bar_a_empty);  // previous A tiles free once their MMAs retire
          mbar_wait(bar_a_full, a_phase);
          a_phase ^= 1;
          tc_fence_after();
          cur_g = g;
        }
        for (int t = 0; t < ntile; ++t) {
          mbar_wait(bar_full(stage), phase);
          tc_fence_after();
          const uint32_t b_base = sB + stage * S::B_STAGE_BYTES;
          for (int r = 0; r < rc; ++r) {
            const int buf = acc_it & 1;
            mbar_wait(bar_t_empty(buf), ((acc_it >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + buf * BN;
            const uint32_t a_base = sA + r * A_TILE_BYTES;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const uint64_t da = make_sw128_desc(a_base + (k >> 2) * A_HALF_BYTES + (k & 3) * 32);
              const uint64_t db = make_sw128_desc(b_base + (k >> 2) * S::B_HALF_BYTES + (k & 3) * 32);
              mma_bf16_ss(d_tmem, da, db, IDESC, k > 0);
            }
            mma_commit(bar_t_full(buf));
            ++acc_it;
          }
          mma_commit(bar_empty(stage));  // page tile consumed by all R query tiles
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ===================================================== epilogue warps
    const int quad = warp & 3;                           // TMEM lane quadrant this warp may read
    const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
    uint32_t acc_it = 0;
    int cur_g = -1;
    float qm[R];
    int qidx[R];
#pragma unroll
    for (int r = 0; r < R; ++r) { qm[r] = 0.f; qidx[r] = -1; }

    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
      const int g = it % prm.n_qgroups, pb = it / prm.n_qgroups;
      const int p0 = pb * prm.ppi, p1 = min(p0 + prm.ppi, prm.np);
      const int c0 = prm.chunk_off[p0], c1 = prm.chunk_off[p1];
      const int nch = c1 - c0;
      const int ntile = (nch * EVDR_CHUNK + BN - 1) / BN;
      const int rc = min(R, prm.nqt - g * R);
      if (g != cur_g) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int q = (g * R + r) * 4 + quad;
          const bool ok = (r < rc) && (q < prm.nq);
          qidx[r] = ok ? q : -1;
          qm[r] = ok ? (float)prm.qmask32[(size_t)q * EVDR_TC_LQ + lane] : 0.f;
        }
        cur_g = g;
      }
      // page-end cache: lane l holds the (item-relative) end chunk of page p0 + 32*cb + l
      int cb = 0;
      auto load_pe = [&](int b) {
        const int p = p0 + 32 * b + lane;
        return (p < p1) ? (prm.chunk_off[p + 1] - c0) : 0x7fffffff;
      };
      int pe_cache = load_pe(0);
      int p = p0;                     // current page
      int pe = __shfl_sync(0xffffffffu, pe_cache, 0);
      float runmax[R];
#pragma unroll
      for (int r = 0; r < R; ++r) runmax[r] = -INFINITY;

      for (int t = 0; t < ntile; ++t) {
        const int p_save = p;
        const int ch_tile0 = t * CH_PER_TILE;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          if (r < rc) {
            // every query tile walks the same chunks of this page tile
            if (p != p_save) {
              p = p_save;
              const int b = (p - p0) >> 5;
              if (b != cb) { cb = b; pe_cache = load_pe(b); }
              pe = __shfl_sync(0xffffffffu, pe_cache, (p - p0) & 31);
            }
            const int buf = acc_it & 1;
            mbar_wait(bar_t_full(buf), (acc_it >> 1) & 1);
            tc_fence_after();
            const uint32_t t_addr = t_lane + buf * BN;
            float rm = runmax[r];
            // page complete for this query tile: mask, reduce the 32 query tokens, store, next page
            auto finalize = [&]() {
              const float s = warp_sum(rm * qm[r]);
              if (lane == 0 && qidx[r] >= 0) prm.out[(size_t)qidx[r] * prm.out_ld + p] = s;
              rm = -INFINITY;
              ++p;
              const int b = (p - p0) >> 5;
              if (b != cb) { cb = b; pe_cache = load_pe(b); }
              pe = __shfl_sync(0xffffffffu, pe_cache, (p - p0) & 31);
            };
            if (pe >= ch_tile0 + CH_PER_TILE) {
              // ---- fast path: the whole tile belongs to the current page.  TMEM loads are
              // double-buffered in registers so that the next 32 columns are in flight while the
              // current 32 go through the FMNMX3 tree.
              float va[32], vb[32];
              tmem_ld32_issue(t_addr, va);
              tmem_wait_ld_tied(va);
#pragma unroll
              for (int grp = 0; grp < BN / 32; grp += 2) {
                tmem_ld32_issue(t_addr + (grp + 1) * 32, vb);
                rm = max_tree32(rm, va);
                tmem_wait_ld_tied(vb);
                if (grp + 2 < BN / 32) {
                  tmem_ld32_issue(t_addr + (grp + 2) * 32, va);
                } else {
                  // last TMEM read of this buffer is complete: hand it back to the MMA warp early
                  tc_fence_before();
                  __syncwarp();
                  if (lane == 0) mbar_arrive(bar_t_empty(buf));
                }
                rm = max_tree32(rm, vb);
                if (grp + 2 < BN / 32) tmem_wait_ld_tied(va);
              }
              if (pe == ch_tile0 + CH_PER_TILE) finalize();
            } else {
              // ---- boundary path: a page ends inside the tile (or the item ends): 32 columns
              // (two chunks) at a time, page ends can fall after either chunk
#pragma unroll 1
              for (int grp = 0; grp < BN / 32; ++grp) {
                const int ch = ch_tile0 + grp * 2;
                if (ch >= nch) break;
                float v[32];
                tmem_ld32_wait(t_addr + grp * 32, v);
                if (pe > ch + 1 && ch + 1 < nch) {        // both chunks belong to the current page
                  rm = max_tree32(rm, v);
                  if (pe == ch + 2) finalize();
                } else {
                  rm = max_tree16(rm, v, 0);
                  if (pe == ch + 1) finalize();
                  if (ch + 1 < nch) {
                    rm = max_tree16(rm, v, 16);
                    if (pe == ch + 2) finalize();
                  }
                }
              }
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(bar_t_empty(buf));
            }
            runmax[r] = rm;
            ++acc_it;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

static int encode_2d_bf16(evdr_handle* h, CUtensorMap* tm, const void* base, uint64_t rows,
                          uint32_t box_rows, const char* what) {
  cuuint64_t dims[2] = {(cuuint64_t)EVDR_TC_DIM, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)EVDR_TC_DIM * 2};
  cuuint32_t box[2] = {64u, box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = h->encode_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims,
                               strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(EVDR_ERR_CUDA, "cuTensorMapEncodeTiled(%s) failed with CUresult %d", what, (int)r);
  return EVDR_OK;
}

template <int R, int BN, int STAGES>
static int launch_tc(evdr_handle* h, const void* Qb, const void* tokens, int64_t n_tokens,
                     TcParams prm, cudaStream_t st) {
  using S = TcSmem<R, BN, STAGES>;
  CUtensorMap tmQ, tmP;
  int rc = encode_2d_bf16(h, &tmQ, Qb, (uint64_t)prm.nq * EVDR_TC_LQ, 128, "queries");
  if (rc != EVDR_OK) return rc;
  rc = encode_2d_bf16(h, &tmP, tokens, (uint64_t)n_tokens, BN, "page tokens");
  if (rc != EVDR_OK) return rc;
  prm.n_qgroups = (prm.nqt + R - 1) / R;
  auto kern = maxsim_tc_kernel<R, BN, STAGES>;
  if ((size_t)S::TOTAL > h->smem_optin)
    return set_error(EVDR_ERR_UNSUPPORTED, "tensor-core kernel needs %d B of shared memory", S::TOTAL);
  EVDR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
  const long long n_items = (long long)prm.n_pblocks * prm.n_qgroups;
  const int grid = (int)(n_items < h->sm_count ? n_items : h->sm_count);
  kern<<<grid, TC_THREADS, S::TOTAL, st>>>(tmQ, tmP, prm);
  EVDR_LAUNCH_CHECK("maxsim_tc_kernel");
  return EVDR_OK;
}

}  // namespace evdr

using namespace evdr;

extern "C" int evdr_maxsim_packed_fwd(evdr_handle* h, const void* Qb, const uint8_t* qmask32, int nq,
                                      const void* tokens, int64_t n_tokens,
                                      const int32_t* chunk_off, int np, int64_t total_chunks,
                                      int pages_per_item, float* out, int64_t out_ld, void* stream) {
  EVDR_REQUIRE_HANDLE(h);
  EVDR_CHECK_ARG(Qb && qmask32 && tokens && chunk_off && out, "evdr_maxsim_packed_fwd: null pointer");
  EVDR_CHECK_ARG(nq >= 1 && np >= 1 && out_ld >= np, "evdr_maxsim_packed_fwd: bad shape nq=%d np=%d ld=%lld",
                 nq, np, (long long)out_ld);
  EVDR_CHECK_ARG(total_chunks >= np && n_tokens >= total_chunks * EVDR_CHUNK,
                 "evdr_maxsim_packed_fwd: token stream (%lld) shorter than %lld chunks",
                 (long long)n_tokens, (long long)total_chunks);
  EVDR_CHECK_ARG(n_tokens < (1LL << 31), "evdr_maxsim_packed_fwd: more than 2^31 tokens per shard");
  EVDR_CHECK_ARG(((uintptr_t)Qb & 15) == 0 && ((uintptr_t)tokens & 15) == 0,
                 "evdr_maxsim_packed_fwd: operands must be 16-byte aligned for TMA");
  TcParams prm;
  prm.qmask32 = qmask32;
  prm.chunk_off = chunk_off;
  prm.out = out;
  prm.out_ld = out_ld;
  prm.nq = nq;
  prm.np = np;
  prm.nqt = (nq + 3) / 4;
  const int R = prm.nqt >= 3 ? 3 : prm.nqt;
  const int n_qgroups = (prm.nqt + R - 1) / R;
  int ppi = pages_per_item;
  if (ppi <= 0) {
    // ~16K tokens per item, tile-aligned for uniform pages; shrink while the grid is underfilled
    const double avg_tok = (double)total_chunks * EVDR_CHUNK / (double)np;
    ppi = (int)(16384.0 / avg_tok + 0.5);
    if (ppi < 1) ppi = 1;
    if (ppi >= 16) ppi = (ppi / 16) * 16;
    while (ppi > 1 && (long long)((np + ppi - 1) / ppi) * n_qgroups < 4LL * h->sm_count) ppi = (ppi + 1) / 2;
  }
  if (ppi > np) ppi = np;
  prm.ppi = ppi;
  prm.n_pblocks = (np + ppi - 1) / ppi;
  prm.n_qgroups = n_qgroups;
  prm.stream_hint = (n_qgroups == 1) ? 1 : 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (R == 1) return launch_tc<1, 256, 3>(h, Qb, tokens, n_tokens, prm, st);
  if (R == 2) return launch_tc<2, 256, 2>(h, Qb, tokens, n_tokens, prm, st);
  return launch_tc<3, 256, 2>(h, Qb, tokens, n_tokens, prm, st);
}
