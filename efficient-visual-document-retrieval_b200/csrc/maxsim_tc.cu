// maxsim_tc.cu -- K1: fused MaxSim on 5th-gen tensor cores (sm_100a).
//
//   out[q,c] = sum_n qmask[q,n] * max_{m in page c} <Q[q,n,:], P[c,m,:]>
//
// (evaluator/retrieval.py:166-213 on the packed page stream, where masked slots hold
//  duplicates of valid tokens and empty pages are zero tokens -- see pack.cu.)
//
// Structure: persistent, warp-specialised, one CTA per SM, 192 or 320 threads:
//   warp 0      TMA producer  : draws work items from a global counter (dynamic scheduler) and publishes
//                               them through a shared-memory ring; cp.async.bulk.tensor of query tiles (A,
//                               resident; double-buffered in the pair flavour) and page-token tiles (B,
//                               STAGES-deep ring) into 128B-swizzled shared memory
//   warp 1      MMA issuer    : one thread issues tcgen05.mma.kind::f16 (K=16 x 8 per pass; x 24 for the
//                               bf16x3 split), fp32 accumulators double-buffered in TMEM (2 x BN columns)
//   warps 2..   epilogue      : one group of 4 warps, or two groups (EW = 8: pass (tile, r) -> TMEM buffer r
//                               -> group r).  tcgen05.ld 32 columns at a time (double-buffered in
//                               registers); lane = query token, so the max over page tokens is a per-thread
//                               FMNMX3 tree; at a page's last chunk: * qmask, warp-shuffle sum over the
//                               32 tokens of the query, one fp32 store (and / or the fused top-k list).
//                               The similarity matrix never leaves the SM.
//
// Flavours of the same kernel (template parameters, see packed_dispatch for which runs when):
//   CG2 = false : one CTA per tile, tcgen05.mma.cta_group::1, M = 128 x N = BN.  3-12 queries, tile counts
//                 that would idle slots of a pair, the argmax (training) variants.
//   CG2 = true  : CTA pairs (cluster of 2), tcgen05.mma.cta_group::2, M = 256 x N = 256: each CTA holds its
//                 own 128 query-token rows and HALF of the page tile, so the shared-memory operand traffic
//                 per SM drops from 96 to 64 B/clk and the L2->SM page traffic halves.  The leader CTA
//                 issues the MMAs and draws the items for both (DSMEM ring).  THE DEFAULT for batched
//                 scoring (>= 13 queries): 1584-1614 TFLOP/s at 1024 queries x 100k pages.
//   (one or two queries run maxsim_scan.cu, where the page tokens are the M side.)
//
// Work item = (page block pb, query group g): the group's resident A tiles x all B tiles of the block.
// Items are ordered g-fastest and handed out dynamically, so that concurrently running CTAs read the
// same page block (L2 hit) and HBM sees every page about once (DRAM traffic 1.07x algorithmic).
#include "common.cuh"
#include "ptx.cuh"

namespace evdr {

using namespace ptx;

constexpr int A_HALF_BYTES = 128 * 64 * 2;       // one 64-wide K sub-tile (128 rows of 128 B)
#ifndef EVDR_PAIR_MIN_TILES
#define EVDR_PAIR_MIN_TILES 4                    // CTA pairs from this many query tiles (13 queries) on
#endif

struct TcParams {
  const uint8_t* qmask32;     // [nq, 32]
  const int32_t* chunk_off;   // [np + 1]
  float* out;                 // [nq, out_ld]
  long long out_ld;
  int nq, np;
  int nqt;                    // ceil(nq / 4) A tiles
  int ppi;                    // pages per item
  int n_pblocks, n_qgroups;
  int g_base, g_rem;          // query tiles per group: g_base (+1 for the first g_rem groups)
  int stream_hint;            // 1: page tiles are read once (evict-first)
  int32_t* argmax;            // AM variant: [np, nq*lq] winning packed slot per (page, query token)
  int lq;                     // AM variant: real query length (rows >= lq are padding)
  int* sched;                 // dynamic scheduler: global item counter (zeroed per launch); CG2: unused
  const __nv_bfloat16* qb;    // TA variant: packed queries [nq, 32, 128] in global memory (staged to TMEM)
  // TK variant (fused local top-k, one query group): per (query, CTA) the k best pages this CTA scored,
  // best first, padded with (-inf, -1): cand_s / cand_i [nq][gridDim.x][tk_k]; ids are page + id_offset
  float* cand_s;
  long long* cand_i;
  int tk_k;
  long long id_offset;
};

// KB: 1 = plain bf16 operands (K = 128); 3 = the bf16x3 split: both operands are stored as TWO 128-wide
// K blocks [hi | lo] (hi = bf16(x), lo = bf16(x - hi)) and the MMA issuer runs THREE K = 128 products
// into the same accumulator, Qhi.Phi + Qhi.Plo + Qlo.Phi, by pointing the descriptors at the matching
// blocks -- the fp32 dot product up to the dropped Qlo.Plo term (~2^-18 relative), the high-fidelity mode
// TA: the resident query tiles (A operand) live in TENSOR MEMORY instead of shared memory: the epilogue
// warps copy their query's 32 token rows global -> registers -> TMEM (tcgen05.st) at every item change and
// the MMAs take A from TMEM.  Shared-memory operand reads drop from 96 to 64 B/clk per SM (only the page
// tile is read), and the 64-96 KB the query tiles occupied go to a deeper page-tile ring.  TMEM holds two
// BN = 192-column accumulators (384) + R = 2 query tiles of 64 columns (128) = 512 columns.
template <int R, int BN, int STAGES, bool CG2, int KB = 1, bool SP = false, int EW = 4, bool TA = false>
struct TcCfg {
  static constexpr int B_ROWS = CG2 ? BN / 2 : BN;          // page tokens held by this CTA per stage
  static constexpr int KSUB = (KB == 3) ? 4 : 2 * KB;        // 64-wide K sub-tiles STORED per operand tile
  static constexpr int KSTEPS = (KB == 3) ? 24 : 8 * KB;      // K = 16 MMA steps per pass
  static constexpr int B_HALF_BYTES = B_ROWS * 64 * 2;       // one 64-wide K sub-tile
  static constexpr int B_STAGE_BYTES = KSUB * B_HALF_BYTES;
  static constexpr int A_TILE_BYTES = KSUB * A_HALF_BYTES;
  // NA = 2 (the CTA-pair flavour with two epilogue groups): the query tiles of the NEXT query group are loaded
  // into a second buffer while the MMAs of the current item drain, so an item change costs no idle tensor time
  static constexpr int NA = (CG2 && EW == 8 && KB == 1 && !SP) ? 2 : 1;   // (SP: the staging rows need the room)
  static constexpr int A_BYTES = TA ? 0 : NA * R * A_TILE_BYTES;
  static constexpr int A_COL = 2 * BN;                       // TA: first TMEM column of query tile 0
  static constexpr int B_BYTES = STAGES * B_STAGE_BYTES;
  static constexpr int BAR_OFF = A_BYTES + B_BYTES;
  // barriers: full[STAGES], empty[STAGES], a_full, a_empty, tmem_full[2], tmem_empty[2], tmem ptr,
  // then the scheduler ring: slot_full[NSLOT], slot_empty[NSLOT], item[NSLOT]
  static constexpr int NSLOT = 4;
  static constexpr int BAR_BYTES = 8 * (2 * STAGES + 6) + 8 + 8 * 2 * NSLOT + 4 * NSLOT + 16;   // + a_full / a_empty of buffer 1
  // short-page variant: per epilogue warp a [16 pages][32 tokens (+1 pad)] fp32 staging area where the
  // per-token maxes of the pages completed in a tile wait for their 32-token sums
  static constexpr int STG_ROW = 33;
  static constexpr int STG_BYTES = SP ? EW * (BN / EVDR_CHUNK) * STG_ROW * 4 : 0;
  static constexpr int STG_OFF = BAR_OFF + ((BAR_BYTES + 15) / 16) * 16;
  static constexpr int TOTAL = STG_OFF + STG_BYTES + 1024;  // + alignment slack
  static constexpr int TILES_PER_GROUP = CG2 ? 2 * R : R;
};

// SP ("short pages"): epilogue variant for corpora whose pages are shorter than a tile, where several
// pages end inside every tile -- page-end work is batched (see `flush`); the default variant keeps the
// leaner code that is fastest when a page spans several tiles.
// EW = number of epilogue warps: 4 (one group, TMEM buffers alternate per pass) or 8 (two groups of 4;
// requires R == 2: pass (tile t, query tile r) always uses TMEM buffer r and is drained by group r, so
// each group owns one query tile's running state and has two MMA periods per pass).
// TK ("top-k"): launches with ONE query group (<= 4 * R queries: the latency path) keep, per epilogue warp,
// the k <= 32 best (score, page) of its query in REGISTERS (lane j = j-th best; an insertion is one ballot
// and two shuffle-ups, and only happens when a page beats the current k-th) and write them once at the
// end, so that only k candidates per (query, CTA) leave the SM -- the (nq, np) score row is optional.
template <int R, int BN, int STAGES, bool CG2, bool AM = false, int KB = 1, bool SP = false, int EW = 4,
          bool TA = false, bool TK = false>
__global__ void __launch_bounds__(64 + 32 * EW, 1)
maxsim_tc_kernel(const __grid_constant__ CUtensorMap tmapQ, const __grid_constant__ CUtensorMap tmapP,
                 const TcParams prm) {
  using S = TcCfg<R, BN, STAGES, CG2, KB, SP, EW, TA>;
  static_assert(!TA || (!CG2 && !AM && KB == 1 && 2 * BN + R * 64 <= 512),
                "TMEM-resident query tiles: single-CTA bf16 scoring flavour, accumulators + tiles within 512 columns");
  constexpr int A_TILE_BYTES = S::A_TILE_BYTES;
  constexpr int KSUB = S::KSUB;
  constexpr int TMEM_NEED = 2 * BN + (TA ? R * 64 : 0);
  constexpr int TMEM_COLS = (TMEM_NEED <= 32) ? 32 : (TMEM_NEED <= 64) ? 64 : (TMEM_NEED <= 128) ? 128
                            : (TMEM_NEED <= 256) ? 256 : 512;
  static_assert(2 * BN <= 512, "two accumulator buffers must fit TMEM");
  static_assert(BN % 64 == 0 && BN >= 64 && BN <= 256, "BN");
  static_assert((BN / 32) % 2 == 0, "the drain loop handles 32-column groups in pairs");
  static_assert(EW == 4 || (EW == 8 && R == 2), "8 epilogue warps need R == 2");
  static_assert(!TK || (!CG2 && !AM && KB == 1), "fused top-k: single-CTA bf16 scoring flavours");
  constexpr uint32_t IDESC = make_idesc_bf16(CG2 ? 256 : 128, BN);
  constexpr int CH_PER_TILE = BN / EVDR_CHUNK;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = smem_base;
  const uint32_t sB = smem_base + S::A_BYTES;
  const uint32_t bars = smem_base + S::BAR_OFF;
  auto bar_full = [&](int s) { return bars + 8u * s; };
  auto bar_empty = [&](int s) { return bars + 8u * (STAGES + s); };
  constexpr int NA = S::NA;
  const uint32_t bar_a_extra = bars + 8u * (2 * STAGES + 6) + 8u + 8u * 2 * S::NSLOT + 4u * S::NSLOT;
  auto bar_a_full = [&](int b) { return b == 0 ? bars + 8u * (2 * STAGES) : bar_a_extra; };
  auto bar_a_empty = [&](int b) { return b == 0 ? bars + 8u * (2 * STAGES + 1) : bar_a_extra + 8u; };
  auto bar_t_full = [&](int b) { return bars + 8u * (2 * STAGES + 2 + b); };
  auto bar_t_empty = [&](int b) { return bars + 8u * (2 * STAGES + 4 + b); };
  const uint32_t tmem_slot = bars + 8u * (2 * STAGES + 6);
  constexpr int NSLOT = S::NSLOT;
  const uint32_t sched_base = tmem_slot + 8u;
  auto bar_s_full = [&](int s) { return sched_base + 8u * s; };
  auto bar_s_empty = [&](int s) { return sched_base + 8u * (NSLOT + s); };
  volatile int* item_ring =
      reinterpret_cast<volatile int*>(smem_raw + (sched_base + 8u * 2 * NSLOT - smem_u32(smem_raw)));
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = CG2 ? (int)cluster_ctarank() : 0;            // 0 = leader of the CTA pair
  const int n_units = CG2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmapQ);
    prefetch_tmap(&tmapP);
    for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), 1); }
    for (int b = 0; b < NA; ++b) {
      mbar_init(bar_a_full(b), TA ? EW : 1);    // TA: every epilogue warp arrives after staging its rows
      mbar_init(bar_a_empty(b), 1);
    }
    // tmem_empty lives in the leader: 4 epilogue warps of each CTA of the pair arrive on it
    for (int b = 0; b < 2; ++b) { mbar_init(bar_t_full(b), 1); mbar_init(bar_t_empty(b), CG2 ? 8 : 4); }
    // scheduler ring: the producer publishes item indices, the MMA thread + 4 epilogue warps consume
    // (pair mode: the leader draws for both CTAs; its slot_empty barriers also collect the peer's producer and
    //  epilogue warps, the peer's slot_full barriers are armed by the leader's remote arrive)
    for (int s = 0; s < NSLOT; ++s) { mbar_init(bar_s_full(s), 1); mbar_init(bar_s_empty(s), (CG2 ? 2 : 1) * (1 + EW)); }
    fence_mbar_init();
  }
  if (warp == 1) {
    if constexpr (CG2) tmem_alloc_cg2<TMEM_COLS>(tmem_slot); else tmem_alloc<TMEM_COLS>(tmem_slot);
  }
  tc_fence_before();
  if constexpr (CG2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int n_items = prm.n_pblocks * prm.n_qgroups;
  // query tiles are dealt to the groups as evenly as possible (sizes differ by at most one), so
  // that the static round-robin item schedule stays balanced
  auto group_start = [&](int g) { return g * prm.g_base + min(g, prm.g_rem); };
  auto group_size = [&](int g) { return prm.g_base + (g < prm.g_rem ? 1 : 0); };
  // passes (query tiles of this CTA) in group g
  auto passes_in_group = [&](int g) { return CG2 ? (group_size(g) + 1) >> 1 : group_size(g); };

  if (warp == 0) {
    // ===================================================== TMA producer (each CTA loads its own
    // query tiles and its half of the page tile; in pair mode all bytes are counted on the
    // LEADER's barriers, which the MMA issuer waits on)
    if (lane == 0) {
      const uint64_t pol = prm.stream_hint ? policy_evict_first() : policy_evict_last();
      int stage = 0, phase = 0, a_loads = 0, cur_g = -1;

      // dynamic tile scheduler: the producer draws the next item from a global counter and publishes it to
      // the other roles through a small shared-memory ring, so all CTAs work on neighbouring items at any
      // time and re-reads of a page block hit L2 however long the run is.  In pair mode the leader CTA
      // draws and also writes the item into the peer's ring through distributed shared memory.
      int sslot = 0, sphase = 0;
      const uint32_t ring_u32 = sched_base + 8u * 2 * NSLOT;       // shared-memory address of item_ring[0]
      auto next_item = [&](int) -> int {
        int v;
        if (!CG2 || rank == 0) {
          mbar_wait(bar_s_empty(sslot), sphase ^ 1);
          v = atomicAdd(prm.sched, 1);
          // every drawing CTA draws exactly one terminal value (>= n_items); the one drawing the last of them
          // hands the counter back zeroed, so the next launch on this slot needs no memset
          if (v == n_items + n_units - 1) atomicExch(prm.sched, 0);
          item_ring[sslot] = v;
          mbar_arrive(bar_s_full(sslot));
          if constexpr (CG2) {   // the same item for the peer CTA: remote store, then a releasing remote arrive
            st_shared_cluster_u32(mapa_u32(ring_u32 + 4u * sslot, 1), (uint32_t)v);
            mbar_arrive_remote(mapa_u32(bar_s_full(sslot), 1));
          }
        } else {                 // peer CTA of a pair: consume what the leader published
          mbar_wait_cluster(bar_s_full(sslot), sphase);
          v = item_ring[sslot];
          mbar_arrive_remote(mapa_u32(bar_s_empty(sslot), 0));
        }
        if (++sslot == NSLOT) { sslot = 0; sphase ^= 1; }
        return v;
      };
      for (int it = next_item(-1); it < n_items; it = next_item(it)) {
        const int g = it % prm.n_qgroups, pb = it / prm.n_qgroups;
        const int p0 = pb * prm.ppi, p1 = min(p0 + prm.ppi, prm.np);
        const int c0 = prm.chunk_off[p0], c1 = prm.chunk_off[p1];
        const int ntile = ((c1 - c0) * EVDR_CHUNK + BN - 1) / BN;
        const int rc = passes_in_group(g);
        if (!TA && g != cur_g) {
          const int ab = a_loads % NA;
          if (a_loads >= NA) mbar_wait(bar_a_empty(ab), ((a_loads / NA) - 1) & 1);
          const uint32_t a_full_tx = CG2 ? mapa_u32(bar_a_full(ab), 0) : bar_a_full(ab);
          const uint32_t a_dst = sA + ab * (R * A_TILE_BYTES);
          // out-of-range query tiles are zero-filled by TMA and still count their bytes
          if (rank == 0) mbar_expect_tx(bar_a_full(ab), (CG2 ? 2 : 1) * rc * A_TILE_BYTES);
          for (int r = 0; r < rc; ++r) {
            const int tile = group_start(g) + (CG2 ? 2 * r + rank : r);
#pragma unroll
            for (int kk = 0; kk < KSUB; ++kk) {
              if constexpr (CG2) tma_load_2d_cg2(a_dst + r * A_TILE_BYTES + kk * A_HALF_BYTES, &tmapQ, a_full_tx, kk * 64, tile * 128);
              else tma_load_2d(a_dst + r * A_TILE_BYTES + kk * A_HALF_BYTES, &tmapQ, a_full_tx, kk * 64, tile * 128);
            }
          }
          ++a_loads;
          cur_g = g;
        }
        for (int t = 0; t < ntile; ++t) {
          mbar_wait(bar_empty(stage), phase ^ 1);
          if (rank == 0) mbar_expect_tx(bar_full(stage), (CG2 ? 2 : 1) * S::B_STAGE_BYTES);
          const int row = c0 * EVDR_CHUNK + t * BN + rank * S::B_ROWS;
          const uint32_t dst = sB + stage * S::B_STAGE_BYTES;
          if constexpr (CG2) {
            const uint32_t full_tx = mapa_u32(bar_full(stage), 0);
#pragma unroll
            for (int kk = 0; kk < KSUB; ++kk)
              tma_load_2d_cg2_hint(dst + kk * S::B_HALF_BYTES, &tmapP, full_tx, kk * 64, row, pol);
          } else {
#pragma unroll
            for (int kk = 0; kk < KSUB; ++kk)
              tma_load_2d_hint(dst + kk * S::B_HALF_BYTES, &tmapP, bar_full(stage), kk * 64, row, pol);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (one thread, leader CTA)
    if (lane == 0 && rank == 0) {
      int stage = 0, phase = 0, a_uses = 0, a_cur = 0, cur_g = -1;
      uint32_t acc_it = 0;
      auto commit = [&](uint32_t bar) {
        if constexpr (CG2) mma_commit_cg2_mc(bar, (uint16_t)0x3); else mma_commit(bar);
      };
      int sslot = 0, sphase = 0;
      uint32_t use0 = 0, use1 = 0;
      auto next_item = [&](int) -> int {               // leader CTA only: its ring is filled locally
        mbar_wait(bar_s_full(sslot), sphase);
        const int v = item_ring[sslot];
        mbar_arrive(bar_s_empty(sslot));
        if (++sslot == NSLOT) { sslot = 0; sphase ^= 1; }
        return v;
      };
      for (int it = next_item(-1); it < n_items; it = next_item(it)) {
        const int g = it % prm.n_qgroups, pb = it / prm.n_qgroups;
        const int p0 = pb * prm.ppi, p1 = min(p0 + prm.ppi, prm.np);
        const int c0 = prm.chunk_off[p0], c1 = prm.chunk_off[p1];
        const int ntile = ((c1 - c0) * EVDR_CHUNK + BN - 1) / BN;
        const int rc = passes_in_group(g);
        if (g != cur_g) {
          if (!TA && cur_g != -1) commit(bar_a_empty(a_cur));  // previous A tiles free once their MMAs retire
          a_cur = a_uses % NA;
          mbar_wait(bar_a_full(a_cur), (a_uses / NA) & 1);
          ++a_uses;
          tc_fence_after();
          cur_g = g;
        }
        for (int t = 0; t < ntile; ++t) {
          mbar_wait(bar_full(stage), phase);
          tc_fence_after();
          const uint32_t b_base = sB + stage * S::B_STAGE_BYTES;
          for (int r = 0; r < rc; ++r) {
            const int buf = (EW == 8) ? r : (int)(acc_it & 1);
            const uint32_t nuse = buf ? use1 : use0;          // how often this buffer was filled before
            mbar_wait(bar_t_empty(buf), (nuse & 1) ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + buf * BN;
            const uint32_t a_base = sA + a_cur * (R * A_TILE_BYTES) + r * A_TILE_BYTES;
#pragma unroll
            for (int k = 0; k < S::KSTEPS; ++k) {
              // bf16x3: steps 0-7 hi.hi, 8-15 Qhi.Plo, 16-23 Qlo.Phi (largest product first)
              const int ka = (KB == 3) ? ((k >> 3) == 2 ? 2 : 0) + ((k & 7) >> 2) : (k >> 2);
              const int kb = (KB == 3) ? ((k >> 3) == 1 ? 2 : 0) + ((k & 7) >> 2) : (k >> 2);
              const uint64_t da = make_sw128_desc(a_base + ka * A_HALF_BYTES + (k & 3) * 32);
              const uint64_t db = make_sw128_desc(b_base + kb * S::B_HALF_BYTES + (k & 3) * 32);
              if constexpr (CG2) mma_bf16_ss_cg2(d_tmem, da, db, IDESC, k > 0);
              else if constexpr (TA) mma_bf16_ts(d_tmem, tmem_base + S::A_COL + r * 64 + k * 8, db, IDESC, k > 0);
              else mma_bf16_ss(d_tmem, da, db, IDESC, k > 0);
            }
            commit(bar_t_full(buf));
            if (buf) ++use1; else ++use0;
            ++acc_it;
          }
          commit(bar_empty(stage));  // page tile consumed by all resident query tiles
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ===================================================== epilogue warps
    const int quad = warp & 3;                           // TMEM lane quadrant this warp may read
    const int egrp = (warp - 2) >> 2;                    // EW == 8: epilogue group 0 / 1 <-> query tile r
    float* stg = reinterpret_cast<float*>(smem_raw + (smem_base - smem_u32(smem_raw)) + S::STG_OFF) +
                 (warp - 2) * (BN / EVDR_CHUNK) * S::STG_ROW;   // SP: this warp's staging rows
    const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
    uint32_t acc_it = 0;
    int cur_g = -1;
    float qm[R];
    int qidx[R];
#pragma unroll
    for (int r = 0; r < R; ++r) { qm[r] = 0.f; qidx[r] = -1; }
    // TK: this warp's register list per resident query tile (entry `lane`) and the score to beat
    float tk_s[TK ? R : 1], tk_thr[TK ? R : 1];
    int tk_p[TK ? R : 1];
#pragma unroll
    for (int r = 0; r < (TK ? R : 1); ++r) { tk_s[r] = -INFINITY; tk_thr[r] = -INFINITY; tk_p[r] = -1; }
    // pages arrive in ascending order within a CTA (items are drawn from an increasing counter), so an equal
    // score never displaces an entry: strict '>' keeps (score desc, page asc).  `s` is warp-uniform.
    auto tk_insert = [&](float& ls, int& lp, float& thr, float s, int page) {
      if (s > thr) {
        const int pos = __popc(__ballot_sync(0xffffffffu, ls >= s));
        const float us = __shfl_up_sync(0xffffffffu, ls, 1);
        const int up = __shfl_up_sync(0xffffffffu, lp, 1);
        if (lane == pos) { ls = s; lp = page; }
        else if (lane > pos) { ls = us; lp = up; }
        thr = __shfl_sync(0xffffffffu, ls, prm.tk_k - 1);
      }
    };
    const uint32_t t_empty0 = CG2 ? mapa_u32(bar_t_empty(0), 0) : bar_t_empty(0);
    const uint32_t t_empty1 = CG2 ? mapa_u32(bar_t_empty(1), 0) : bar_t_empty(1);
    auto release_tmem = [&](int buf) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        const uint32_t a = buf ? t_empty1 : t_empty0;
        if constexpr (CG2) mbar_arrive_remote_relaxed(a); else mbar_arrive(a);
      }
    };

    int sslot = 0, sphase = 0;
    const uint32_t s_empty_leader = CG2 ? mapa_u32(bar_s_empty(0), 0) : bar_s_empty(0);
    auto next_item = [&](int) -> int {
      if (CG2 && rank != 0) mbar_wait_cluster(bar_s_full(sslot), sphase);   // written by the leader CTA
      else mbar_wait(bar_s_full(sslot), sphase);
      const int v = item_ring[sslot];
      __syncwarp();                                  // every lane has read the slot
      if (lane == 0) {
        if (CG2 && rank != 0) mbar_arrive_remote(s_empty_leader + 8u * sslot);
        else mbar_arrive(bar_s_empty(sslot));
      }
      if (++sslot == NSLOT) { sslot = 0; sphase ^= 1; }
      return v;
    };
    for (int it = next_item(-1); it < n_items; it = next_item(it)) {
      const int g = it % prm.n_qgroups, pb = it / prm.n_qgroups;
      const int p0 = pb * prm.ppi, p1 = min(p0 + prm.ppi, prm.np);
      const int c0 = prm.chunk_off[p0], c1 = prm.chunk_off[p1];
      const int nch = c1 - c0;
      const int ntile = (nch * EVDR_CHUNK + BN - 1) / BN;
      const int rc = passes_in_group(g);
      if (g != cur_g) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int tile = group_start(g) + (CG2 ? 2 * r + rank : r);
          const int q = tile * 4 + quad;
          // in pair mode the peer's last tile of an odd-sized group belongs to the next group
          const bool ok = (r < rc) && (tile < group_start(g) + group_size(g)) && (q < prm.nq);
          qidx[r] = ok ? q : -1;
          qm[r] = ok ? (float)prm.qmask32[(size_t)q * EVDR_TC_LQ + lane] : 0.f;
          if constexpr (TA) {
            // stage this warp's query (32 token rows = TMEM lanes of its quadrant) as query tile r.  The warp
            // has drained the last pass that read tile r, so the columns are free; the other tile may still
            // be in use by a running MMA and is not touched.
            if (r < rc && (EW == 4 || r == egrp)) {
              uint32_t a[64];
              if (ok) {
                const uint4* src = reinterpret_cast<const uint4*>(prm.qb + ((size_t)q * EVDR_TC_LQ + lane) * EVDR_TC_DIM);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  const uint4 v = __ldg(src + i);
                  a[4 * i] = v.x; a[4 * i + 1] = v.y; a[4 * i + 2] = v.z; a[4 * i + 3] = v.w;
                }
              } else {
#pragma unroll
                for (int i = 0; i < 64; ++i) a[i] = 0u;
              }
              tmem_st64(t_lane + S::A_COL + r * 64, a);
            }
          }
        }
        if constexpr (TA) {
          tmem_wait_st();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_a_full(0));
        }
        cur_g = g;
      }
      // page-end cache: lane l holds the (item-relative) end chunk of page p0 + 32*cb + l
      // (the NEXT batch of 32 is loaded as soon as a batch becomes current, so the global-load latency
      //  is hidden behind 32 pages of work)
      int cb = 0;
      auto load_pe = [&](int b) {
        const int pp = p0 + 32 * b + lane;
        return (pp < p1) ? (prm.chunk_off[pp + 1] - c0) : 0x7fffffff;
      };
      int pe_cache = load_pe(0);
      int pe_next = SP ? load_pe(1) : 0;
      auto set_batch = [&](int b) {
        if constexpr (SP) {
          if (b == cb + 1) { pe_cache = pe_next; }
          else { pe_cache = load_pe(b); }
          pe_next = load_pe(b + 1);
        } else {
          pe_cache = load_pe(b);
        }
        cb = b;
      };
      int p = p0;                     // current page
      int pe = __shfl_sync(0xffffffffu, pe_cache, 0);
      float runmax[R];
      int best_chunk[R];                 // AM: item-relative chunk holding the running max ...
      int best_pos[R];                   // ... and the winner's position inside that chunk
#pragma unroll
      for (int r = 0; r < R; ++r) { runmax[r] = -INFINITY; best_chunk[r] = 0; best_pos[r] = 0; }
      int pstart = 0;                    // AM: item-relative first chunk of the current page

      for (int t = 0; t < ntile; ++t) {
        const int p_save = p;
        const int pstart_save = pstart;
        const int ch_tile0 = t * CH_PER_TILE;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          if (r < rc && (EW == 4 || r == egrp)) {
            // every query tile walks the same chunks of this page tile
            if (p != p_save) {
              p = p_save;
              pstart = pstart_save;
              const int b = (p - p0) >> 5;
              if (b != cb) set_batch(b);
              pe = __shfl_sync(0xffffffffu, pe_cache, (p - p0) & 31);
            }
            // EW == 4: buffers alternate per pass; EW == 8: this group always drains buffer r
            const int buf = (EW == 8) ? r : (int)(acc_it & 1);
            mbar_wait(bar_t_full(buf), (EW == 8) ? (acc_it & 1) : ((acc_it >> 1) & 1));
            tc_fence_after();
            const uint32_t t_addr = t_lane + buf * BN;
            float rm = runmax[r];
            int bc = best_chunk[r], bp = best_pos[r];
            // page complete for this query tile: mask, reduce the 32 query tokens, store, next page
            // Completed pages are queued (up to 4 per flush): their 32-token sums are reduced with
            // interleaved shuffles and written by 4 lanes in one store, instead of one serial
            // shuffle chain + store per page (short compressed pages end 4 or more times per tile).
            float f0 = 0.f, f1 = 0.f, f2 = 0.f, f3 = 0.f;
            int nfin = 0, fin_page0 = 0;
            auto flush = [&]() {
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) {
                f0 += __shfl_xor_sync(0xffffffffu, f0, o);
                f1 += __shfl_xor_sync(0xffffffffu, f1, o);
                f2 += __shfl_xor_sync(0xffffffffu, f2, o);
                f3 += __shfl_xor_sync(0xffffffffu, f3, o);
              }
              const float v = lane == 0 ? f0 : lane == 1 ? f1 : lane == 2 ? f2 : f3;
              if (lane < nfin && qidx[r] >= 0)
                prm.out[(size_t)qidx[r] * prm.out_ld + fin_page0 + lane] = v;
              nfin = 0;
            };
            auto finalize = [&]() {
              if constexpr (AM) {
                // winning packed slot of this (query token, page): chunk within the page * 16 + position
                const int slot = (bc - pstart) * EVDR_CHUNK + bp;
                if (qidx[r] >= 0 && lane < prm.lq)
                  prm.argmax[(size_t)p * ((size_t)prm.nq * prm.lq) + (size_t)qidx[r] * prm.lq + lane] = slot;
                if (prm.out != nullptr) {                 // scores too (gradient calls: search + score in one pass)
                  const float sc = warp_sum(rm * qm[r]);
                  if (lane == 0 && qidx[r] >= 0) prm.out[(size_t)qidx[r] * prm.out_ld + p] = sc;
                }
                pstart = pe;
                bc = pe;
              } else {
                const float s = warp_sum(rm * qm[r]);        // every lane holds the sum
                if constexpr (TK) {
                  if (prm.out != nullptr && lane == 0 && qidx[r] >= 0) prm.out[(size_t)qidx[r] * prm.out_ld + p] = s;
                  tk_insert(tk_s[r], tk_p[r], tk_thr[r], s, p);
                } else {
                  if (lane == 0 && qidx[r] >= 0) prm.out[(size_t)qidx[r] * prm.out_ld + p] = s;
                }
              }
              rm = -INFINITY;
              ++p;
              const int b = (p - p0) >> 5;
              if (b != cb) set_batch(b);
              pe = __shfl_sync(0xffffffffu, pe_cache, (p - p0) & 31);
            };
            // ---- drain: per-chunk (16-column) maxes of this tile into registers.  TMEM loads are
            // double-buffered in registers: the next 32 columns are in flight while the current 32
            // go through the FMNMX3 trees.  The buffer goes back to the MMA warp as soon as the
            // last load has landed, before any page logic.
            float h[CH_PER_TILE];
            int hp[AM ? CH_PER_TILE : 1];   // AM: first position of the chunk max inside its chunk
            {
              float va[32], vb[32];
              tmem_ld32_issue(t_addr, va);
              tmem_wait_ld_tied(va);
#pragma unroll
              for (int grp = 0; grp < BN / 32; grp += 2) {
                tmem_ld32_issue(t_addr + (grp + 1) * 32, vb);
                h[2 * grp] = max16(va, 0);
                h[2 * grp + 1] = max16(va, 16);
                if constexpr (AM) {
                  hp[2 * grp] = first_pos16(va, 0, h[2 * grp]);
                  hp[2 * grp + 1] = first_pos16(va, 16, h[2 * grp + 1]);
                }
                tmem_wait_ld_tied(vb);
                if (grp + 2 < BN / 32) tmem_ld32_issue(t_addr + (grp + 2) * 32, va);
                else release_tmem(buf);
                h[2 * grp + 2] = max16(vb, 0);
                h[2 * grp + 3] = max16(vb, 16);
                if constexpr (AM) {
                  hp[2 * grp + 2] = first_pos16(vb, 0, h[2 * grp + 2]);
                  hp[2 * grp + 3] = first_pos16(vb, 16, h[2 * grp + 3]);
                }
                if (grp + 2 < BN / 32) tmem_wait_ld_tied(va);
              }
            }
            // ---- fold the chunk maxes into the running per-page max; a page can end after any chunk
            if constexpr (SP && !AM) {
              // short pages: which of this tile's chunks close a page, as a bit mask built with one
              // redux.or over the lane-cached page ends of the current and the next batch of 32 pages
              // (pages already finished end at or before the tile start, pages past the item's end are
              // INT_MAX: neither sets a bit); then ONE pass over the 16 chunk maxes.
              auto end_bit = [&](int pev) {
                const int pos = pev - ch_tile0 - 1;
                return (pos >= 0 && pos < CH_PER_TILE) ? (1u << pos) : 0u;
              };
              const unsigned endmask = __reduce_or_sync(0xffffffffu, end_bit(pe_cache) | end_bit(pe_next));
              const float w = qm[r];
              int k = 0;                               // pages completed in this tile so far
#pragma unroll
              for (int j = 0; j < CH_PER_TILE; ++j) {
                rm = fmaxf(rm, h[j]);
                if (endmask & (1u << j)) {             // warp-uniform: chunk j closes page p + k
                  stg[k * S::STG_ROW + lane] = rm * w;  // per-token max of the finished page
                  ++k;
                  rm = -INFINITY;
                }
              }
              if (k > 0) {
                // 32-token sums of the k finished pages: lane l adds up row l (conflict-free thanks to
                // the 33-float row pitch), then k lanes store k consecutive scores at once
                __syncwarp();
                float sacc = 0.f;
                if (lane < k) {
#pragma unroll
                  for (int t = 0; t < 32; ++t) sacc += stg[lane * S::STG_ROW + t];
                  if ((!TK || prm.out != nullptr) && qidx[r] >= 0) prm.out[(size_t)qidx[r] * prm.out_ld + p + lane] = sacc;
                }
                if constexpr (TK) {                   // the k finished pages, in page order, into the register list
                  for (int i = 0; i < k; ++i)
                    tk_insert(tk_s[r], tk_p[r], tk_thr[r], __shfl_sync(0xffffffffu, sacc, i), p + i);
                }
                __syncwarp();
                p += k;
              }
              const int b = (p - p0) >> 5;
              if (b != cb) set_batch(b);
              pe = __shfl_sync(0xffffffffu, pe_cache, (p - p0) & 31);
            } else {
            const bool whole_tile = !AM && (pe >= ch_tile0 + CH_PER_TILE);
            if (whole_tile) {
              float m = rm;
#pragma unroll
              for (int j = 0; j + 1 < CH_PER_TILE; j += 2) m = max3(m, h[j], h[j + 1]);
              rm = m;
              if (pe == ch_tile0 + CH_PER_TILE) finalize();
            } else {
              // compact (not unrolled) segment loop: ONE copy of finalize per resident tile keeps the
              // kernel inside the instruction cache.  Segment [lo, hi) of this tile's chunks belongs
              // to the current page; membership is a warp-uniform predicate per chunk register.
              int lo = 0;
#pragma unroll 1
              while (lo < CH_PER_TILE) {
                const int hi = min(pe - ch_tile0, CH_PER_TILE);   // pe is INT_MAX past the item's end
                if constexpr (AM) {
#pragma unroll
                  for (int j = 0; j < CH_PER_TILE; ++j) {
                    // strict '>' in ascending chunk order keeps the first chunk attaining the max
                    const bool take = (j >= lo && j < hi) && (h[j] > rm);
                    rm = take ? h[j] : rm;
                    bc = take ? ch_tile0 + j : bc;
                    bp = take ? hp[j] : bp;
                  }
                } else {
#pragma unroll
                  for (int j = 0; j < CH_PER_TILE; ++j)
                    rm = (j >= lo && j < hi) ? fmaxf(rm, h[j]) : rm;
                }
                if (pe - ch_tile0 <= CH_PER_TILE) finalize();
                lo = hi;
              }
            }
            }
            runmax[r] = rm;
            best_chunk[r] = bc;
            best_pos[r] = bp;
            ++acc_it;
          }
        }
      }
    }
    if constexpr (TK) {
      // one query group: query tile r of this CTA is tile r of the launch whatever items it drew (a CTA
      // that drew none writes pure padding, so every (query, CTA) slot of the candidate buffer is defined)
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int q = r * 4 + quad;
        if ((EW == 4 || r == egrp) && q < prm.nq && lane < prm.tk_k) {
          const size_t o = ((size_t)q * gridDim.x + blockIdx.x) * prm.tk_k + lane;
          prm.cand_s[o] = tk_s[r];
          prm.cand_i[o] = tk_p[r] >= 0 ? (long long)tk_p[r] + prm.id_offset : -1LL;
        }
      }
    }
  }

  // teardown: in pair mode the peer's shared memory / TMEM must stay alive until both CTAs are done
  // (multicast commits and remote arrives target it)
  tc_fence_before();
  if constexpr (CG2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    __syncwarp();
    if constexpr (CG2) tmem_dealloc_cg2<TMEM_COLS>(tmem_base); else tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

static int encode_2d_bf16(evdr_handle* h, CUtensorMap* tm, const void* base, uint64_t rows,
                          uint32_t box_rows, int kb, const char* what) {
  cuuint64_t dims[2] = {(cuuint64_t)EVDR_TC_DIM * kb, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)EVDR_TC_DIM * 2 * kb};
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

template <int R, int BN, int STAGES, bool CG2, bool AM = false, int KB = 1, bool SP = false, int EW = 4,
          bool TA = false, bool TK = false>
static int launch_tc(evdr_handle* h, const void* Qb, const void* tokens, int64_t n_tokens,
                     TcParams prm, cudaStream_t st) {
  using S = TcCfg<R, BN, STAGES, CG2, KB, SP, EW, TA>;
  prm.qb = static_cast<const __nv_bfloat16*>(Qb);
  CUtensorMap tmQ, tmP;
  constexpr int STORED = (KB == 3) ? 2 : KB;                  // 128-wide K blocks per stored row
  int rc = encode_2d_bf16(h, &tmQ, Qb, (uint64_t)prm.nq * EVDR_TC_LQ, 128, STORED, "queries");
  if (rc != EVDR_OK) return rc;
  rc = encode_2d_bf16(h, &tmP, tokens, (uint64_t)n_tokens, S::B_ROWS, STORED, "page tokens");
  if (rc != EVDR_OK) return rc;
  auto kern = maxsim_tc_kernel<R, BN, STAGES, CG2, AM, KB, SP, EW, TA, TK>;
  if ((size_t)S::TOTAL > h->smem_optin)
    return set_error(EVDR_ERR_UNSUPPORTED, "tensor-core kernel needs %d B of shared memory", S::TOTAL);
  EVDR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
  const long long n_items = (long long)prm.n_pblocks * prm.n_qgroups;
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  if (CG2) {
    const int pairs = h->sm_count / 2;
    const int units = (int)(n_items < pairs ? n_items : pairs);
    cfg.gridDim = dim3(2 * units);
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  } else {
    cfg.gridDim = dim3((unsigned)(n_items < h->sm_count ? n_items : h->sm_count));
  }
  cfg.blockDim = dim3(64 + 32 * EW);
  cfg.dynamicSmemBytes = S::TOTAL;
  cfg.stream = st;
  prm.sched = nullptr;
  {
    // this launch's item counter: next slot of the handle's ring.  Slots start zeroed (evdr_create) and the
    // kernel hands its slot back zeroed (the CTA that draws the last terminal value resets it)
    prm.sched = h->sched_ring + (h->sched_next.fetch_add(1, std::memory_order_relaxed) % EVDR_SCHED_RING);
  }
  note_kernel("maxsim_tc_kernel<R=%d,BN=%d,STAGES=%d,CG2=%d,AM=%d,KB=%d,SP=%d,EW=%d,TA=%d,TK=%d> grid=%u", R, BN, STAGES,
              (int)CG2, (int)AM, KB, (int)SP, EW, (int)TA, (int)TK, cfg.gridDim.x);
  cudaError_t le = cudaLaunchKernelEx(&cfg, kern, tmQ, tmP, prm);
  if (le != cudaSuccess)
    return set_error(EVDR_ERR_CUDA, "launch of maxsim_tc_kernel failed: %s", cudaGetErrorString(le));
  return EVDR_OK;
}

}  // namespace evdr

using namespace evdr;

// Can this device / context co-schedule clusters of two CTAs of the pair flavour?  Probed once per handle with
// cudaOccupancyMaxActiveClusters; when it cannot, the dispatch silently stays on the single-CTA flavours.
static bool pairs_schedulable(evdr_handle* h) {
  if (h->pair_state == 0) {
    using S = TcCfg<2, 256, 3, true, 1, false, 8, false>;
    auto kern = maxsim_tc_kernel<2, 256, 3, true, false, 1, false, 8>;
    int n = 0;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.gridDim = dim3(2 * (h->sm_count / 2));
    cfg.blockDim = dim3(64 + 32 * 8);
    cfg.dynamicSmemBytes = S::TOTAL;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
    if (e != cudaSuccess) (void)cudaGetLastError();
    h->pair_state = (e == cudaSuccess && n >= 1) ? 1 : -1;
  }
  return h->pair_state > 0;
}

// fused top-k request of a dispatch (evdr::scan_topk): candidate buffers in, CTA count out
struct TkRequest {
  float* cand_s;
  long long* cand_i;
  int k;
  long long id_offset;
  int n_cta;
};

static int packed_dispatch(evdr_handle* h, const void* Qb, const uint8_t* qmask32, int nq,
                           const void* tokens, int64_t n_tokens, const int32_t* chunk_off, int np,
                           int64_t total_chunks, int pages_per_item, float* out, int64_t out_ld,
                           int32_t* argmax, int lq, int k_blocks, void* stream, TkRequest* tk = nullptr) {
  if (k_blocks != 1 && k_blocks != 3)
    return set_error(EVDR_ERR_INVALID, "evdr_maxsim_packed: k_blocks=%d (1 = bf16, 3 = bf16x3 split)", k_blocks);
  EVDR_REQUIRE_HANDLE(h);
  EVDR_CHECK_ARG(Qb && qmask32 && tokens && chunk_off && (out || argmax || tk), "evdr_maxsim_packed: null pointer");
  EVDR_CHECK_ARG(nq >= 1 && np >= 1 && (argmax || tk || out_ld >= np), "evdr_maxsim_packed: bad shape nq=%d np=%d ld=%lld",
                 nq, np, (long long)out_ld);
  EVDR_CHECK_ARG(total_chunks >= np && n_tokens >= total_chunks * EVDR_CHUNK,
                 "evdr_maxsim_packed_fwd: token stream (%lld) shorter than %lld chunks",
                 (long long)n_tokens, (long long)total_chunks);
  EVDR_CHECK_ARG(n_tokens < (1LL << 31), "evdr_maxsim_packed_fwd: more than 2^31 tokens per shard");
  EVDR_CHECK_ARG(((uintptr_t)Qb & 15) == 0 && ((uintptr_t)tokens & 15) == 0,
                 "evdr_maxsim_packed_fwd: operands must be 16-byte aligned for TMA");
  TcParams prm = {};
  prm.qmask32 = qmask32;
  prm.chunk_off = chunk_off;
  prm.out = out;
  prm.out_ld = out_ld;
  prm.argmax = argmax;
  prm.lq = lq;
  prm.nq = nq;
  prm.np = np;
  prm.nqt = (nq + 3) / 4;
  // kernel flavour (EVDR_TC_MODE=pair selects the cta_group::2 kernels, for A/B runs)
  const char* mode = getenv("EVDR_TC_MODE");
  if (mode && !*mode) mode = nullptr;
  // one or two queries: the operand-swapped streaming scan (maxsim_scan.cu: page tokens on the M side, 4x / 2x
  // fewer tensor cycles per byte -- the scan is power-bound as well as HBM-bound).  EVDR_TC_MODE=scan forces
  // it, any other mode keeps the query-major kernels below.
  // By page length (profiles/r2_scan_page_lengths.log, r2_short_pages_one_tile.log; GB/s of page bytes): the
  // operand-swapped kernel wins from ~160 (one query) / ~400 (two queries) tokens per page on (7.3-7.5 vs 6.8
  // TB/s); shorter pages take the query-major kernel with the short-page epilogue below (6.4-6.5 TB/s down to
  // 16-token pages, where per-page work would otherwise dominate: 1.8 TB/s with the long-page epilogue).
  const double scan_avg_tok = (double)total_chunks * EVDR_CHUNK / (double)np;
  const bool scan_len_ok = scan_avg_tok >= (nq == 1 ? 160.0 : 400.0);
  if (nq <= 2 && !argmax && k_blocks == 1 && ((!mode && scan_len_ok) || (mode && strcmp(mode, "scan") == 0))) {
    if (tk && tk->k > 32) return EVDR_ERR_UNSUPPORTED;
    return scan_swapped(h, Qb, qmask32, nq, tokens, n_tokens, chunk_off, np, total_chunks, pages_per_item, out,
                        out_ld, tk ? tk->k : 0, tk ? tk->id_offset : 0, tk ? tk->cand_s : nullptr,
                        tk ? tk->cand_i : nullptr, tk ? &tk->n_cta : nullptr, static_cast<cudaStream_t>(stream));
  }
  // CTA pairs (cta_group::2, M = 256: each CTA holds its own 128 query-token rows and HALF of the page tile,
  // so per SM the shared-memory operand reads drop from 96 to 64 B/clk and the L2 -> SM page traffic halves)
  // with two epilogue groups per CTA and the dynamic scheduler are the default from 4 query tiles on long
  // pages: 1587 vs 1508 TFLOP/s at 1024 queries x 100k pages, 1405 vs 1316 at 32, 1311 vs 1191 at 16
  // (profiles/r2_pair_vs_single_cta.log).  "pair" = the one-epilogue-group pair flavours (R = 1..4 per CTA),
  // "1cta" / "ew8" / "sp8" / "ta" = the single-CTA flavours.
  const double avg_chunks_all = (double)total_chunks / (double)np;
  // A pair holds 4 query tiles per group; a group of 3 leaves one of the four 128-row slots idle.  The pair
  // flavour is worth ~5 % (profiles/r2_pair_sweep.log: 100 queries = 25 tiles = 3 idle slots of 28 lose),
  // so it is taken when at most 4 % of the slots would idle.
  const int pair_groups = (prm.nqt + 3) / 4;
  const int idle_slots = 4 * pair_groups - prm.nqt;
  const bool pair8 = (mode && strcmp(mode, "pair8") == 0) ||
                     (!mode && prm.nqt >= EVDR_PAIR_MIN_TILES && 25 * idle_slots <= 4 * pair_groups &&
                      pairs_schedulable(h));
  // pages shorter than two tiles on average: the pair flavour with the short-page (end-mask + staged-sum)
  // epilogue -- 1560 vs 1464 TFLOP/s on 64-token pages at 1024 queries (profiles/r2_short_pages_pair.log)
  const bool pair_sp = pair8 && avg_chunks_all < 32.0;
  // six query tiles (21-24 queries) fit ONE group of a pair with R = 3 per CTA and one epilogue group: no idle
  // slot, no query-tile reloads, every page tile read once (1376 vs 1311 TFLOP/s, profiles/r2_pair_sweep.log)
  const bool pair6 = !mode && prm.nqt == 6 && avg_chunks_all >= 32.0 && argmax == nullptr && k_blocks == 1 &&
                     pairs_schedulable(h);
  const bool want_pair = (mode && strcmp(mode, "pair") == 0) || pair8 || pair6;
  // bf16x3: one resident query tile per CTA (64 KB) and 2 x 64 KB page-tile stages; CTA pairs (2 query tiles
  // per page tile, half a tile per CTA: 21 instead of 42 B/clk of L2 -> SM traffic per SM) unless an odd tile
  // count would idle more than ~10 % of the slots
  const bool pair_x3 = k_blocks == 3 && !argmax && prm.nqt >= 2 && (prm.nqt % 2 == 0 || prm.nqt >= 9) &&
                       (!mode || strcmp(mode, "1cta") != 0) && pairs_schedulable(h);
  const bool pair = (want_pair && prm.nqt >= 2 && argmax == nullptr && k_blocks == 1) || pair_x3;
  int R, tpg;
  if (pair) { R = (prm.nqt + 1) / 2; if (R > 4) R = 4; if (pair8) R = 2; tpg = 2 * R; }
  else { R = prm.nqt >= 3 ? 3 : prm.nqt; tpg = R; }
  const bool pair8_on = pair && pair8;
  if (argmax) { R = 3; tpg = 3; }   // the argmax variant is instantiated for R = 3 only
  if (k_blocks == 3) { R = 1; tpg = pair_x3 ? 2 : 1; }
  // pages shorter than half a tile end several times per tile: the epilogue (not the MMA) is the
  // limit there -> two epilogue groups, one per resident query tile
  // Two epilogue groups (8 warps, R = 2) measured best whenever there are enough query tiles to
  // balance groups of two: pages shorter than two tiles use the end-mask + staged-sum epilogue (SP,
  // 1480-1494 TFLOP/s from 16- to 208-token pages vs 775-1195 with one group), longer pages the tree
  // fold (1500 vs 1478 TFLOP/s at 1024 queries).  EVDR_TC_MODE=1cta keeps one group with R <= 3.
  const bool one_group = mode && strcmp(mode, "1cta") == 0;
  const bool two_groups_ok = !pair && !argmax && k_blocks == 1 && !one_group;
  const double avg_chunks = (double)total_chunks / (double)np;
  const bool force_sp8 = mode && strcmp(mode, "sp8") == 0;
  const bool short_pages = two_groups_ok && prm.nqt >= 2 && (avg_chunks < 32.0 || force_sp8);
  // one query tile (<= 4 queries) on short pages: the short-page epilogue with one epilogue group ("sp4")
  const bool one_tile_sp = !pair && !argmax && k_blocks == 1 && prm.nqt == 1 &&
                           ((mode && strcmp(mode, "sp4") == 0) ||
                            (!mode && scan_avg_tok < (nq == 1 ? 160.0 : nq == 2 ? 400.0 : 512.0)));
  const bool want_ew8 = two_groups_ok && !short_pages &&
                        (prm.nqt >= 16 || (mode && strcmp(mode, "ew8") == 0 && prm.nqt >= 2));
  // Query tiles in TMEM + a 4-stage ring of 192-token page tiles: the default for exactly two query tiles
  // (5-8 queries) on long pages, where the scan is HBM-bound and the deeper ring is worth 1-12 % (4.9-5.8 ms
  // vs 5.2-5.9 ms per 100k pages, profiles/r1g_tmem_resident_query_tiles.log); everywhere else it measured
  // equal or slower and stays opt-in (EVDR_TC_MODE=ta).
  const bool want_ta = two_groups_ok && prm.nqt >= 2 &&
                       ((mode && strcmp(mode, "ta") == 0) || (!mode && prm.nqt == 2 && !short_pages));
  if (short_pages || want_ew8 || want_ta) { R = 2; tpg = 2; }
  const int n_qgroups = (prm.nqt + tpg - 1) / tpg;
  const int units = pair ? h->sm_count / 2 : h->sm_count;
  int ppi = pages_per_item;
  if (ppi <= 0) { const char* e = getenv("EVDR_TC_PPI"); if (e) ppi = atoi(e); }
  if (ppi <= 0) {
    // ~64K tokens per item (amortises the query-tile reload between items), tile-aligned for
    // uniform pages ...
    const double avg_tok = (double)total_chunks * EVDR_CHUNK / (double)np;
    ppi = (int)(65536.0 / avg_tok + 0.5);
    if (ppi < 1) ppi = 1;
    // ... with few query groups every page block is re-read by only a handful of CTAs while the grid
    // spans units / n_qgroups blocks: keep that window L2-resident (~48 MB of the 126 MB) ...
    if (n_qgroups > 1 && n_qgroups < 8) {
      const double cap = 48e6 * n_qgroups / ((double)units * avg_tok * EVDR_TC_DIM * 2 * (k_blocks == 3 ? 2 : 1));
      if (ppi > cap) ppi = cap < 8.0 ? 8 : (int)cap;   // >= 8 pages: keep the query-tile reload amortised
    } else if (n_qgroups >= 8 && n_qgroups < 64) {
      // ... from 8 groups on the CTAs of a block run in step (dynamic scheduler) and the window is not
      // the limit: measured DRAM traffic stays 1.03x algorithmic up to 64K-token items at 64 and 256
      // queries, and 32K-token items amortise the query-tile reload at every item change best
      // (64 queries: 1444 -> 1494 TFLOP/s; profiles/r1g_item_size_sweep.log)
      const int p32k = (int)(32768.0 / avg_tok + 0.5);
      if (ppi > p32k) ppi = p32k < 1 ? 1 : p32k;
    }
    if (ppi >= 16) ppi = (ppi / 16) * 16;
    // ... and there must be enough items to balance the grid -- but every item change reloads the query
    // tiles (~2 us, about 3.5 tile passes, with the tensor pipe idle), which dominates SMALL problems when items
    // shrink to a page or two (a 32 x 1000-page teacher call ran 12-pass items: 25 % reload time).  With
    // P passes per CTA in total, items of p passes lose ~3.5/p to reloads and ~p/(2P) to the tail: the optimum
    // is p = sqrt(7 P).  Items stay within the caps above and never fewer than one per CTA.
    {
      const int bn = (k_blocks == 3) ? (pair_x3 ? 256 : 128) : (want_ta ? 192 : 256);
      const double avg_tok = (double)total_chunks * EVDR_CHUNK / (double)np;
      const double passes_per_cta = (double)total_chunks * EVDR_CHUNK / bn * prm.nqt / (double)units;
      const double item_passes = sqrt(7.0 * passes_per_cta);
      int ppi_bal = (int)(item_passes * bn / (avg_tok * tpg) + 0.5);
      if (ppi_bal < 1) ppi_bal = 1;
      if (ppi > ppi_bal) ppi = ppi_bal;
      while (ppi > 1 && (long long)((np + ppi - 1) / ppi) * n_qgroups < (long long)units) ppi = (ppi + 1) / 2;
    }
  }
  if (ppi > np) ppi = np;
  prm.ppi = ppi;
  prm.n_pblocks = (np + ppi - 1) / ppi;
  prm.n_qgroups = n_qgroups;
  prm.g_base = prm.nqt / n_qgroups;
  prm.g_rem = prm.nqt % n_qgroups;
  prm.stream_hint = (n_qgroups == 1) ? 1 : 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (tk) {
    // the top-k is folded into the scan when every CTA keeps the same queries for the whole launch (one query
    // group: <= 4 queries, or 5-8 with the query tiles in TMEM) and the long-page epilogue is used
    const bool r1 = !pair && !argmax && k_blocks == 1 && prm.nqt == 1;
    const bool ta2 = want_ta && !short_pages && n_qgroups == 1;
    if (!(r1 || ta2) || tk->k > 32) return EVDR_ERR_UNSUPPORTED;
    const long long n_items = (long long)prm.n_pblocks * prm.n_qgroups;
    tk->n_cta = (int)(n_items < h->sm_count ? n_items : h->sm_count);
    prm.cand_s = tk->cand_s;
    prm.cand_i = tk->cand_i;
    prm.tk_k = tk->k;
    prm.id_offset = tk->id_offset;
    if (ta2) return launch_tc<2, 192, 4, false, false, 1, false, 8, true, true>(h, Qb, tokens, n_tokens, prm, st);
    if (one_tile_sp) return launch_tc<1, 256, 2, false, false, 1, true, 4, false, true>(h, Qb, tokens, n_tokens, prm, st);
    return launch_tc<1, 256, 3, false, false, 1, false, 4, false, true>(h, Qb, tokens, n_tokens, prm, st);
  }
  if (k_blocks == 3) {
    if (argmax) return launch_tc<1, 128, 2, false, true, 3>(h, Qb, tokens, n_tokens, prm, st);
    if (pair_x3) return launch_tc<1, 256, 2, true, false, 3>(h, Qb, tokens, n_tokens, prm, st);
    return launch_tc<1, 128, 2, false, false, 3>(h, Qb, tokens, n_tokens, prm, st);
  }
  if (argmax) return launch_tc<3, 256, 2, false, true>(h, Qb, tokens, n_tokens, prm, st);
  if (pair) {
    if (pair8_on && pair_sp) return launch_tc<2, 256, 4, true, false, 1, true, 8>(h, Qb, tokens, n_tokens, prm, st);
    if (pair8_on) return launch_tc<2, 256, 3, true, false, 1, false, 8>(h, Qb, tokens, n_tokens, prm, st);
    if (R == 1) return launch_tc<1, 256, 6, true>(h, Qb, tokens, n_tokens, prm, st);
    if (R == 2) return launch_tc<2, 256, 5, true>(h, Qb, tokens, n_tokens, prm, st);
    if (R == 3) return launch_tc<3, 256, 4, true>(h, Qb, tokens, n_tokens, prm, st);
    return launch_tc<4, 256, 3, true>(h, Qb, tokens, n_tokens, prm, st);
  }
  // query tiles in TMEM (EVDR_TC_MODE=ta): 192-token page tiles, 4-stage ring
  if (want_ta) {
    if (short_pages) return launch_tc<2, 192, 4, false, false, 1, true, 8, true>(h, Qb, tokens, n_tokens, prm, st);
    return launch_tc<2, 192, 4, false, false, 1, false, 8, true>(h, Qb, tokens, n_tokens, prm, st);
  }
  if (short_pages) return launch_tc<2, 256, 2, false, false, 1, true, 8>(h, Qb, tokens, n_tokens, prm, st);
  if (want_ew8) return launch_tc<2, 256, 2, false, false, 1, false, 8>(h, Qb, tokens, n_tokens, prm, st);
  if (one_tile_sp) return launch_tc<1, 256, 2, false, false, 1, true, 4>(h, Qb, tokens, n_tokens, prm, st);
  if (R == 1) return launch_tc<1, 256, 3, false>(h, Qb, tokens, n_tokens, prm, st);
  if (R == 2) return launch_tc<2, 256, 2, false>(h, Qb, tokens, n_tokens, prm, st);
  return launch_tc<3, 256, 2, false>(h, Qb, tokens, n_tokens, prm, st);
}

int evdr::scan_topk(evdr_handle* h, const void* Qb, const uint8_t* qmask32, int nq, const void* tokens,
                    int64_t n_tokens, const int32_t* chunk_off, int np, int64_t total_chunks, int k,
                    int64_t id_offset, float* cand_s, long long* cand_i, int* n_cta, cudaStream_t st) {
  TkRequest tk = {cand_s, cand_i, k, id_offset, 0};
  const int rc = packed_dispatch(h, Qb, qmask32, nq, tokens, n_tokens, chunk_off, np, total_chunks, 0, nullptr, 0,
                                 nullptr, EVDR_TC_LQ, 1, st, &tk);
  *n_cta = tk.n_cta;
  return rc;
}

extern "C" int evdr_maxsim_packed_fwd(evdr_handle* h, const void* Qb, const uint8_t* qmask32, int nq,
                                      const void* tokens, int64_t n_tokens,
                                      const int32_t* chunk_off, int np, int64_t total_chunks,
                                      int pages_per_item, int k_blocks, float* out, int64_t out_ld,
                                      void* stream) {
  if (!out) return set_error(EVDR_ERR_INVALID, "evdr_maxsim_packed_fwd: out is null");
  return packed_dispatch(h, Qb, qmask32, nq, tokens, n_tokens, chunk_off, np, total_chunks,
                         pages_per_item, out, out_ld, nullptr, EVDR_TC_LQ, k_blocks, stream);
}

extern "C" int evdr_maxsim_packed_argmax_scores(evdr_handle* h, const void* Qb, const uint8_t* qmask32, int nq,
                                                int lq, const void* tokens, int64_t n_tokens,
                                                const int32_t* chunk_off, int np, int64_t total_chunks,
                                                int k_blocks, int32_t* argmax_slots, float* out, int64_t out_ld,
                                                void* stream) {
  if (!argmax_slots || !out)
    return set_error(EVDR_ERR_INVALID, "evdr_maxsim_packed_argmax_scores: null output pointer");
  if (lq < 1 || lq > EVDR_TC_LQ)
    return set_error(EVDR_ERR_INVALID, "evdr_maxsim_packed_argmax_scores: lq=%d outside [1,32]", lq);
  if (out_ld < np) return set_error(EVDR_ERR_INVALID, "evdr_maxsim_packed_argmax_scores: out_ld < np");
  return packed_dispatch(h, Qb, qmask32, nq, tokens, n_tokens, chunk_off, np, total_chunks, 0, out, out_ld,
                         argmax_slots, lq, k_blocks, stream);
}

extern "C" int evdr_maxsim_packed_argmax(evdr_handle* h, const void* Qb, const uint8_t* qmask32, int nq,
                                         int lq, const void* tokens, int64_t n_tokens,
                                         const int32_t* chunk_off, int np, int64_t total_chunks,
                                         int k_blocks, int32_t* argmax_slots, void* stream) {
  if (!argmax_slots) return set_error(EVDR_ERR_INVALID, "evdr_maxsim_packed_argmax: argmax_slots is null");
  if (lq < 1 || lq > EVDR_TC_LQ) return set_error(EVDR_ERR_INVALID, "evdr_maxsim_packed_argmax: lq=%d outside [1,32]", lq);
  return packed_dispatch(h, Qb, qmask32, nq, tokens, n_tokens, chunk_off, np, total_chunks, 0, nullptr,
                         0, argmax_slots, lq, k_blocks, stream);
}
