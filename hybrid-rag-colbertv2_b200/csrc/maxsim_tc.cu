// maxsim_tc.cu — tensor-core MaxSim: TMA -> tcgen05.mma (TMEM accumulators) -> fused segmented
// max / query-token-sum epilogue.  The [Lq x Ld] similarity matrix never leaves the SM.
//
// Data layout in HBM
//   tokens  : bf16 [total_tokens][128], packed, padding-free           (TMA map: 2-D, 128B swizzle)
//   offsets : int64 [n_docs + 1] CSR document boundaries
//   queries : bf16 [n_queries][lq][128]; 32 query tokens per tile slot, a longer query (lq <= 256) is scored as
//             ceil(lq / 32) "virtual queries" whose partial scores are summed in slot order (sum_slots_kernel)
//
// Two kernels (DESIGN.md §4 has the measurements behind every choice):
//
// maxsim_dm_kernel<TK> — ONE query, DOC-MAJOR: A = document tokens (M = 128), B = the query (N = 32), so exactly the
//   useful tensor work is issued; a tile is 4 streams x 32 tokens, one stream of whole documents per epilogue warp;
//   the max over a document's tokens is one redux.sync.max.f32 per query token.  HRC_PATH_AUTO's choice for a single
//   query; TK fuses the per-CTA top-k into the epilogue (hrc_search = 2 launches).  Work is handed out in UNITS of
//   whole documents: one own unit per CTA (7/8 of the corpus) + small shared units claimed at run time, resolved one or
//   two units ahead by a scheduler warp (SMs pull data at different rates; equal shares left 5 % waiting for the
//   slowest GPC).  See the comments above it.
//
// maxsim_tc_kernel<MT, ZP, CG, TK, RR> — QUERY-MAJOR: A = queries (M = 4 slots x 32 tokens), B = document tokens
//   (N = 128): D[row = query token][col = doc token].  One CTA = one contiguous run of whole documents ("segment") x
//   one group of 4*MT (virtual) queries.
//   warp 0      : TMA producer  — streams 128-token x 128-dim tiles through a shared-memory ring
//   warp 1      : MMA issuer    — per tile and M-tile: 8 x tcgen05.mma (M=128 [256 over a CTA pair], N=128, K=16),
//                                 accumulator = 128 TMEM columns; owns TMEM alloc/dealloc
//   warps 2..   : epilogue      — a warp owns TMEM lanes 32*(w%4).. = one query slot: thread i holds query
//                                 token i, so the max over a document's tokens is a per-thread FMNMX3 tree over
//                                 32 TMEM columns (loads aligned to the DOCUMENT, not the tile: no masking for
//                                 >= 32-token pieces), document boundaries are warp-uniform, and a document's
//                                 score is one shuffle butterfly, emitted after the tile has been released.
//                                 Every warp owns WHOLE documents (no cross-warp combine).
//   <1,ZP=1>          2-4 queries (HBM-bound) and candidate mode: the unused rows of the A tile are zero, the 4
//                     epilogue warps are stacked on the used lane groups, 5-stage smem ring, 4 accumulator stages,
//                     one tcgen05.commit per tile.  TK: fused top-k (one query, explicit HRC_PATH_TC).
//                     RR: candidate (rerank) mode with the sorted top-k written by the last CTA of a query to finish
//                     (hrc_rerank = one launch).
//   <2,ZP=0,CG=2>     tensor-bound, batched: CTA PAIR (cta_group::2, cluster of 2), one M=256 MMA per K slice,
//                     each CTA stages half of every document tile; 8 epilogue warps per CTA (two per lane
//                     group, alternating documents), 2 x 2 accumulators.            [from 9 queries]
//   <2,ZP=0,CG=1>     single-CTA batched kernel (5..8 queries, or an odd last query group)
//   The organisations measured slower — query replicated over the A tile, A operand in TMEM, per-M-tile accumulator
//   units (round 1), the M=64 variant (round 2) — are no longer in the source: profiles/experiments/ keeps their diffs.
//
// The product library reads no environment variables.  Building with -DHRC_EXPERIMENTS (make exp -> libhrc_exp.so)
// adds hrc_exp_set_debug() / hrc_exp_set_stages() / hrc_exp_set_ctas() / hrc_exp_set_dyn() / hrc_exp_set_cta_times():
// kernel skeletons (no epilogue math / no document TMA / no MMA), the shared-memory-ring cap, the CTA count, the
// dynamic share and per-CTA finish times behind the measurements in DESIGN.md §4 (scripts/exp_*.py).
//
// Reference semantics: local_rag_complete.py:807-812 (docstring), :813-817 (shapes), summed over
// query tokens per BASELINE.json north_star.  Algorithmic traffic: 256 B per document token.
#include <cuda.h>
#include <climits>
#include <cstdio>
#include <mutex>

#include "hrc_common.cuh"

namespace hrc {

namespace {

constexpr int TN = 128;                           // document tokens per tile (MMA N; N/2 per CTA of a pair)
constexpr int kSlotBytes = 32 * 128;              // one 32-row query slot inside a 64-dim slab
constexpr int kTmemCols = 512;
constexpr int kEpiWarp0 = 2;
constexpr int kMaxSmem = 232448;                  // 227 KB opt-in limit per CTA

// Fused top-k (TK kernels): every epilogue warp keeps, per query it scores, a shared-memory list of the best keys it
// has emitted: [0, k) survivors of the last compaction + appended keys above the current k-th best.  A full list is
// compacted by a warp-wide bitonic sort; at the end the warps of a CTA that share a query merge their lists and the
// CTA hands kListOut keys per (query, segment) to the merge kernel.  The [n_queries x n_docs] score matrix is never
// written (unless the caller asks for it).
constexpr int kListCap = kKeyListCap;             // keys per list (warp_sort256_desc, hrc_common.cuh)
constexpr int kListOut = kKeyListOut;             // keys per (query, segment) handed to the merge: k <= 128

#ifdef HRC_EXPERIMENTS
#define HRC_DBG(p, bit) (((p).debug & (bit)) != 0)
#else
#define HRC_DBG(p, bit) false
#endif

__host__ __device__ constexpr int epi_warps(int mt) { return mt == 2 ? 8 : 4; }
// ZP ("zero padded") variants for <= 4 queries: with slots_used = 1, 2 or 4 queries the other rows of the A tile
// stay ZERO (measured: replicating the query instead costs ~10 % under sustained load, because the extra
// tensor-core switching power pushes the GPU into its 1 kW cap), and the 4 epilogue warps are stacked on the
// USED TMEM lane groups: a warp can only read lanes 32*(warp%4).., so with one query they are warps 4, 8, 12, 16
// (all on lane group 0, splitting the documents 4 ways), with two queries warps 4, 5, 8, 9, with four 4..7.
// The warps in between have no role and exit.
// (An M=64 variant — half the tensor work, a query's token halves in two lane groups combined with atomicAdd — was
// measured between the two at the power cap, C2 4.51-4.99 ms against 4.41-4.89 for the doc-major kernel, and removed:
// profiles/experiments/r02_removed_m64_variant.diff.)
__host__ __device__ constexpr int cta_threads(int mt, int zp) {
  return zp == 1 ? 17 * 32 : (2 + epi_warps(mt)) * 32;
}

struct TcParams {
  const int64_t* offsets;
  const int32_t* cand_ids;  // nullptr: corpus mode
  float* scores;            // TK kernels: may be nullptr (scores are not materialised)
  uint64_t* cand_keys;      // TK kernels: [n_queries][n_segments][kListOut] best keys per (query, segment), sorted
  int k;                    // TK kernels: 1 <= k <= kListOut
  int32_t id_base;          // TK kernels: global id of document 0
  // candidate mode, fused rerank (all null / 0 otherwise): the LAST CTA of a query to finish ranks its n_items scores
  uint32_t* rr_counter;     // [n_queries] zero on entry; reset to zero by the last CTA
  int rr_k;                 // results per query
  int32_t* rr_pos;          // [n_queries][rr_k] position in the candidate list
  int32_t* rr_ids;          // [n_queries][rr_k] candidate document id (optional)
  float* rr_scores;         // [n_queries][rr_k]
  int64_t n_docs;
  int64_t total_tokens;
  int64_t n_items;          // row stride of scores (n_docs, or n_cand)
  int n_queries;            // (virtual) queries: one per 32-token slot of a real query
  int q_slots;              // 32-token slots per real query: ceil(lq / 32); virtual query v = real v / q_slots, slot v % q_slots
  int vq_base;              // first virtual query of this launch
  int n_segments;           // corpus mode: CTAs along the corpus
  int n_qgroups;            // corpus mode: query groups (4*MT queries each)
  int n_stages;             // smem ring depth
  int slots_used;           // distinct queries per A tile: 1, 2 or 4
  int n_units;              // doc-major kernel: work units (>= n_segments); the first n_segments are the CTAs' own
  int64_t static_tokens;    // doc-major kernel: tokens [0, static_tokens) form the n_segments own units, the rest the shared ones
  uint32_t* unit_counter;   // doc-major kernel: shared units are claimed here (zeroed before the launch); nullptr = none
  int debug;                // HRC_EXPERIMENTS builds only
  unsigned long long* cta_times;   // HRC_EXPERIMENTS builds only (hrc_exp_set_cta_times): [2 * CTAs] end time (ns), SM id
  uint64_t watchdog_ns;     // mbarrier waits trap after this long (0 = never)
  uint64_t doc_policy;      // L2 policy for document tiles (evict-first when read once)
};

// Spin on an mbarrier with a wall-clock watchdog: a protocol bug must fault, not hang the GPU.
__device__ __forceinline__ void mbar_wait_wd(uint64_t* bar, uint32_t parity, uint64_t limit_ns) {
  if (mbar_try_wait(bar, parity)) return;
  uint64_t t0 = 0;
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3ffu) == 0 && limit_ns != 0) {
      uint64_t now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      if (now - t0 > limit_ns) {
        printf("hrc: mbarrier timeout block %d thread %d\n", blockIdx.x, threadIdx.x);
        __trap();
      }
    }
  }
}

// first d in [0, n] with offsets[d] >= target
__device__ __forceinline__ int64_t lower_bound_doc(const int64_t* __restrict__ offsets, int64_t n,
                                                   int64_t target) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (offsets[mid] < target) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// 3-input max: one FMNMX3
__device__ __forceinline__ float max3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }
// max(m, 32 accumulator columns) in 16 FMNMX3 (the minimum: each removes two values), depth 4 — the ALU pipe
// issues a warp instruction every 2 cycles, so the instruction count IS the cost of the epilogue's arithmetic
__device__ __forceinline__ float max32_acc(const uint32_t (&v)[32], float m) {
  float a[10];
#pragma unroll
  for (int i = 0; i < 10; ++i)
    a[i] = max3(__uint_as_float(v[3 * i]), __uint_as_float(v[3 * i + 1]), __uint_as_float(v[3 * i + 2]));
  const float b0 = max3(a[0], a[1], a[2]);
  const float b1 = max3(a[3], a[4], a[5]);
  const float b2 = max3(a[6], a[7], a[8]);
  const float b3 = max3(a[9], __uint_as_float(v[30]), __uint_as_float(v[31]));
  return max3(max3(b0, b1, b2), b3, m);
}
// same, over the columns whose bit is set (a piece of a document shorter than 32 columns)
__device__ __forceinline__ float max32_masked_acc(const uint32_t (&v)[32], uint32_t bits, float m) {
  uint32_t t[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) t[i] = ((bits >> i) & 1u) ? v[i] : 0xff800000u;   // -inf
  return max32_acc(t, m);
}

template <int MT, int ZP, int CG, bool TK, bool RR>
__global__ void __launch_bounds__(cta_threads(MT, ZP), 1)
maxsim_tc_kernel(const __grid_constant__ CUtensorMap tmap_d, const __grid_constant__ CUtensorMap tmap_q,
                 const TcParams p) {
  constexpr int kEpiWarps = epi_warps(MT);
  constexpr int kM = 128;                              // MMA M (rows of the A tile)
  constexpr int kQBytes = kM * HRC_DIM * 2;            // one A tile in shared memory
  constexpr int kSplit = kEpiWarps / 4;                 // warps sharing one TMEM lane group split the documents
  // CG == 2 (CTA pair, cta_group::2): the pair issues one M=256 MMA per K slice; this CTA stages TN/2 tokens
  // of every tile (its half of the B operand), its own MT query tiles and its own accumulators.
  constexpr int kTileRows = TN / CG;                    // document tokens of a tile staged by THIS CTA
  constexpr int kTileBytes = kTileRows * HRC_DIM * 2;
  constexpr int kHalfTileBytes = kTileBytes / 2;        // one 64-dim (128-byte-row) slab
  constexpr int kTileStages = kTmemCols / (MT * TN);    // tiles in flight between MMA and epilogue (4 or 2)
  constexpr uint32_t kIdesc = make_idesc_bf16_f32(kM * CG, TN);
  static_assert(kTileBytes % 2048 == 0, "tile slabs must stay 1024-byte aligned for the 128B swizzle");
  static_assert((MT == 1 && ZP == 1 && CG == 1) || (MT == 2 && ZP == 0 && (CG == 1 || CG == 2)),
                "instantiations: <1,1,1> <2,0,1> <2,0,2>");
  static_assert(!TK || (MT == 1 && ZP == 1), "fused top-k: the single-query kernel only (see tc_topk_supported)");
  static_assert(!RR || (MT == 1 && ZP == 1 && !TK), "fused rerank: the candidate-mode instantiation only");
  // A tile's accumulators (all M-tiles) are ONE unit: one tfull / tempty pair per stage.  The warps of a lane group
  // alternate DOCUMENTS and each reads all M-tiles of a tile in one walk.
  // HBM-bound kernels (MT == 1): ONE tcgen05.commit per tile (tfull); the shared-memory slot is released by the
  // first epilogue warp when it sees tfull (the same event, ~100 cycles later, irrelevant with a 6-deep TMA ring).
  // In-process A/B: C2 4.79 -> 4.70 ms, ragged 10.32 -> 10.07 ms.  The batched kernels keep the second commit
  // (1,349 vs 1,332 TFLOP/s).
  constexpr bool kForwardEmpty = MT == 1;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                        // MT x A tile
  uint8_t* sD = smem + MT * kQBytes;                         // n_stages x tile
  uint64_t* bars = reinterpret_cast<uint64_t*>(sD + p.n_stages * kTileBytes);
  uint64_t* full = bars;                           // [n_stages]    TMA -> MMA
  uint64_t* empty = bars + 10;                     // [n_stages]    MMA -> TMA
  uint64_t* tfull = bars + 20;                     // [kTileStages] MMA -> epilogue
  uint64_t* tempty = bars + 24;                    // [kTileStages] epilogue -> MMA
  uint64_t* qfull = bars + 28;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 29);
  int64_t* seg = reinterpret_cast<int64_t*>(bars + 30);  // [0]=doc_begin [1]=doc_end [2]=tok_begin [3]=tok_end
  uint64_t* lists = bars + 64;                           // TK: [kEpiWarps][kListCap] keys

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = CG == 2 ? cluster_ctarank() : 0u;   // 0 = leader of the pair (issues the MMAs)
  const uint64_t wd = p.watchdog_ns;

  // ---- which segment / queries does this CTA own? -------------------------------------------
  int q_base;   // first query of slot 0, M-tile 0
  int64_t item = 0;
  if (p.cand_ids == nullptr) {
    q_base = p.vq_base + int(blockIdx.x % p.n_qgroups) * 4 * MT;   // query groups vary fastest: CTAs sharing a
    item = blockIdx.x / p.n_qgroups;                                // corpus segment run together (L2 reuse)
  } else {
    q_base = blockIdx.y;
    item = blockIdx.x;
  }
  if (threadIdx.x == 0) {
    int64_t d0, d1;
    if (p.cand_ids == nullptr) {
      const int64_t b0 = (p.total_tokens * item) / p.n_segments;
      const int64_t b1 = (p.total_tokens * (item + 1)) / p.n_segments;
      d0 = lower_bound_doc(p.offsets, p.n_docs, b0);
      d1 = (item + 1 == p.n_segments) ? p.n_docs : lower_bound_doc(p.offsets, p.n_docs, b1);
      if (item == 0) d0 = 0;
    } else {
      const int64_t id = p.cand_ids[int64_t(q_base / p.q_slots) * p.n_items + item];   // the REAL query's list
      if (id < 0 || id >= p.n_docs) {
        p.scores[int64_t(q_base) * p.n_items + item] = -INFINITY;
        d0 = d1 = 0;
      } else {
        d0 = id; d1 = id + 1;
      }
    }
    seg[0] = d0; seg[1] = d1;
    seg[2] = p.offsets[d0];
    seg[3] = p.offsets[d1];
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_d);
    tma_prefetch_desc(&tmap_q);
    for (int i = 0; i < p.n_stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < kTileStages; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], kEpiWarps * CG); }
    mbar_init(qfull, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    if constexpr (CG == 2) { tmem_alloc_cg2(tmem_slot, kTmemCols); tmem_relinquish_cg2(); }
    else { tmem_alloc(tmem_slot, kTmemCols); tmem_relinquish(); }
  }
  tc_fence_before_sync();
  if constexpr (CG == 2) cluster_sync_all();   // the peer's barriers must be initialised before anything arrives on them
  else __syncthreads();
  tc_fence_after_sync();

  const uint32_t acc_base = *tmem_slot;
  const int64_t doc_begin = seg[0], doc_end = seg[1], tok_begin = seg[2], tok_end = seg[3];
  const int n_tiles = int((tok_end - tok_begin + TN - 1) / TN);

  if (warp == 0) {
    // =============================== TMA producer =============================================
    // (elect.sync, not `lane == 0`: the compiler then knows a single thread runs this and feeds the
    //  uniform datapath directly instead of emitting a per-instruction uniformisation loop)
    if (n_tiles > 0 && elect_one()) {
      if (cta_rank == 0) mbar_arrive_expect_tx(qfull, CG * MT * kQBytes);   // the peer's tiles count here too
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        // slot g (32 rows of the A tile) holds query g of this M-tile; unused slots (ZP) and rows >= lq /
        // queries >= n_queries are out of bounds of the map and arrive as zeros.
#pragma unroll
        for (int g = 0; g < kM / 32; ++g) {
          const int vq = (ZP && g >= p.slots_used) ? p.n_queries : q_base + 4 * mt + g;
          // virtual query -> (real query, first token row); a query longer than 32 tokens is scored slot by slot
          const int q = vq / p.q_slots, row0 = (vq % p.q_slots) * 32;
          uint8_t* dst = sQ + mt * kQBytes + g * kSlotBytes;
          if constexpr (CG == 2) {
            tma_load_3d_cg2(dst, &tmap_q, qfull, 0, row0, q, kEvictLast);
            tma_load_3d_cg2(dst + kQBytes / 2, &tmap_q, qfull, 64, row0, q, kEvictLast);
          } else {
            tma_load_3d(dst, &tmap_q, qfull, 0, row0, q, kEvictLast);
            tma_load_3d(dst + kQBytes / 2, &tmap_q, qfull, 64, row0, q, kEvictLast);
          }
        }
      }
      int stage = 0; uint32_t phase = 0;
      for (int t = 0; t < n_tiles; ++t) {
        mbar_wait_wd(&empty[stage], phase ^ 1, wd);
        if (HRC_DBG(p, 2)) {
          if (cta_rank == 0) mbar_arrive(&full[stage]);
        } else {
          const int row = int(tok_begin + int64_t(t) * TN) + int(cta_rank) * kTileRows;
          uint8_t* dst = sD + stage * kTileBytes;
          if constexpr (CG == 2) {
            // both halves of the tile are counted on the LEADER's barrier (the MMA issuer waits there)
            if (cta_rank == 0) mbar_arrive_expect_tx(&full[stage], 2 * kTileBytes);
            tma_load_2d_cg2(dst, &tmap_d, &full[stage], 0, row, p.doc_policy);
            tma_load_2d_cg2(dst + kHalfTileBytes, &tmap_d, &full[stage], 64, row, p.doc_policy);
          } else {
            mbar_arrive_expect_tx(&full[stage], kTileBytes);
            tma_load_2d(dst, &tmap_d, &full[stage], 0, row, p.doc_policy);
            tma_load_2d(dst + kHalfTileBytes, &tmap_d, &full[stage], 64, row, p.doc_policy);
          }
        }
        if (++stage == p.n_stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================================
    // The whole warp walks the tile loop and waits on the barriers; one elected lane issues.  Gating with
    // elect.sync (rather than `lane == 0`) matters: otherwise every tcgen05.mma is preceded by an
    // ELECT/R2UR.BROADCAST loop and the issue rate, not the tensor core, bounds the kernel.
    if (n_tiles > 0 && cta_rank == 0) {
      mbar_wait_wd(qfull, 0, wd);
      tc_fence_after_sync();
      const uint32_t sQ_addr = smem_u32(sQ);
      const uint32_t sD_addr = smem_u32(sD);
      int stage = 0; uint32_t phase = 0;
      int ts = 0; uint32_t tphase = 0;
      for (int t = 0; t < n_tiles; ++t) {
        mbar_wait_wd(&full[stage], phase, wd);
        const uint32_t b_addr = sD_addr + stage * kTileBytes;
        mbar_wait_wd(&tempty[ts], tphase ^ 1, wd);     // the accumulators' previous readers are done
        tc_fence_after_sync();
        if (elect_one()) {
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const uint32_t d_tmem = acc_base + uint32_t((ts * MT + mt) * TN);
#pragma unroll
            for (int k = 0; k < HRC_DIM / 16; ++k) {
              if (HRC_DBG(p, 4)) break;   // perf experiment: no tensor work at all (results are garbage)
              // k-th 16-element K slice: slab (k / 4), 32 bytes per slice inside the 128-byte row
              const uint64_t b_desc = make_kmajor_sw128_desc(b_addr + (k >> 2) * kHalfTileBytes + (k & 3) * 32);
              const uint32_t a_addr = sQ_addr + mt * kQBytes + (k >> 2) * (kQBytes / 2) + (k & 3) * 32;
              if constexpr (CG == 2) umma_bf16_ss_cg2(d_tmem, make_kmajor_sw128_desc(a_addr), b_desc, kIdesc, k > 0 ? 1u : 0u);
              else umma_bf16_ss(d_tmem, make_kmajor_sw128_desc(a_addr), b_desc, kIdesc, k > 0 ? 1u : 0u);
            }
          }
          // commits (CG == 2: multicast, the peer's producer and epilogue wait on their own copies).
          // tfull first: the epilogue is on the critical path, the smem slot is not.
          if constexpr (CG == 2) umma_commit_cg2(&tfull[ts]); else umma_commit(&tfull[ts]);
          if constexpr (!kForwardEmpty) {   // smem slot reusable once these MMAs have read it
            if constexpr (CG == 2) umma_commit_cg2(&empty[stage]); else umma_commit(&empty[stage]);
          }
        }
        __syncwarp();
        if (++stage == p.n_stages) { stage = 0; phase ^= 1; }
        if (++ts == kTileStages) { ts = 0; tphase ^= 1; }
      }
    }
  } else if (ZP == 0 || (warp >= 4 && (warp & 3) < p.slots_used && (warp >> 2) - 1 < 4 / p.slots_used)) {
    // =============================== epilogue ==================================================
    // The warp of (slot g, share `sub`) scores the query in slot g for the documents whose local index is
    // congruent to `residue` modulo `rep`.
    const int slot = warp & 3;                           // TMEM lanes 32*slot .. 32*slot+31
    const uint32_t lane_base = uint32_t(slot * 32) << 16;
    const int sub = ZP ? (warp >> 2) - 1 : (warp - kEpiWarp0) >> 2;
    const int rep = ZP ? 4 / p.slots_used : kSplit;
    const int residue = sub;
    bool active[MT];
    int64_t out_row[MT];
    bool any_active = false;
#pragma unroll
    for (int j = 0; j < MT; ++j) {
      const int q = q_base + 4 * j + slot;
      active[j] = q < p.n_queries;
      out_row[j] = int64_t(q) * p.n_items;
      any_active |= active[j];
    }

    const int n_docs_seg = int(doc_end - doc_begin);
    // Document ends are fetched 32 at a time, one per lane, one batch ahead, and broadcast with a shuffle.  The lanes
    // keep the RAW low word of offsets[] and subtract tok_begin only after the shuffle, so nothing consumes a load
    // right after it is issued (a consumer there stalls the warp for a whole global-memory latency).
    const uint32_t tok_begin_lo = uint32_t(tok_begin);
    const uint32_t raw_none = tok_begin_lo + uint32_t(INT_MAX);     // decodes to INT_MAX: "no such document"
    int batch = 0;
    uint32_t ends = raw_none, ends_next = raw_none;
    auto load_ends = [&](int b) -> uint32_t {
      const int d = b * 32 + lane;
      uint32_t r = raw_none;
      if (d < n_docs_seg) r = uint32_t(p.offsets[doc_begin + d + 1]);   // predicated load, no consumer
      return r;
    };
    auto end_of = [&](int d) -> int {   // d non-decreasing over calls, -1 <= d < n_docs_seg
      if (d < 0) return 0;
#pragma unroll 1
      while ((d >> 5) > batch) {
        ends = ends_next;
        ++batch;
        ends_next = load_ends(batch + 1);
      }
      return int(__shfl_sync(0xffffffffu, ends, d & 31) - tok_begin_lo);
    };

    int my = residue;                       // local index of the document this warp is accumulating
    bool have_doc = any_active && my < n_docs_seg;
    int s_tok = 0, e_tok = 0;               // its token range
    int ns_tok = 0, ne_tok = 0;             // token range of this warp's NEXT document, fetched one document ahead
    if (any_active) {
      ends = load_ends(0);
      ends_next = load_ends(1);
    }
    if (have_doc) {
      s_tok = end_of(my - 1);
      e_tok = end_of(my);
      if (my + rep < n_docs_seg) {
        ns_tok = end_of(my + rep - 1);
        ne_tok = end_of(my + rep);
      }
    }
    float m[MT];
#pragma unroll
    for (int j = 0; j < MT; ++j) m[j] = -INFINITY;

    // Emitting a score (a 5-step shuffle butterfly + a store, ~500 cycles of latency) is taken OFF the
    // accumulator hand-shake: finish_doc only parks the finished maxima; they are reduced and stored after this
    // warp has released the tile (or when the next document of the same tile finishes).
    float pend_m[MT];
    int64_t pend_col = 0;
    bool pending = false;
#pragma unroll
    for (int j = 0; j < MT; ++j) pend_m[j] = 0.f;

    // TK (single-query kernel only): this warp's key list.  After warp_sum EVERY lane holds the document's score, and
    // every lane keeps the list's fill count, threshold key and threshold score, so "does this score enter the list?" is
    // one float compare and a warp-uniform branch — no vote, no shuffle.  Only then is the 64-bit key formed (by every
    // lane, identically), stored by lane 0, and a full list compacted with a bitonic sort.
    const int lidx = slot * rep + sub;                       // 0 .. kEpiWarps-1
    uint64_t* lst = lists + size_t(lidx) * kListCap;
    uint32_t cnt = 0;                                        // warp-uniform
    uint64_t thr = 0;                                        // warp-uniform: the k-th best key so far (0: accept everything)
    float thr_f = -INFINITY;
    if constexpr (TK) {
      for (int i = lane; i < kListCap; i += 32) lst[i] = 0;
      __syncwarp();
    }
    auto offer = [&](float score, int64_t doc) {             // whole warp, uniform arguments
      if (score < thr_f) return;                             // (false for NaN, which make_key orders as -inf)
      const uint64_t key = make_key(score, int32_t(p.id_base + int32_t(doc)));
      if (key <= thr) return;
      if (lane == 0) lst[cnt] = key;
      if (++cnt == uint32_t(kListCap)) {
        warp_sort256_desc(lst, lane);
        cnt = uint32_t(p.k);
        thr = lst[p.k - 1];
        thr_f = key_score(thr);
      }
    };

    auto emit_pending = [&]() {
      float sc_all = 0.f;                   // MT == 1: the document's score, on every lane
      if constexpr (MT == 2) {
        // both M-tiles in ONE butterfly: after the first exchange lanes 0-15 carry M-tile 0 and lanes 16-31 M-tile 1
        const bool lo_half = lane < 16;
        float a = lo_half ? pend_m[0] : pend_m[1];
        const float b = lo_half ? pend_m[1] : pend_m[0];
        a += __shfl_xor_sync(0xffffffffu, b, 16);
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        const bool act = lo_half ? active[0] : active[1];
        const int64_t row = lo_half ? out_row[0] : out_row[1];
        if ((lane & 15) == 0 && act) p.scores[row + pend_col] = a;
      } else {
        const float sc = warp_sum(pend_m[0]);
        sc_all = sc;
        if (lane == 0 && active[0]) {
          if (!TK || p.scores != nullptr) p.scores[out_row[0] + pend_col] = sc;
          if constexpr (RR) __threadfence();                // fused rerank: the score must be visible to the last CTA
        }
      }
      if constexpr (TK) {
        if (active[0]) offer(sc_all, pend_col);
      }
      pending = false;
    };

    auto finish_doc = [&]() {   // park the score(s) of document `my`, move to this warp's next document
      if (pending) emit_pending();
#pragma unroll
      for (int j = 0; j < MT; ++j) { pend_m[j] = m[j]; m[j] = -INFINITY; }
      pend_col = (p.cand_ids == nullptr) ? (doc_begin + my) : item;
      pending = true;
      my += rep;
      have_doc = my < n_docs_seg;
      s_tok = ns_tok;
      e_tok = ne_tok;
      if (my + rep < n_docs_seg) {          // results are consumed one document later: the shuffle latency is hidden
        ns_tok = end_of(my + rep - 1);
        ne_tok = end_of(my + rep);
      }
    };

    int ts = 0; uint32_t tphase = 0;
    int stage_e = 0;                        // shared-memory slot of the tile being read (kForwardEmpty)
    for (int t = 0; t < n_tiles; ++t) {
      mbar_wait_wd(&tfull[ts], tphase, wd);
      if constexpr (kForwardEmpty) {        // the tile's MMAs are done: its shared-memory slot may be refilled
        if (warp == 4 && lane == 0) mbar_arrive(&empty[stage_e]);
        if (++stage_e == p.n_stages) stage_e = 0;
      }
      tc_fence_after_sync();
      const int tile0 = t * TN, tile1 = tile0 + TN;
      // accumulator columns of this tile for this warp's j-th M-tile: tacc + j * TN + (token position - tile0)
      const uint32_t tacc = acc_base + lane_base + uint32_t(ts * MT * TN);
      uint32_t v[2][32];
      while (have_doc && s_tok < tile1) {
        const int lo = max(s_tok, tile0), hi = min(e_tok, tile1);
        const int len = hi - lo;              // this document's tokens inside this tile
        if (len > 0 && !HRC_DBG(p, 1)) {
          if (len >= 32) {
            // Whole 32-column loads that START AT the document's first column (TMEM columns are addressable one by
            // one); the last load is pulled back so that it ENDS at the document's last column — the overlap is
            // harmless under max — so no column is ever masked.
            const int last = hi - 32;
            int c = lo;
            if constexpr (MT == 2) {
              // software pipeline over (chunk, M-tile) steps: one TMEM load is in flight while the max tree of the
              // previous step runs (v[0] always holds M-tile 0, v[1] M-tile 1)
              uint32_t col = uint32_t(min(c, last) - tile0);
              tmem_ld_32x32(tacc + col, v[0]);
              while (true) {
                tmem_ld_wait();
                tmem_ld_32x32(tacc + uint32_t(TN) + col, v[1]);
                m[0] = max32_acc(v[0], m[0]);
                tmem_ld_wait();
                const bool more = c < last;
                if (more) {
                  c += 32;
                  col = uint32_t(min(c, last) - tile0);
                  tmem_ld_32x32(tacc + col, v[0]);
                }
                m[1] = max32_acc(v[1], m[1]);
                if (!more) break;
              }
            } else {
              while (true) {
                tmem_ld_32x32(tacc + uint32_t(min(c, last) - tile0), v[0]);
                tmem_ld_wait();
                m[0] = max32_acc(v[0], m[0]);
                if (c >= last) break;
                c += 32;
              }
            }
          } else {
            // fewer than 32 of its tokens here (a short document, or the head / tail a tile boundary cut off)
            const int cc = min(lo, tile1 - 32);
#pragma unroll
            for (int j = 0; j < MT; ++j) tmem_ld_32x32(tacc + uint32_t(j * TN) + uint32_t(cc - tile0), v[j]);
            tmem_ld_wait();
            const int a = lo - cc, b = hi - cc;   // 0 <= a < b <= 32, b - a < 32
            const uint32_t bits = ((1u << (b - a)) - 1u) << a;
#pragma unroll
            for (int j = 0; j < MT; ++j) m[j] = max32_masked_acc(v[j], bits, m[j]);
          }
        }
        if (e_tok > tile1) break;             // the document continues in the next tile
        finish_doc();
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG == 2) mbar_arrive_cluster(&tempty[ts], 0);   // the leader's MMA issuer owns both accumulators
        else mbar_arrive(&tempty[ts]);
      }
      if (pending) emit_pending();          // after the release: off the MMA <-> epilogue critical path
      if (++ts == kTileStages) { ts = 0; tphase ^= 1; }
    }
    while (have_doc) finish_doc();          // trailing empty documents (no tokens, no tile): -inf
    if (pending) emit_pending();

    if constexpr (TK) {
      warp_sort256_desc(lst, lane);         // sorted, best first
      // the `rep` warps of a lane group scored disjoint documents for the same query: tree-merge their lists
      // (partner's best kListOut into the upper half, sort).  All epilogue warps meet at the named barrier.
      for (int st = 1; st < rep; st <<= 1) {
        asm volatile("bar.sync 1, %0;" ::"r"(kEpiWarps * 32) : "memory");
        if ((sub & (2 * st - 1)) == 0 && sub + st < rep) {
          const uint64_t* other = lists + size_t(lidx + st) * kListCap;
          for (int i = lane; i < kListOut; i += 32) lst[kListOut + i] = other[i];
          warp_sort256_desc(lst, lane);
        }
      }
      if (sub == 0 && active[0]) {
        uint64_t* out = p.cand_keys + (int64_t(q_base + slot) * p.n_segments + item) * kListOut;
        for (int i = lane; i < kListOut; i += 32) out[i] = lst[i];
      }
    }
  }

  tc_fence_before_sync();
  if constexpr (CG == 2) {
    cluster_sync_all();             // the peer may still be reading operands / arriving on this CTA's barriers
    if (warp == 1) tmem_dealloc_cg2(acc_base, kTmemCols);
  } else {
    __syncthreads();
    if (warp == 1) tmem_dealloc(acc_base, kTmemCols);
  }

  if constexpr (RR) {
    // Fused rerank (candidate mode): every CTA scored ONE candidate; the last CTA of a query to get here ranks the
    // query's n_items scores (rank by counting over 64-bit (score, position) keys: n_items <= 1024) and writes the
    // sorted top rr_k — what torch.argsort(descending)[:k] does at local_rag_complete.py:789-792 — so that a rerank
    // is ONE launch.  The tile ring is idle by now and holds the keys.
    {
      volatile int* s_last = reinterpret_cast<volatile int*>(bars + 40);   // (no static shared memory: the dynamic
                                                                            //  allocation already takes the whole 227 KB)
      const int q = q_base;                               // candidate mode: one (virtual = real) query per blockIdx.y
      if (threadIdx.x == 0) {
        __threadfence();
        *s_last = atomicAdd(&p.rr_counter[q], 1u) == uint32_t(p.n_items) - 1u;
      }
      __syncthreads();
      if (*s_last) {
        __threadfence();
        const int n = int(p.n_items);
        uint64_t* keys = reinterpret_cast<uint64_t*>(sD);
        const float* row = p.scores + int64_t(q) * n;
        for (int i = threadIdx.x; i < n; i += blockDim.x) keys[i] = make_key(__ldcg(row + i), i);
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
          const uint64_t ki = keys[i];
          int rank = 0;
          for (int j = 0; j < n; ++j) rank += keys[j] > ki ? 1 : 0;
          if (rank < p.rr_k) {
            const int64_t o = int64_t(q) * p.rr_k + rank;
            p.rr_pos[o] = i;
            if (p.rr_ids != nullptr) p.rr_ids[o] = p.cand_ids[int64_t(q) * n + i];
            p.rr_scores[o] = key_score(ki);
          }
        }
        if (threadIdx.x == 0) p.rr_counter[q] = 0;
      }
    }
  }
}

// =====================================================================================================================
// maxsim_dm_kernel — the DOC-MAJOR orientation for ONE query: A = document tokens (M = 128), B = the query (N = 32).
//
// The query-major kernel above puts the query on the M axis, so a single 32-token query occupies 32 of the 128 rows of
// every MMA: three quarters of the issued tensor work multiplies zeros.  That costs nothing at burst clocks, but the
// single-query search is what runs for seconds on end, at the GPU's power cap, where every watt spent on the tensor
// pipe is a watt the memory system does not get.  Here D[row = document token][col = query token] = 128 lanes x 32
// columns: exactly the useful work (2 * 32 * 128 flop per token), a quarter of the tensor-pipe time and of the B-operand
// traffic, 16 accumulator stages of 32 TMEM columns instead of 4 of 128.
//
// The price is the epilogue: the max over a document's tokens now runs ACROSS LANES.  To keep it warp-local, a tile is
// not 128 consecutive tokens but 4 x 32: the CTA's segment is cut into FOUR token streams of whole documents, one per
// epilogue warp / TMEM lane quadrant, and tile t holds tokens [32t, 32t + 32) of each stream (4 TMA boxes of 32 rows).
// A warp therefore sees ITS stream chunk by chunk on its own 32 lanes — lane i = token 32t + i of the stream, register
// j = query token j — and reduces a document's piece with ONE warp-wide max instruction per query token
// (redux.sync.max.f32, sm_100a); the 32 running maxima max_t <q_j, d_t> live replicated in every lane across chunks, and
// a finished document is 31 adds in the order of the query-major kernels' butterfly (bit-identical scores).  No
// cross-warp combine, no atomics, documents never split between warps.
// =====================================================================================================================
constexpr int kDmQBytes = 32 * HRC_DIM * 2;         // the query as the B operand: 32 rows, 8 KB
constexpr int kDmTileBytes = 128 * HRC_DIM * 2;      // 4 streams x 32 tokens, 32 KB
constexpr int kDmStages = 5;
constexpr int kDmAcc = 16;                          // accumulator stages: 512 TMEM columns / 32
constexpr int kDmThreads = 7 * 32;                  // TMA warp, MMA warp, 4 epilogue warps, unit scheduler warp
constexpr int kDmBarBytes = 768;                    // mbarriers + the unit-descriptor ring
constexpr int kDmSharedShare = 8;                   // 1 / this of the corpus is handed out dynamically ...
constexpr int kDmSharedPerCta = 16;                 // ... in this many units per CTA
constexpr int64_t kDmDynMinTokensPerCta = 4096;     // below this a shared unit would be shorter than two tiles

// max over the 32 lanes of v, returned to every lane: ONE instruction on sm_100a (CREDUX.MAX.F32, result in a uniform
// register).  scripts/micro/redux_bench.cu: 32 of these + 32 FMNMX take 174 cycles per 32 x 32 block and warp, the 5-level
// shuffle reduce-scatter (31 SHFL + 62 SEL + 31 FMNMX) 346.
__device__ __forceinline__ float lanes_max(float v) {
  float r;
  asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
  return r;
}
// x[0] + ... + x[31] in exactly the association order of warp_sum's xor butterfly over 32 lanes (16, 8, 4, 2, 1), so that
// a score summed in one thread is bit-identical to the query-major kernels' score summed across lanes
__device__ __forceinline__ float sum32_butterfly_order(const float (&x)[32]) {
  float a[16], b[8], c[4];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = x[i] + x[i + 16];
#pragma unroll
  for (int i = 0; i < 8; ++i) b[i] = a[i] + a[i + 8];
#pragma unroll
  for (int i = 0; i < 4; ++i) c[i] = b[i] + b[i + 4];
  return (c[0] + c[2]) + (c[1] + c[3]);
}

// Work distribution.  The corpus is NOT cut into one equal token range per CTA: SMs pull data at different rates (GPCs
// of 16 / 18 / 20 SMs share a port to the L2, near / far die), and with equal ranges ~20 CTAs of one GPC finished 190-220 us
// after the first (scripts/exp_cta_times.py) — 5 % of the kernel waiting for its slowest members.  Instead the first
// 7/8 of the tokens form one OWN unit per CTA and the rest is cut into 16 small SHARED units per CTA, claimed with an
// atomic counter by whoever is free (same-box A/B, scripts/ab_dynamic_units.py / exp_dyn_sweep.py: C2 kernel 4.38-4.42 ms
// against 4.46-4.49 with equal ranges, the CTAs now finish within 63 us of each other; ragged corpus and power-capped
// runs: unchanged).  A dedicated scheduler warp claims units and resolves their document boundaries
// (binary searches over `offsets`, ~10 us of dependent loads) one or two units ahead and hands them to the other warps
// through a two-entry descriptor ring, so the TMA ring, the accumulator ring and the epilogue never drain between units.
// Every unit is a range of WHOLE documents, so the result does not depend on which CTA scored it.
__device__ __forceinline__ int64_t dm_unit_first_token(const TcParams& p, int64_t u) {
  if (u >= p.n_units) return p.total_tokens;
  if (u < p.n_segments) return (p.static_tokens * u) / p.n_segments;
  return p.static_tokens + ((p.total_tokens - p.static_tokens) * (u - p.n_segments)) / (p.n_units - p.n_segments);
}

template <bool TK>
__global__ void __launch_bounds__(kDmThreads, 1)
maxsim_dm_kernel(const __grid_constant__ CUtensorMap tmap_d, const __grid_constant__ CUtensorMap tmap_q, const TcParams p) {
  constexpr uint32_t kIdesc = make_idesc_bf16_f32(128, 32);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                        // 8 KB (1024-aligned)
  uint8_t* sD = smem + kDmQBytes;                            // kDmStages x 32 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(sD + kDmStages * kDmTileBytes);
  uint64_t* full = bars;                           // [kDmStages] TMA -> MMA
  uint64_t* empty = bars + 8;                      // [kDmStages] epilogue (forwarding the MMA's completion) -> TMA
  uint64_t* tfull = bars + 16;                     // [kDmAcc]    MMA -> epilogue
  uint64_t* tempty = bars + 32;                    // [kDmAcc]    epilogue -> MMA
  uint64_t* qfull = bars + 48;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 49);
  uint64_t* ufull = bars + 50;                     // [2] scheduler -> the six consumer warps: a unit descriptor is ready
  uint64_t* uempty = bars + 52;                    // [2] consumers -> scheduler: descriptor read
  int64_t* udesc = reinterpret_cast<int64_t*>(bars + 56);    // [2][11]: first document of each stream (+ end), first token
                                                             // of each stream (+ end), tiles (< 0: no more units)
  uint64_t* lists = bars + kDmBarBytes / 8;                  // TK: [4][kListCap] keys

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint64_t wd = p.watchdog_ns;
  const int64_t item = blockIdx.x;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_d);
    tma_prefetch_desc(&tmap_q);
    for (int i = 0; i < kDmStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < kDmAcc; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
    for (int i = 0; i < 2; ++i) { mbar_init(&ufull[i], 1); mbar_init(&uempty[i], 6); }
    mbar_init(qfull, 1);
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, kTmemCols); tmem_relinquish(); }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();

  const uint32_t acc_base = *tmem_slot;
  int uslot = 0; uint32_t uphase = 0;              // position in the descriptor ring (every warp walks it in step)

  if (warp == 6) {
    // =============================== unit scheduler ===========================================
    int64_t u = item;                                // the CTA's own unit first
    while (true) {
      mbar_wait_wd(&uempty[uslot], uphase ^ 1, wd);
      int64_t* d = udesc + uslot * 11;
      const bool more = u < p.n_units;
      if (more) {
        if (lane < 5) {      // stream boundaries: the unit's token range cut in four at document starts
          const int64_t b0 = dm_unit_first_token(p, u), b1 = dm_unit_first_token(p, u + 1);
          const int64_t d0 = u == 0 ? 0 : lower_bound_doc(p.offsets, p.n_docs, b0);
          const int64_t d1 = (u + 1 == p.n_units) ? p.n_docs : lower_bound_doc(p.offsets, p.n_docs, b1);
          const int64_t t0 = p.offsets[d0], t1 = p.offsets[d1];
          int64_t dd;
          if (lane == 0) dd = d0;
          else if (lane == 4) dd = d1;
          else {
            dd = lower_bound_doc(p.offsets, p.n_docs, t0 + ((t1 - t0) * int64_t(lane)) / 4);
            dd = dd < d0 ? d0 : (dd > d1 ? d1 : dd);
          }
          d[lane] = dd;
          d[5 + lane] = p.offsets[dd];
        }
        __syncwarp();
        if (lane == 0) {
          int64_t longest = 0;
#pragma unroll
          for (int s4 = 0; s4 < 4; ++s4) longest = max(longest, d[5 + s4 + 1] - d[5 + s4]);
          d[10] = (longest + 31) / 32;
        }
      } else if (lane == 0) {
        d[10] = -1;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&ufull[uslot]);
      if (!more) break;
      if (++uslot == 2) { uslot = 0; uphase ^= 1; }
      int64_t nu = p.n_units;                        // claim a shared unit
      if (lane == 0 && p.unit_counter != nullptr) nu = int64_t(p.n_segments) + int64_t(atomicAdd(p.unit_counter, 1u));
      u = __shfl_sync(0xffffffffu, nu, 0);
    }
  } else if (warp == 0) {
    // =============================== TMA producer =============================================
    if (lane == 0) {
      mbar_arrive_expect_tx(qfull, kDmQBytes);
      tma_load_3d(sQ, &tmap_q, qfull, 0, 0, p.vq_base, kEvictLast);                  // rows >= lq arrive as zeros
      tma_load_3d(sQ + kDmQBytes / 2, &tmap_q, qfull, 64, 0, p.vq_base, kEvictLast);
      int stage = 0; uint32_t phase = 0;
      while (true) {
        mbar_wait_wd(&ufull[uslot], uphase, wd);
        const int64_t* d = udesc + uslot * 11;
        const int n_tiles = int(d[10]);
        int64_t row0[4];
        int tiles_of[4];                             // chunks of 32 tokens each stream has in this unit
#pragma unroll
        for (int s4 = 0; s4 < 4; ++s4) {
          row0[s4] = d[5 + s4];
          tiles_of[s4] = int((d[5 + s4 + 1] - d[5 + s4] + 31) / 32);
        }
        mbar_arrive(&uempty[uslot]);
        if (++uslot == 2) { uslot = 0; uphase ^= 1; }
        if (n_tiles < 0) break;
        for (int t = 0; t < n_tiles; ++t) {
          mbar_wait_wd(&empty[stage], phase ^ 1, wd);
          uint8_t* dst = sD + stage * kDmTileBytes;
          if (HRC_DBG(p, 2)) {
            mbar_arrive(&full[stage]);
          } else {
            // a stream that has ended loads nothing more (its quarter of the tile keeps stale tokens: harmless, no
            // document of its warp owns those lanes) — with whole-document streams the four lengths of a unit differ by
            // up to a document, and reading on would waste that much HBM traffic per unit
            int live = 0;
#pragma unroll
            for (int s4 = 0; s4 < 4; ++s4) live += t < tiles_of[s4] ? 1 : 0;
            mbar_arrive_expect_tx(&full[stage], uint32_t(live) * (kDmTileBytes / 4));
#pragma unroll
            for (int s4 = 0; s4 < 4; ++s4) {
              if (t >= tiles_of[s4]) continue;
              const int row = int(row0[s4] + int64_t(t) * 32);
              tma_load_2d(dst + s4 * 4096, &tmap_d, &full[stage], 0, row, p.doc_policy);
              tma_load_2d(dst + kDmTileBytes / 2 + s4 * 4096, &tmap_d, &full[stage], 64, row, p.doc_policy);
            }
          }
          if (++stage == kDmStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================================
    mbar_wait_wd(qfull, 0, wd);
    tc_fence_after_sync();
    const uint32_t sQ_addr = smem_u32(sQ);
    const uint32_t sD_addr = smem_u32(sD);
    int stage = 0; uint32_t phase = 0;
    int ts = 0; uint32_t tphase = 0;
    while (true) {
      mbar_wait_wd(&ufull[uslot], uphase, wd);
      const int n_tiles = int(udesc[uslot * 11 + 10]);
      __syncwarp();
      if (lane == 0) mbar_arrive(&uempty[uslot]);
      if (++uslot == 2) { uslot = 0; uphase ^= 1; }
      if (n_tiles < 0) break;
      for (int t = 0; t < n_tiles; ++t) {
        mbar_wait_wd(&full[stage], phase, wd);
        mbar_wait_wd(&tempty[ts], tphase ^ 1, wd);
        tc_fence_after_sync();
        if (elect_one()) {
          const uint32_t a_tile = sD_addr + stage * kDmTileBytes;
          const uint32_t d_tmem = acc_base + uint32_t(ts * 32);
#pragma unroll
          for (int k = 0; k < HRC_DIM / 16; ++k) {
            if (HRC_DBG(p, 4)) break;
            const uint64_t a_desc = make_kmajor_sw128_desc(a_tile + (k >> 2) * (kDmTileBytes / 2) + (k & 3) * 32);   // documents
            const uint64_t b_desc = make_kmajor_sw128_desc(sQ_addr + (k >> 2) * (kDmQBytes / 2) + (k & 3) * 32);     // query
            umma_bf16_ss(d_tmem, a_desc, b_desc, kIdesc, k > 0 ? 1u : 0u);
          }
          umma_commit(&tfull[ts]);          // the first epilogue warp forwards it to empty[stage]
        }
        __syncwarp();
        if (++stage == kDmStages) { stage = 0; phase ^= 1; }
        if (++ts == kDmAcc) { ts = 0; tphase ^= 1; }
      }
    }
  } else {
    // =============================== epilogue: one stream per warp ============================
    const int quad = warp & 3;                           // TMEM lanes 32*quad.. = rows 32*quad.. of the tile = stream `quad`
    const uint32_t lane_base = uint32_t(quad * 32) << 16;
    const bool q_active = p.vq_base < p.n_queries;
    const int64_t out_row = int64_t(p.vq_base) * p.n_items;

    float mr[32];                           // mr[j]: running max_t <q_j, d_t> of the current document, in every lane
#pragma unroll
    for (int j = 0; j < 32; ++j) mr[j] = -INFINITY;

    // fused top-k state (see the query-major kernel); the list spans all the units this CTA scores
    uint64_t* lst = lists + size_t(quad) * kListCap;
    uint32_t cnt = 0;                        // warp-uniform, like thr and thr_f: every lane holds the score after warp_sum
    uint64_t thr = 0;
    float thr_f = -INFINITY;
    if constexpr (TK) {
      for (int i = lane; i < kListCap; i += 32) lst[i] = 0;
      __syncwarp();
    }
    int ts = 0; uint32_t tphase = 0;
    int stage_e = 0;

    while (true) {
      mbar_wait_wd(&ufull[uslot], uphase, wd);
      const int64_t* ud = udesc + uslot * 11;
      const int64_t doc_begin = ud[quad], doc_end = ud[quad + 1], tok_begin = ud[5 + quad], tok_end = ud[5 + quad + 1];
      const int n_tiles = int(ud[10]);
      __syncwarp();
      if (lane == 0) mbar_arrive(&uempty[uslot]);
      if (++uslot == 2) { uslot = 0; uphase ^= 1; }
      if (n_tiles < 0) break;

      const int n_docs_seg = int(doc_end - doc_begin);
      const int my_chunks = int((tok_end - tok_begin + 31) / 32);
      const uint32_t tok_begin_lo = uint32_t(tok_begin);
      const uint32_t raw_none = tok_begin_lo + uint32_t(INT_MAX);
      int batch = 0;
      uint32_t ends = raw_none, ends_next = raw_none;
      auto load_ends = [&](int b) -> uint32_t {
        const int d = b * 32 + lane;
        uint32_t r = raw_none;
        if (d < n_docs_seg) r = uint32_t(p.offsets[doc_begin + d + 1]);
        return r;
      };
      auto end_of = [&](int d) -> int {   // d non-decreasing over calls, -1 <= d < n_docs_seg
        if (d < 0) return 0;
#pragma unroll 1
        while ((d >> 5) > batch) {
          ends = ends_next;
          ++batch;
          ends_next = load_ends(batch + 1);
        }
        return int(__shfl_sync(0xffffffffu, ends, d & 31) - tok_begin_lo);
      };
      ends = load_ends(0);
      ends_next = load_ends(1);

      int my = 0;                              // local index of the document being accumulated
      bool have_doc = q_active && n_docs_seg > 0;
      int s_tok = 0, e_tok = 0;
      if (have_doc) e_tok = end_of(0);

      auto finish_doc = [&]() {
        const float sc = sum32_butterfly_order(mr);
        const int64_t col = doc_begin + my;
        if (lane == 0 && (!TK || p.scores != nullptr)) p.scores[out_row + col] = sc;
        if constexpr (TK) {
          if (!HRC_DBG(p, 8) && !(sc < thr_f)) {   // one compare per document; warp-uniform (also taken for NaN)
            const uint64_t key = make_key(sc, int32_t(p.id_base + int32_t(col)));
            if (key > thr) {
              if (lane == 0) lst[cnt] = key;
              if (++cnt == uint32_t(kListCap)) {
                warp_sort256_desc(lst, lane);
                cnt = uint32_t(p.k);
                thr = lst[p.k - 1];
                thr_f = key_score(thr);
              }
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) mr[j] = -INFINITY;
        ++my;
        have_doc = my < n_docs_seg;
        s_tok = e_tok;
        if (have_doc) e_tok = end_of(my);
      };

      for (int t = 0; t < n_tiles; ++t) {
        mbar_wait_wd(&tfull[ts], tphase, wd);
        if (quad == 0 && lane == 0) mbar_arrive(&empty[stage_e]);   // the tile's MMAs are done: its smem slot may be refilled
        if (++stage_e == kDmStages) stage_e = 0;
        tc_fence_after_sync();
        if (t < my_chunks && have_doc && !HRC_DBG(p, 1)) {
          uint32_t v[32];
          tmem_ld_32x32(acc_base + lane_base + uint32_t(ts * 32), v);
          tmem_ld_wait();
          const int c0 = t * 32, c1 = c0 + 32;
          while (have_doc && s_tok < c1) {
            const int lo = max(s_tok, c0), hi = min(e_tok, c1);     // this document's tokens inside this chunk: lanes [lo-c0, hi-c0)
            if (hi - lo == 32) {                  // the chunk lies inside the document
#pragma unroll
              for (int j = 0; j < 32; ++j) mr[j] = fmaxf(mr[j], lanes_max(__uint_as_float(v[j])));
            } else if (hi > lo) {                 // a boundary inside the chunk: the other documents' lanes read as -inf
              const bool mine = lane >= lo - c0 && lane < hi - c0;
#pragma unroll
              for (int j = 0; j < 32; ++j) mr[j] = fmaxf(mr[j], lanes_max(mine ? __uint_as_float(v[j]) : -INFINITY));
            }
            if (e_tok > c1) break;              // the document continues in the next chunk
            finish_doc();
          }
        }
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[ts]);
        if (++ts == kDmAcc) { ts = 0; tphase ^= 1; }
      }
      while (have_doc) finish_doc();            // trailing empty documents: -inf
    }

    if constexpr (TK) {
      if (!HRC_DBG(p, 32)) warp_sort256_desc(lst, lane);
      for (int st = 1; st < (HRC_DBG(p, 32) ? 1 : 4); st <<= 1) {
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if ((quad & (2 * st - 1)) == 0) {
          const uint64_t* other = lists + size_t(quad + st) * kListCap;
          for (int i = lane; i < kListOut; i += 32) lst[kListOut + i] = other[i];
          warp_sort256_desc(lst, lane);
        }
      }
      if (quad == 0 && q_active) {
        uint64_t* out = p.cand_keys + (int64_t(p.vq_base) * p.n_segments + item) * kListOut;
        for (int i = lane; i < kListOut; i += 32) out[i] = lst[i];
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(acc_base, kTmemCols);
#ifdef HRC_EXPERIMENTS
  if (p.cta_times != nullptr && threadIdx.x == 0) {        // when did this CTA finish, and where did it run
    unsigned long long now;
    unsigned smid;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    p.cta_times[2 * blockIdx.x] = now;
    p.cta_times[2 * blockIdx.x + 1] = smid;
  }
#endif
}

// --- host side -------------------------------------------------------------------------------

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess || sym == nullptr) {
    return nullptr;
  }
  fn = reinterpret_cast<EncodeTiledFn>(sym);
  return fn;
}

// Tensor maps are a pure function of (base address, extents, box), never of the data, so encoded maps are kept in
// a small process-wide cache: after the first call on a store (or query buffer) no cuTensorMapEncodeTiled runs on
// the launch path.  hrc_store_register pre-encodes a store's maps; hrc_store_release drops them.
struct MapEntry {
  const void* base = nullptr;
  uint64_t d1 = 0, d2 = 0;     // tokens: (total_tokens, 0); queries: (lq, n_queries)
  uint32_t box_rows = 0;
  CUtensorMap map;
};
constexpr int kMapCache = 64;
std::mutex g_map_mu;
MapEntry g_maps[kMapCache];
unsigned g_map_next = 0;

int cached_map(const void* base, uint64_t d1, uint64_t d2, uint32_t box_rows, CUtensorMap* out) {
  {
    std::lock_guard<std::mutex> lk(g_map_mu);
    for (int i = 0; i < kMapCache; ++i) {
      const MapEntry& e = g_maps[i];
      if (e.base == base && e.d1 == d1 && e.d2 == d2 && e.box_rows == box_rows) { *out = e.map; return 0; }
    }
  }
  EncodeTiledFn encode = get_encode_fn();
  HRC_REQUIRE(encode != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  CUtensorMap m;
  CUresult r;
  if (d2 == 0) {   // document tokens: 2-D {128, total_tokens}, box {64, box_rows}
    cuuint64_t dims[2] = {HRC_DIM, (cuuint64_t)d1};
    cuuint64_t strides[1] = {HRC_DIM * 2};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    r = encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {         // queries: 3-D {128, lq, n_queries}, box {64, 32, 1}; out-of-bounds rows / queries read as zeros
    cuuint64_t dims[3] = {HRC_DIM, (cuuint64_t)d1, (cuuint64_t)d2};
    cuuint64_t strides[2] = {HRC_DIM * 2, (cuuint64_t)d1 * HRC_DIM * 2};
    cuuint32_t box[3] = {64, box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    r = encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  HRC_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed: %d", int(r));
  {
    std::lock_guard<std::mutex> lk(g_map_mu);
    MapEntry& e = g_maps[g_map_next++ % kMapCache];
    e.base = base; e.d1 = d1; e.d2 = d2; e.box_rows = box_rows; e.map = m;
  }
  *out = m;
  return 0;
}

#ifdef HRC_EXPERIMENTS
unsigned long long* g_exp_cta_times = nullptr;   // hrc_exp_set_cta_times: device buffer the doc-major kernel's CTAs stamp at exit
int g_exp_dyn_share = 0, g_exp_dyn_per_cta = 0;  // hrc_exp_set_dyn: 1/share of the corpus in per_cta shared units per CTA (0 = default)
int g_exp_ctas = 0;                               // hrc_exp_set_ctas: CTAs (corpus segments) of the persistent kernels, 0 = one per SM
#endif

int sm_count() {
#ifdef HRC_EXPERIMENTS
  if (g_exp_ctas > 0) return g_exp_ctas;
#endif
  static int cached[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && cached[dev] > 0) return cached[dev];
  int n = 0;
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  if (n <= 0) n = 148;
  if (dev >= 0 && dev < 64) cached[dev] = n;
  return n;
}

uint64_t g_watchdog_ns = 20ull * 1000000000ull;   // hrc_set_watchdog_ms
#ifdef HRC_EXPERIMENTS
int g_debug = 0;                                  // hrc_exp_set_debug: 1 no epilogue math, 2 no document TMA, 4 no MMA,
                                                  // 8 doc-major fused top-k never offers, 32 ... skips its final list merge
int g_stages = 0;                                 // hrc_exp_set_stages: cap of the shared-memory ring depth (0 = default)
#endif

template <int MT, int ZP, int CG, bool TK = false, bool RR = false>
int launch_cfg(const void* d_tokens, const void* d_queries, int lq, int n_real_queries, TcParams p, dim3 grid,
               cudaStream_t stream) {
  constexpr int kTileBytes = (TN / CG) * HRC_DIM * 2;   // what ONE CTA stages per tile
  constexpr int kQBytes = 128 * HRC_DIM * 2;
  CUtensorMap tmap_d, tmap_q;
  if (int rc = cached_map(d_tokens, uint64_t(p.total_tokens), 0, TN / CG, &tmap_d)) return rc;
  if (int rc = cached_map(d_queries, uint64_t(lq), uint64_t(n_real_queries), 32, &tmap_q)) return rc;
  const int q_bytes = MT * kQBytes;
  const int list_bytes = TK ? epi_warps(MT) * kListCap * 8 : 0;
  int stages = (kMaxSmem - 1024 - 512 - q_bytes - list_bytes) / kTileBytes;
  // Ring depth: 8 x 16 KB for the batched kernels; FIVE x 32 KB for the HBM-bound ones — a sixth stage fits but is
  // slower (libhrc_exp stage sweep, same box: C2 4.67-4.90 ms with 6, 4.69-4.78 with 5, 4.60-4.64 with 4, 5.09 with 3;
  // ragged single query 10.63 / 9.89 / 9.84 / 10.41): 160 KB in flight per SM already covers the latency-bandwidth
  // product several times, and more outstanding requests per SM only spread the DRAM access pattern.
  if (stages > (MT == 1 ? 5 : 8)) stages = MT == 1 ? 5 : 8;
#ifdef HRC_EXPERIMENTS
  if (g_stages > 0 && stages > g_stages) stages = g_stages;
#endif
  p.n_stages = stages;
  const int smem_bytes = 1024 + q_bytes + stages * kTileBytes + 512 + list_bytes;
  static PerDeviceOnce once;
  int dev;
  if (once.pending(&dev)) {
    HRC_CHECK_CUDA(cudaFuncSetAttribute(maxsim_tc_kernel<MT, ZP, CG, TK, RR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        kMaxSmem));
    once.mark(dev);
  }
  trace_begin(stream);
  if constexpr (CG == 2) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;                                  // grid.x is even: consecutive CTAs form a pair (one TPC)
    cfg.blockDim = dim3(cta_threads(MT, ZP));
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    HRC_CHECK_CUDA(cudaLaunchKernelEx(&cfg, maxsim_tc_kernel<MT, ZP, CG, TK, RR>, tmap_d, tmap_q, p));
  } else {
    maxsim_tc_kernel<MT, ZP, CG, TK, RR><<<grid, cta_threads(MT, ZP), smem_bytes, stream>>>(tmap_d, tmap_q, p);
  }
  trace_end(stream);
  count_launch();
  HRC_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <bool TK>
int launch_dm(const void* d_tokens, const void* d_queries, int lq, int n_real_queries, TcParams p, cudaStream_t stream) {
  CUtensorMap tmap_d, tmap_q;
  if (int rc = cached_map(d_tokens, uint64_t(p.total_tokens), 0, 32, &tmap_d)) return rc;     // one stream's 32-token box
  if (int rc = cached_map(d_queries, uint64_t(lq), uint64_t(n_real_queries), 32, &tmap_q)) return rc;
  p.n_stages = kDmStages;
  // work units: one own unit per CTA over the first 3/4 of the tokens + 16 shared units per CTA over the rest, when the
  // caller's workspace holds the claim counter and a CTA's share is long enough for the SMs' speed spread to matter
  p.n_units = p.n_segments;
  p.static_tokens = p.total_tokens;
  int share = kDmSharedShare, per_cta = kDmSharedPerCta;
  // A unit is cut into four streams of WHOLE documents: a shared unit must hold >= 16 average documents or its streams
  // are badly balanced, and there must be >= 8 shared units per CTA or the last one to be claimed IS the new tail
  // (32k documents of 4,000 tokens: one 27-document shared unit per CTA made the kernel 8 % slower than equal
  // ranges, profiles/r02_logs/r02_ab_dynamic_units_long.log) — long documents keep the static distribution.
  const int64_t mean_len = p.total_tokens / (p.n_docs > 0 ? p.n_docs : 1) + 1;
  const int64_t fit = (p.total_tokens / share / p.n_segments) / (16 * mean_len);
  if (fit < per_cta) per_cta = int(fit);
  bool dynamic = p.unit_counter != nullptr && p.total_tokens >= int64_t(p.n_segments) * kDmDynMinTokensPerCta && per_cta >= 8;
#ifdef HRC_EXPERIMENTS
  if (g_exp_dyn_share > 0 && g_exp_dyn_per_cta > 0 && p.unit_counter != nullptr) {
    share = g_exp_dyn_share;
    per_cta = g_exp_dyn_per_cta;
    dynamic = true;
  }
#endif
  if (dynamic) {
    p.n_units = p.n_segments * (1 + per_cta);
    p.static_tokens = p.total_tokens - p.total_tokens / share;
    HRC_CHECK_CUDA(cudaMemsetAsync(p.unit_counter, 0, sizeof(uint32_t), stream));
  } else {
    p.unit_counter = nullptr;
  }
  const int smem_bytes = 1024 + kDmQBytes + kDmStages * kDmTileBytes + kDmBarBytes + (TK ? 4 * kListCap * 8 : 0);
  static PerDeviceOnce once;
  int dev;
  if (once.pending(&dev)) {
    HRC_CHECK_CUDA(cudaFuncSetAttribute(maxsim_dm_kernel<TK>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    once.mark(dev);
  }
  trace_begin(stream);
  maxsim_dm_kernel<TK><<<dim3((unsigned)p.n_segments), kDmThreads, smem_bytes, stream>>>(tmap_d, tmap_q, p);
  trace_end(stream);
  count_launch();
  HRC_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// out[q][i] = sum over the slots of query q, in slot order (deterministic)
__global__ void sum_slots_kernel(const float* __restrict__ part, int q_slots, int64_t n_items, int64_t total,
                                 float* __restrict__ out) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int64_t q = i / n_items, d = i - q * n_items;
  float acc = 0.f;
  for (int sl = 0; sl < q_slots; ++sl) acc += part[(q * q_slots + sl) * n_items + d];
  out[i] = acc;
}

// single-query kernel organisations.  Auto (HRC_PATH_AUTO): doc-major for one query of <= 32 tokens over the corpus —
// same time as the query-major kernel at burst clocks, 4-7 % faster than round 1's organisation once the GPU sits at its
// power cap (3 s back to back, same box: 4.41-4.77 vs 4.59-5.10 ms per C2 scan) — query-major for everything else.
constexpr int kVariantDefault = 0, kVariantDocMajor = 2, kVariantAuto = 3;

struct TopkOut {            // fused top-k request (corpus mode, lq <= 32, k <= kListOut)
  uint64_t* cand_keys;
  int k;
  int32_t id_base;
};
struct RerankOut {          // fused rerank request (candidate mode, lq <= 32, n_cand <= kRerankFusedMax)
  uint32_t* counter;
  int k;
  int32_t* pos;
  int32_t* ids;
  float* scores;
};
constexpr int kRerankFusedMax = 1024;

int launch_tc_slots(const void* d_tokens, const int64_t* d_offsets, int64_t n_docs, int64_t total_tokens,
                    const int32_t* d_cand_ids, int64_t n_items, const void* d_queries, int n_real_queries,
                    int q_slots, int lq, float* d_scores, int variant, const TopkOut* tk, const RerankOut* rr,
                    uint32_t* unit_counter, cudaStream_t stream) {
  const bool dm = variant == kVariantDocMajor || (variant == kVariantAuto && d_cand_ids == nullptr && q_slots == 1);
  const int n_queries = n_real_queries * q_slots;      // virtual queries from here on
  HRC_REQUIRE(total_tokens > 0 && total_tokens < (1ll << 31), "tc path: total_tokens=%lld out of range",
              (long long)total_tokens);
  HRC_REQUIRE((reinterpret_cast<uintptr_t>(d_tokens) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_queries) & 15) == 0,
              "tc path: token / query buffers must be 16-byte aligned");

  TcParams p;
  p.offsets = d_offsets;
  p.cand_ids = d_cand_ids;
  p.scores = d_scores;
  p.cand_keys = tk ? tk->cand_keys : nullptr;
  p.k = tk ? tk->k : 0;
  p.id_base = tk ? tk->id_base : 0;
  p.rr_counter = rr ? rr->counter : nullptr;
  p.rr_k = rr ? rr->k : 0;
  p.rr_pos = rr ? rr->pos : nullptr;
  p.rr_ids = rr ? rr->ids : nullptr;
  p.rr_scores = rr ? rr->scores : nullptr;
  p.n_docs = n_docs;
  p.total_tokens = total_tokens;
  p.n_items = n_items;
  p.n_queries = n_queries;
  p.q_slots = q_slots;
  p.vq_base = 0;
  p.n_segments = 1;
  p.n_qgroups = 1;
  p.n_stages = 0;
  p.slots_used = 1;
  p.n_units = 1;
  p.static_tokens = total_tokens;
  p.unit_counter = unit_counter;
  p.debug = 0;
  p.cta_times = nullptr;
#ifdef HRC_EXPERIMENTS
  p.debug = g_debug;
  p.cta_times = g_exp_cta_times;
#endif
  p.watchdog_ns = g_watchdog_ns;
  p.doc_policy = kEvictFirst;

  if (d_cand_ids != nullptr) {
    HRC_REQUIRE(n_queries <= 65535, "tc path: too many queries for a candidate launch (%d)", n_queries);
    const dim3 cgrid((unsigned)n_items, (unsigned)n_queries);
    return rr ? launch_cfg<1, 1, 1, false, true>(d_tokens, d_queries, lq, n_real_queries, p, cgrid, stream)
              : launch_cfg<1, 1, 1>(d_tokens, d_queries, lq, n_real_queries, p, cgrid, stream);
  }
  const int64_t tiles = (total_tokens + TN - 1) / TN;
  p.n_segments = int(tiles < sm_count() ? tiles : sm_count());
  if (dm && n_queries == 1) {                   // doc-major orientation: exactly the useful tensor work
    // every stream of a CTA wants at least a chunk or two: 4 streams per segment
    const int64_t chunks = (total_tokens + 127) / 128;
    p.n_segments = int(chunks < sm_count() ? chunks : sm_count());
    return tk ? launch_dm<true>(d_tokens, d_queries, lq, n_real_queries, p, stream)
              : launch_dm<false>(d_tokens, d_queries, lq, n_real_queries, p, stream);
  }
  if (n_queries <= 4) {
    p.slots_used = n_queries == 1 ? 1 : (n_queries == 2 ? 2 : 4);
    if (tk) return launch_cfg<1, 1, 1, true>(d_tokens, d_queries, lq, n_real_queries, p, dim3((unsigned)p.n_segments), stream);
    return launch_cfg<1, 1, 1>(d_tokens, d_queries, lq, n_real_queries, p, dim3((unsigned)p.n_segments), stream);
  }
  p.n_qgroups = (n_queries + 7) / 8;
  p.slots_used = 4;
  p.doc_policy = kEvictNormal;  // the other query groups re-read this tile from L2
  // CTA pairs (cta_group::2) from 2 query groups up: two query groups of the same corpus segment share every
  // document tile — each CTA stages half of it — so the L2 -> shared-memory traffic and the B-operand reads per SM
  // halve.  C3 (256 queries, 1M ragged documents, power-capped): 1233 vs 1186 TFLOP/s at the time of the switch.
  // An odd last query group runs on the single-CTA kernel.
  if (p.n_qgroups >= 2) {
    const int paired = p.n_qgroups & ~1;                // query groups handled by pairs
    TcParams pp = p;
    pp.n_qgroups = paired;
    pp.n_queries = n_queries < paired * 8 ? n_queries : paired * 8;
    const dim3 pgrid((unsigned)(pp.n_segments * paired));
    int rc = launch_cfg<2, 0, 2>(d_tokens, d_queries, lq, n_real_queries, pp, pgrid, stream);
    if (rc != 0 || paired == p.n_qgroups) return rc;
    TcParams pl = p;                                    // the odd group: (virtual) queries [paired * 8, n_queries)
    pl.n_qgroups = 1;
    pl.vq_base = paired * 8;
    return launch_cfg<2, 0, 1>(d_tokens, d_queries, lq, n_real_queries, pl, dim3((unsigned)pl.n_segments), stream);
  }
  return launch_cfg<2, 0, 1>(d_tokens, d_queries, lq, n_real_queries, p, dim3((unsigned)(p.n_segments * p.n_qgroups)), stream);
}

}  // namespace

void set_watchdog_ns(uint64_t ns) { g_watchdog_ns = ns; }
uint64_t get_watchdog_ns() { return g_watchdog_ns; }
#ifdef HRC_EXPERIMENTS
void set_debug(int bits) { g_debug = bits; }
void set_stages(int n) { g_stages = n; }
void set_ctas(int n) { g_exp_ctas = n; }
void set_dyn(int share, int per_cta) { g_exp_dyn_share = share; g_exp_dyn_per_cta = per_cta; }
void set_cta_times(unsigned long long* d) { g_exp_cta_times = d; }
#endif

int store_register(const void* d_tokens, int64_t total_tokens) {
  HRC_REQUIRE(d_tokens != nullptr && total_tokens > 0 && total_tokens < (1ll << 31),
              "store_register: bad store (total_tokens=%lld)", (long long)total_tokens);
  CUtensorMap m;
  if (int rc = cached_map(d_tokens, uint64_t(total_tokens), 0, TN, &m)) return rc;
  return cached_map(d_tokens, uint64_t(total_tokens), 0, TN / 2, &m);
}

void store_release(const void* base) {
  std::lock_guard<std::mutex> lk(g_map_mu);
  for (int i = 0; i < kMapCache; ++i)
    if (g_maps[i].base == base) g_maps[i] = MapEntry();
}

// ---- fused MaxSim + per-segment top-k (hrc_search's default) -------------------------------------------------------
// Fused top-k is used for ONE query (the HBM-bound kernel, whose epilogue has slack).  Measured in-process on one box,
// fused vs score matrix + top-k: C2 one query 4.59-4.69 vs 4.85-5.04 ms, ragged one query 9.6-10.0 vs 10.7-11.0 ms; two
// and four queries 2 % SLOWER.  In the tensor-bound batched kernels the epilogue IS the critical resource: even a
// never-taken append costs 2.2 % (registers, code size) and the real thing 5 % (C3 64 queries 109.8 vs 104.4 ms) against
// a top-k pass that costs 1 %, so everything but the single-query search keeps writing scores and runs the streaming
// top-k of topk.cu (profiles/r02_summary.md, "fused top-k").
bool tc_topk_supported(int64_t total_tokens, int n_queries, int lq, int k) {
  return total_tokens > 0 && total_tokens < (1ll << 31) && n_queries >= 1 && n_queries <= HRC_FUSED_TOPK_MAX_QUERIES &&
         lq >= 1 && lq <= HRC_TC_MAX_LQ && k >= 1 && k <= kListOut;
}
int tc_topk_segments(int64_t total_tokens) {     // CTAs along the corpus = key lists per query
  const int64_t tiles = (total_tokens + TN - 1) / TN;
  return int(tiles < sm_count() ? tiles : sm_count());
}
int tc_topk_list_len() { return kListOut; }
// d_cand_keys: uint64 [n_queries][tc_topk_segments()][kListOut]; d_scores optional (the full matrix, if wanted)
int launch_maxsim_tc_topk(const void* d_tokens, const int64_t* d_offsets, int64_t n_docs, int64_t total_tokens,
                          const void* d_queries, int n_queries, int lq, int k, int32_t id_base, float* d_scores,
                          uint64_t* d_cand_keys, int variant, uint32_t* d_unit_counter, cudaStream_t stream) {
  if (n_docs == 0 || n_queries == 0) return 0;
  HRC_REQUIRE(tc_topk_supported(total_tokens, n_queries, lq, k), "fused top-k: needs <= %d queries, lq <= %d and k <= %d",
              HRC_FUSED_TOPK_MAX_QUERIES, HRC_TC_MAX_LQ, kListOut);
  HRC_REQUIRE(d_cand_keys != nullptr, "fused top-k: null candidate buffer");
  const TopkOut tk{d_cand_keys, k, id_base};
  return launch_tc_slots(d_tokens, d_offsets, n_docs, total_tokens, nullptr, n_docs, d_queries, n_queries, 1, lq, d_scores,
                         variant, &tk, nullptr, d_unit_counter, stream);
}

// ---- fused rerank: candidate MaxSim + sorted top-k in ONE launch (hrc_rerank's default) ----------------------------
bool tc_rerank_supported(int64_t total_tokens, int lq, int n_cand, int k) {
  return total_tokens > 0 && total_tokens < (1ll << 31) && lq >= 1 && lq <= HRC_TC_MAX_LQ && n_cand >= 1 &&
         n_cand <= kRerankFusedMax && k >= 1 && k <= n_cand;
}
// d_scores: fp32 [n_queries][n_cand] (every candidate's score); d_counter: uint32 [n_queries], zeroed here
int launch_maxsim_tc_rerank(const void* d_tokens, const int64_t* d_offsets, int64_t n_docs, int64_t total_tokens,
                            const int32_t* d_cand_ids, int n_cand, const void* d_queries, int n_queries, int lq, int k,
                            float* d_scores, uint32_t* d_counter, int32_t* d_pos, int32_t* d_ids, float* d_scores_out,
                            cudaStream_t stream) {
  if (n_queries == 0) return 0;
  HRC_REQUIRE(tc_rerank_supported(total_tokens, lq, n_cand, k), "fused rerank: needs lq <= %d, n_cand <= %d", HRC_TC_MAX_LQ,
              kRerankFusedMax);
  HRC_REQUIRE(d_scores != nullptr && d_counter != nullptr && d_pos != nullptr && d_scores_out != nullptr,
              "fused rerank: null buffer");
  HRC_CHECK_CUDA(cudaMemsetAsync(d_counter, 0, size_t(n_queries) * sizeof(uint32_t), stream));
  const RerankOut rr{d_counter, k, d_pos, d_ids, d_scores_out};
  return launch_tc_slots(d_tokens, d_offsets, n_docs, total_tokens, d_cand_ids, n_cand, d_queries, n_queries, 1, lq,
                         d_scores, kVariantDefault, nullptr, &rr, nullptr, stream);
}

// bytes of caller workspace the tensor-core path needs: the per-slot partial scores of queries longer than 32 tokens
size_t maxsim_tc_workspace_bytes(int64_t n_items, int n_queries, int lq) {
  const int q_slots = (lq + HRC_TC_MAX_LQ - 1) / HRC_TC_MAX_LQ;
  if (q_slots == 1 && n_queries == 1) return 256;      // the doc-major kernel's claim counter (optional: see launch_maxsim_tc)
  if (q_slots <= 1 || q_slots > HRC_TC_MAX_SLOTS) return 0;
  return size_t(n_queries) * size_t(q_slots) * size_t(n_items) * sizeof(float);
}

int launch_maxsim_tc(const void* d_tokens, const int64_t* d_offsets, int64_t n_docs, int64_t total_tokens,
                     const int32_t* d_cand_ids, int64_t n_items, const void* d_queries, int n_queries,
                     int lq, float* d_scores, int variant, void* d_workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (n_items == 0 || n_queries == 0) return 0;
  HRC_REQUIRE(lq >= 1 && lq <= HRC_TC_MAX_LQ * HRC_TC_MAX_SLOTS, "tc path: lq=%d not in [1,%d]", lq,
              HRC_TC_MAX_LQ * HRC_TC_MAX_SLOTS);
  const int q_slots = (lq + HRC_TC_MAX_LQ - 1) / HRC_TC_MAX_LQ;
  if (q_slots == 1) {
    // (one query over the corpus: the doc-major kernel hands part of the work out dynamically when the caller's
    // workspace can hold its 4-byte claim counter; without one it falls back to equal shares)
    uint32_t* counter = (d_cand_ids == nullptr && n_queries == 1 && d_workspace != nullptr && workspace_bytes >= sizeof(uint32_t) &&
                         (reinterpret_cast<uintptr_t>(d_workspace) & 3) == 0)
                            ? static_cast<uint32_t*>(d_workspace) : nullptr;
    return launch_tc_slots(d_tokens, d_offsets, n_docs, total_tokens, d_cand_ids, n_items, d_queries, n_queries, 1, lq,
                           d_scores, variant, nullptr, nullptr, counter, stream);
  }
  // A query of more than 32 tokens is scored as q_slots virtual queries of <= 32 tokens (rows beyond lq arrive as
  // zeros from TMA and add max_t <0, d_t> = 0); their partial scores (caller workspace) are summed in slot order.
  HRC_REQUIRE(int64_t(n_queries) * q_slots <= 65535, "tc path: too many query slots (%d x %d)", n_queries, q_slots);
  const size_t need = maxsim_tc_workspace_bytes(n_items, n_queries, lq);
  HRC_REQUIRE(d_workspace != nullptr && workspace_bytes >= need,
              "tc path: lq=%d > 32 needs %zu bytes of workspace (hrc_maxsim_workspace_bytes), got %zu", lq, need,
              workspace_bytes);
  float* part = static_cast<float*>(d_workspace);
  const int64_t total = int64_t(n_queries) * n_items;
  if (int rc = launch_tc_slots(d_tokens, d_offsets, n_docs, total_tokens, d_cand_ids, n_items, d_queries, n_queries,
                               q_slots, lq, part, variant, nullptr, nullptr, nullptr, stream))
    return rc;
  sum_slots_kernel<<<unsigned((total + 255) / 256), 256, 0, stream>>>(part, q_slots, n_items, total, d_scores);
  count_launch();
  HRC_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace hrc
