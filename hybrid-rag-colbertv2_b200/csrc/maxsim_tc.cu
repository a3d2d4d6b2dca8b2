// maxsim_tc.cu — tensor-core MaxSim: TMA -> tcgen05.mma (TMEM accumulators) -> fused segmented
// row-max / query-token-sum epilogue.  The [Lq x Ld] similarity matrix never leaves the SM.
//
// Data layout in HBM
//   tokens  : bf16 [total_tokens][128], packed, padding-free           (TMA map: 2-D, 128B swizzle)
//   offsets : int64 [n_docs + 1] CSR document boundaries
//   queries : bf16 [n_queries][lq][128]; 32 query tokens per A-tile slot, a longer query (lq <= 256) is scored as
//             ceil(lq / 32) "virtual queries" whose partial scores are summed in slot order (sum_slots_kernel)
//
// One CTA = one contiguous run of whole documents ("segment") x one group of 4*MT (virtual) queries.
//   warp 0      : TMA producer  — streams TN-token x 128-dim tiles through a shared-memory ring
//   warp 1      : MMA issuer    — per tile and M-tile: 8 x tcgen05.mma (M=128 [256 over a CTA pair], N=TN, K=16),
//                                 accumulator = TN TMEM columns; owns TMEM alloc/dealloc
//   warps 2..   : epilogue      — a warp owns TMEM lanes 32*(w%4).. = one query slot: thread i holds query
//                                 token i, so the max over a document's tokens is a per-thread FMNMX3 tree over
//                                 32 TMEM columns (loads aligned to the DOCUMENT, not the tile: no masking for
//                                 >= 32-token pieces), document boundaries are warp-uniform, and a document's
//                                 score is one shuffle butterfly, emitted after the tile has been released.
//                                 Every warp owns WHOLE documents (no cross-warp combine).
// MMA orientation: A = queries (M = 4 slots x 32 tokens), B = document tokens (N):
// D[row = query token][col = doc token].
//
// Instantiations <MT, TN, TS, ZP, CG, EPI> (DESIGN.md §4.1 has the measurements behind each choice):
//   <1,128,SS,ZP=1>       HBM-bound (<= 4 queries) and candidate (rerank) mode: the unused rows of the A tile are
//                         zero, the 4 epilogue warps are stacked on the used lane groups, 6-stage smem ring,
//                         4 accumulator stages, one tcgen05.commit per tile.            [default]
//   <1,128,SS,ZP=2>       M=64 MMA for <= 2 queries (env HRC_TC_M64=1); <1,128,SS,ZP=0> replicated query (HRC_TC_ZP=0)
//   <2,128,SS,CG=2>       tensor-bound, batched: CTA PAIR (cta_group::2, cluster of 2), one M=256 MMA per K slice,
//                         each CTA stages half of every document tile; 8 epilogue warps per CTA (two per lane
//                         group, alternating documents), 2 x 2 accumulators.            [default from 9 queries]
//   <2,128,SS>            single-CTA batched kernel (5..8 queries, an odd last query group, or HRC_TC_PAIR=0)
//   <2,128,SS,CG=2,EPI=1..3>  other organisations of the batched epilogue (env HRC_TC_EPI), parity-tested, slower
//   <2,96,TS>             query tiles in TMEM as the A operand (env HRC_TC_TS=1), parity-tested, slower
//
// Reference semantics: local_rag_complete.py:807-812 (docstring), :813-817 (shapes), summed over
// query tokens per BASELINE.json north_star.  Algorithmic traffic: 256 B per document token.
#include <cuda.h>
#include <climits>
#include <cstdio>
#include <cstdlib>

#include "hrc_common.cuh"

namespace hrc {

namespace {

constexpr int kQTileBytes = 128 * HRC_DIM * 2;    // 128 query rows x 128 dims, 32 KB (SS mode only)
constexpr int kSlotBytes = 32 * 128;              // one 32-row query slot inside a 64-dim slab
constexpr int kTmemCols = 512;
constexpr int kEpiWarp0 = 2;
constexpr int kMaxSmem = 232448;                  // 227 KB opt-in limit per CTA

__host__ __device__ constexpr int epi_warps(int mt) { return mt == 1 ? 4 : 8; }
// ZP ("zero padded") variant for <= 4 queries: with slots_used = 1, 2 or 4 queries the other rows of the A tile
// stay ZERO (measured: replicating the query instead costs ~10 % under sustained load, because the extra
// tensor-core switching power pushes the GPU into its 1 kW cap), and the 4 epilogue warps are stacked on the
// USED TMEM lane groups: a warp can only read lanes 32*(warp%4).., so with one query they are warps 4, 8, 12, 16
// (all on lane group 0, splitting the documents 4 ways), with two queries warps 4, 5, 8, 9, with four 4..7.
// The warps in between have no role and exit.
// ZP == 2 additionally uses an M=64 MMA for <= 2 queries (half the tensor work and A-operand traffic).  Its
// accumulator layout puts rows 0-15 / 16-31 / 32-47 / 48-63 in lanes 0-15 of lane groups 0 / 1 / 2 / 3, so a
// query's tokens 0-15 and 16-31 are summed by two different warps, which each atomicAdd their half into the
// (zero-initialised) score: two commutative additions, hence still deterministic.  8 epilogue warps.
__host__ __device__ constexpr int cta_threads(int mt, int zp = 0) {
  return zp == 2 ? 18 * 32 : (zp == 1 ? 17 * 32 : (2 + epi_warps(mt)) * 32);
}

struct TcParams {
  const int64_t* offsets;
  const int32_t* cand_ids;  // nullptr: corpus mode
  const __nv_bfloat16* queries;
  float* scores;
  int64_t n_docs;
  int64_t total_tokens;
  int64_t n_items;          // row stride of scores (n_docs, or n_cand)
  int n_queries;            // (virtual) queries: one per 32-token slot of a real query
  int n_real_queries;       // rows of the query tensor
  int q_slots;              // 32-token slots per real query: ceil(lq / 32); virtual query v = real v / q_slots, slot v % q_slots
  int vq_base;              // first virtual query of this launch
  int lq;
  int n_segments;           // corpus mode: CTAs along the corpus
  int n_qgroups;            // corpus mode: query groups (4*MT queries each)
  int n_stages;             // smem ring depth
  int slots_used;           // distinct queries per A tile: 1, 2 or 4 (each replicated 4/slots_used times)
  int debug;                // perf experiments only (env HRC_TC_DEBUG): 1 = skip epilogue math, 2 = skip TMA of documents, 4 = skip MMA
  uint64_t doc_policy;      // L2 policy for document tiles (evict-first when read once)
};

// Spin on an mbarrier with a wall-clock watchdog: a protocol bug must fault, not hang the GPU.
__device__ __forceinline__ void mbar_wait_wd(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  uint64_t t0 = 0;
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3ffu) == 0) {
      uint64_t now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      if (now - t0 > 4000000000ull) {
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

// D[tmem] (+)= A[tmem] * B[smem]^T : the A operand (queries) is read from tensor memory.
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 lanes x 64 consecutive 32-bit columns, registers -> TMEM: thread i writes lane (base_lane + i).
__device__ __forceinline__ void tmem_st_32x64(uint32_t taddr, const uint32_t (&v)[64]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x64.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, "
      "%33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, "
      "%49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63, %64};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
      "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
      "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]), "r"(v[32]), "r"(v[33]), "r"(v[34]), "r"(v[35]), "r"(v[36]),
      "r"(v[37]), "r"(v[38]), "r"(v[39]), "r"(v[40]), "r"(v[41]), "r"(v[42]), "r"(v[43]), "r"(v[44]), "r"(v[45]),
      "r"(v[46]), "r"(v[47]), "r"(v[48]), "r"(v[49]), "r"(v[50]), "r"(v[51]), "r"(v[52]), "r"(v[53]), "r"(v[54]),
      "r"(v[55]), "r"(v[56]), "r"(v[57]), "r"(v[58]), "r"(v[59]), "r"(v[60]), "r"(v[61]), "r"(v[62]), "r"(v[63])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

template <int MT, int TN, bool TS, int ZP = 0, int CG = 1, int EPI = 0>
__global__ void __launch_bounds__(cta_threads(MT, ZP), 1)
maxsim_tc_kernel(const __grid_constant__ CUtensorMap tmap_d, const __grid_constant__ CUtensorMap tmap_q,
                 const TcParams p) {
  constexpr int kEpiWarps = ZP == 2 ? 8 : epi_warps(MT);
  constexpr int kM = ZP == 2 ? 64 : 128;               // MMA M (rows of the A tile)
  constexpr int kQBytes = kM * HRC_DIM * 2;            // one A tile in shared memory (SS mode)
  constexpr int kSplit = kEpiWarps / 4;                 // warps sharing one TMEM lane group split the documents
  // CG == 2 (CTA pair, cta_group::2): the pair issues one M=256 MMA per K slice; this CTA stages TN/2 tokens
  // of every tile (its half of the B operand), its own MT query tiles and its own accumulators.
  constexpr int kTileRows = TN / CG;                    // document tokens of a tile staged by THIS CTA
  constexpr int kTileBytes = kTileRows * HRC_DIM * 2;
  constexpr int kHalfTileBytes = kTileBytes / 2;        // one 64-dim (128-byte-row) slab
  constexpr int kQCols = TS ? MT * 64 : 0;              // TMEM columns holding the query tiles (bf16 pairs)
  constexpr int kTileStages = (kTmemCols - kQCols) / (MT * TN);   // tiles in flight between MMA and epilogue
  constexpr uint32_t kIdesc = make_idesc_bf16_f32(kM * CG, TN);
  static_assert(TN % 32 == 0 && TN % 16 == 0 && TN <= 256 && kTileStages >= 2, "bad tile configuration");
  static_assert(kTileBytes % 2048 == 0, "tile slabs must stay 1024-byte aligned for the 128B swizzle");
  static_assert(ZP == 0 || (MT == 1 && !TS), "ZP is a few-query variant");
  static_assert(CG == 1 || (CG == 2 && MT == 2 && !TS && ZP == 0), "CTA pairs: batched SS kernel only");
  // How the epilogue warps of a lane group share the work (EPI):
  //  0  the warps alternate DOCUMENTS and each reads all M-tiles of a tile in one walk; a tile's accumulators are one
  //     unit (one tfull / tempty pair per stage).  Default.
  //  1  the warps take one M-TILE each and both walk every document; every (stage, M-tile) is its own unit.
  //  2  the warps alternate documents, and each walks a tile once PER M-TILE, releasing M-tile 0 before it reads
  //     M-tile 1; every (stage, M-tile) is its own unit, so 2 x MT hand-shakes are in flight.
  //  3  as 0, but the MMA issuer publishes every M-TILE of a tile separately (one tfull per (stage, M-tile), still
  //     one tempty per stage) and a warp reads a document's columns M-tile by M-tile: it works on M-tile 0 while
  //     the MMAs of M-tile 1 are still executing, which takes half an epilogue off the hand-shake chain.
  static_assert(EPI == 0 || (MT == 2 && ZP == 0), "bad EPI");
  constexpr int MTW = (EPI == 0 || EPI == 3) ? MT : 1;   // M-tiles read in one walk
  constexpr int kFullPerStage = EPI == 0 ? 1 : MT;       // tfull barriers per accumulator stage
  constexpr int kUnitsPerStage = (EPI == 0 || EPI == 3) ? 1 : MT;   // tempty barriers per accumulator stage
  constexpr int kUnits = kTileStages * kUnitsPerStage;
  constexpr int kReadersPerUnit = EPI == 1 ? epi_warps(MT) / MT : (ZP == 2 ? 8 : epi_warps(MT));
  // HBM-bound kernels (MT == 1): ONE tcgen05.commit per tile (tfull); the shared-memory slot is released by the
  // first epilogue warp when it sees tfull (the same event, ~100 cycles later, irrelevant with a 6-deep TMA ring).
  // In-process A/B: C2 4.79 -> 4.70 ms, ragged 10.32 -> 10.07 ms.  The batched kernels keep the second commit
  // (1,349 vs 1,332 TFLOP/s).
  constexpr bool kForwardEmpty = MT == 1;
  static_assert(kUnits <= 4, "tfull / tempty hold 4 barriers each");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                        // SS: MT x 32 KB query tiles
  uint8_t* sD = smem + (TS ? 0 : MT * kQBytes);              // n_stages x tile
  uint64_t* bars = reinterpret_cast<uint64_t*>(sD + p.n_stages * kTileBytes);
  uint64_t* full = bars;                           // [n_stages]    TMA -> MMA
  uint64_t* empty = bars + 10;                     // [n_stages]    MMA -> TMA
  uint64_t* tfull = bars + 20;                     // [kTileStages] MMA -> epilogue
  uint64_t* tempty = bars + 24;                    // [kTileStages] epilogue -> MMA
  uint64_t* qfull = bars + 28;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 29);
  int64_t* seg = reinterpret_cast<int64_t*>(bars + 30);  // [0]=doc_begin [1]=doc_end [2]=tok_begin [3]=tok_end

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = CG == 2 ? cluster_ctarank() : 0u;   // 0 = leader of the pair (issues the MMAs)

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
    if (!TS) tma_prefetch_desc(&tmap_q);
    for (int i = 0; i < p.n_stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < kTileStages * kFullPerStage; ++i) mbar_init(&tfull[i], 1);
    for (int i = 0; i < kUnits; ++i) mbar_init(&tempty[i], kReadersPerUnit * CG);
    mbar_init(qfull, TS ? (EPI == 1 ? kEpiWarps : 4) : 1);
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

  const uint32_t tmem_base = *tmem_slot;
  const uint32_t acc_base = tmem_base + kQCols;
  const int64_t doc_begin = seg[0], doc_end = seg[1], tok_begin = seg[2], tok_end = seg[3];
  const int n_tiles = int((tok_end - tok_begin + TN - 1) / TN);

  if (warp == 0) {
    // =============================== TMA producer =============================================
    // (elect.sync, not `lane == 0`: the compiler then knows a single thread runs this and feeds the
    //  uniform datapath directly instead of emitting a per-instruction uniformisation loop)
    if (n_tiles > 0 && elect_one()) {
      if constexpr (!TS) {
        if (cta_rank == 0) mbar_arrive_expect_tx(qfull, CG * MT * kQBytes);   // the peer's tiles count here too
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          // slot g (32 rows of the A tile) holds query (g % slots_used): with fewer than 4 queries the
          // query is REPLICATED, so every epilogue warp sees complete rows and takes its own documents.
          // Rows >= lq and queries >= n_queries are out of bounds of the map and arrive as zeros.
#pragma unroll
          for (int g = 0; g < kM / 32; ++g) {
            const int vq = (ZP && g >= p.slots_used) ? p.n_queries : q_base + 4 * mt + (g % p.slots_used);   // ZP: out of bounds -> zeros
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
      }
      int stage = 0; uint32_t phase = 0;
      for (int t = 0; t < n_tiles; ++t) {
        mbar_wait_wd(&empty[stage], phase ^ 1);
        if (p.debug & 2) {
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
      mbar_wait_wd(qfull, 0);
      tc_fence_after_sync();
      const uint32_t sQ_addr = smem_u32(sQ);
      const uint32_t sD_addr = smem_u32(sD);
      int stage = 0; uint32_t phase = 0;
      int ts = 0; uint32_t tphase = 0;
      for (int t = 0; t < n_tiles; ++t) {
        mbar_wait_wd(&full[stage], phase);
        const uint32_t b_addr = sD_addr + stage * kTileBytes;
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          const int unit = ts * kUnitsPerStage + (kUnitsPerStage == 1 ? 0 : mt);
          const int funit = ts * kFullPerStage + (kFullPerStage == 1 ? 0 : mt);
          if (kUnitsPerStage == MT || mt == 0) {   // first M-tile of an accumulator unit: wait until its readers are done
            mbar_wait_wd(&tempty[unit], tphase ^ 1);
            tc_fence_after_sync();
          }
          if (elect_one()) {
            const uint32_t d_tmem = acc_base + uint32_t((ts * MT + mt) * TN);
#pragma unroll
            for (int k = 0; k < HRC_DIM / 16; ++k) {
              if (p.debug & 4) break;   // perf experiment: no tensor work at all (results are garbage)
              // k-th 16-element K slice of B: slab (k / 4), 32 bytes per slice inside the 128-byte row
              const uint64_t b_desc = make_kmajor_sw128_desc(b_addr + (k >> 2) * kHalfTileBytes + (k & 3) * 32);
              if constexpr (TS) {
                // A from TMEM: lane = query row, 8 columns (16 bf16) per K slice
                umma_bf16_ts(d_tmem, tmem_base + uint32_t(mt * 64 + k * 8), b_desc, kIdesc, k > 0 ? 1u : 0u);
              } else {
                const uint32_t a_addr = sQ_addr + mt * kQBytes + (k >> 2) * (kQBytes / 2) + (k & 3) * 32;
                if constexpr (CG == 2) umma_bf16_ss_cg2(d_tmem, make_kmajor_sw128_desc(a_addr), b_desc, kIdesc, k > 0 ? 1u : 0u);
                else umma_bf16_ss(d_tmem, make_kmajor_sw128_desc(a_addr), b_desc, kIdesc, k > 0 ? 1u : 0u);
              }
            }
            // commits (CG == 2: multicast, the peer's producer and epilogue wait on their own copies)
            if (kFullPerStage == MT || mt == MT - 1) {   // accumulators (of this M-tile) ready for the epilogue
              if constexpr (CG == 2) umma_commit_cg2(&tfull[funit]); else umma_commit(&tfull[funit]);
            }
            if (!kForwardEmpty && mt == MT - 1) {   // smem slot reusable once these MMAs have read it
              if constexpr (CG == 2) umma_commit_cg2(&empty[stage]); else umma_commit(&empty[stage]);
            }
          }
          __syncwarp();
        }
        if (++stage == p.n_stages) { stage = 0; phase ^= 1; }
        if (++ts == kTileStages) { ts = 0; tphase ^= 1; }
      }
    }
  } else if (ZP == 0 || (warp >= 4 && (warp & 3) < (ZP == 2 ? 2 : 1) * p.slots_used && (warp >> 2) - 1 < 4 / p.slots_used)) {
    // =============================== epilogue ==================================================
    // slots_used (1, 2 or 4) queries occupy an A tile and each is replicated 4 / slots_used times; the
    // kSplit warps sharing a lane group split further.  The warp of (slot g, share `sub`) scores query
    // g % slots_used for the documents whose local index is congruent to `residue` modulo `rep`.
    const int slot = warp & 3;                           // TMEM lanes 32*slot .. 32*slot+31
    const uint32_t lane_base = uint32_t(slot * 32) << 16;
    const int sub = ZP ? (warp >> 2) - 1 : (warp - kEpiWarp0) >> 2;
    constexpr bool kMtSplit = EPI == 1;                  // the warps of a lane group split M-tiles, not documents
    constexpr bool kMtPass = EPI == 2;                   // one walk per M-tile
    constexpr int kPasses = kMtPass ? MT : 1;
    int mt0 = kMtSplit ? sub : 0;                        // this walk reads M-tiles mt0 .. mt0 + MTW - 1
    const int rep = ZP ? 4 / p.slots_used : (4 / p.slots_used) * (kMtSplit ? 1 : kSplit);
    const int residue = ZP ? sub : (kMtSplit ? slot / p.slots_used : (slot / p.slots_used) * kSplit + sub);
    bool active[MTW];
    int64_t out_row[MTW];
    bool any_active = false;
#pragma unroll
    for (int j = 0; j < MTW; ++j) {
      const int q = ZP == 2 ? q_base + (slot >> 1) : q_base + 4 * (mt0 + j) + (slot % p.slots_used);
      active[j] = q < p.n_queries;
      out_row[j] = int64_t(q) * p.n_items;
      any_active |= active[j];
    }
    // kMtPass: per-pass copies of what differs between the M-tiles (swapped into the [0] slots around each walk)
    bool active_p[kPasses];
    int64_t out_row_p[kPasses];
    float m_p[kPasses], pend_m_p[kPasses];
    int64_t pend_col_p[kPasses];
    bool pending_p[kPasses];
#pragma unroll
    for (int ps = 0; ps < kPasses; ++ps) {
      const int q = q_base + 4 * ps + (slot % p.slots_used);
      active_p[ps] = q < p.n_queries;
      out_row_p[ps] = int64_t(q) * p.n_items;
      m_p[ps] = -INFINITY; pend_m_p[ps] = 0.f; pend_col_p[ps] = 0; pending_p[ps] = false;
      if (kMtPass) any_active |= active_p[ps];
    }

    if constexpr (TS) {
      // Stage the query tiles in TMEM (A operand): this thread owns row (slot, lane) = query token `lane`.
      if ((kMtSplit || sub == 0) && n_tiles > 0) {   // (TS is only instantiated with EPI == 0)
#pragma unroll
        for (int j = 0; j < MTW; ++j) {
          const int mt = mt0 + j;
          uint32_t qv[64];
          const int q = q_base + 4 * mt + (slot % p.slots_used);
          if (q < p.n_queries && lane < p.lq) {
            const uint4* src = reinterpret_cast<const uint4*>(p.queries + (int64_t(q) * p.lq + lane) * HRC_DIM);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const uint4 x = __ldg(src + i);
              qv[4 * i] = x.x; qv[4 * i + 1] = x.y; qv[4 * i + 2] = x.z; qv[4 * i + 3] = x.w;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 64; ++i) qv[i] = 0u;
          }
          tmem_st_32x64(tmem_base + lane_base + uint32_t(mt * 64), qv);
        }
        tmem_st_wait();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(qfull);
      }
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
    float m[MTW];
#pragma unroll
    for (int j = 0; j < MTW; ++j) m[j] = -INFINITY;

    // Emitting a score (a 5-step shuffle butterfly + a store, ~500 cycles of latency) is taken OFF the
    // accumulator hand-shake: finish_doc only parks the finished maxima; they are reduced and stored after this
    // warp has released the tile (or when the next document of the same tile finishes).
    float pend_m[MTW];
    int64_t pend_col = 0;
    bool pending = false;
#pragma unroll
    for (int j = 0; j < MTW; ++j) pend_m[j] = 0.f;

    auto emit_pending = [&]() {
      if constexpr (MTW == 2) {
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
        // M=64: only lanes 0-15 of a lane group hold accumulator rows (16 query tokens)
        const float sc = warp_sum((ZP == 2 && lane >= 16) ? 0.f : pend_m[0]);
        if (lane == 0 && active[0]) {
          if constexpr (ZP == 2) atomicAdd(&p.scores[out_row[0] + pend_col], sc);   // the other token half adds its part
          else p.scores[out_row[0] + pend_col] = sc;
        }
      }
      pending = false;
    };

    auto finish_doc = [&]() {   // park the score(s) of document `my`, move to this warp's next document
      if (pending) emit_pending();
#pragma unroll
      for (int j = 0; j < MTW; ++j) { pend_m[j] = m[j]; m[j] = -INFINITY; }
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
      // kMtPass: the walk state at the start of the tile, replayed for every M-tile
      const int sv_my = my, sv_s = s_tok, sv_e = e_tok, sv_ns = ns_tok, sv_ne = ne_tok, sv_batch = batch;
      const bool sv_have = have_doc;
      const uint32_t sv_ends = ends, sv_ends_next = ends_next;
#pragma unroll
      for (int ps = 0; ps < kPasses; ++ps) {
        if constexpr (kMtPass) {
          my = sv_my; s_tok = sv_s; e_tok = sv_e; ns_tok = sv_ns; ne_tok = sv_ne; batch = sv_batch; have_doc = sv_have;
          ends = sv_ends; ends_next = sv_ends_next;
          mt0 = ps;
          active[0] = active_p[ps]; out_row[0] = out_row_p[ps];
          m[0] = m_p[ps]; pend_m[0] = pend_m_p[ps]; pend_col = pend_col_p[ps]; pending = pending_p[ps];
        }
        const int unit = ts * kUnitsPerStage + (kMtSplit ? sub : (kMtPass ? ps : 0));
        mbar_wait_wd(&tfull[EPI == 3 ? ts * MT : unit], tphase);
        if constexpr (kForwardEmpty) {        // the tile's MMAs are done: its shared-memory slot may be refilled
          if (warp == (ZP ? 4 : kEpiWarp0) && lane == 0) mbar_arrive(&empty[stage_e]);
          if (++stage_e == p.n_stages) stage_e = 0;
        }
        tc_fence_after_sync();
        bool got1 = EPI != 3;                 // EPI == 3: M-tile 1 of this tile has been waited for
        auto need_mt1 = [&]() {
          if constexpr (EPI == 3) {
            if (!got1) { mbar_wait_wd(&tfull[ts * MT + 1], tphase); tc_fence_after_sync(); got1 = true; }
          }
        };
        const int tile0 = t * TN, tile1 = tile0 + TN;
        // accumulator columns of this tile for this warp's j-th M-tile: tacc + j * TN + (token position - tile0)
        const uint32_t tacc = acc_base + lane_base + uint32_t((ts * MT + mt0) * TN);
        uint32_t v[2][32];
        while (have_doc && s_tok < tile1) {
          const int lo = max(s_tok, tile0), hi = min(e_tok, tile1);
          const int len = hi - lo;              // this document's tokens inside this tile
          if (len > 0 && !(p.debug & 1)) {
            if (len >= 32) {
              // Whole 32-column loads that START AT the document's first column (TMEM columns are addressable one by
              // one); the last load is pulled back so that it ENDS at the document's last column — the overlap is
              // harmless under max — so no column is ever masked.
              const int last = hi - 32;
              int c = lo;
              if constexpr (EPI == 3) {
                // M-tile major: all of this document's columns of M-tile 0 (while M-tile 1's MMAs may still be
                // executing), then M-tile 1; software pipeline over chunks with two register buffers
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                  if (j == 1) need_mt1();
                  const uint32_t tj = tacc + uint32_t(j * TN);
                  c = lo;
                  tmem_ld_32x32(tj + uint32_t(min(c, last) - tile0), v[0]);
                  while (true) {
                    tmem_ld_wait();
                    const bool more1 = c < last;
                    if (more1) { c += 32; tmem_ld_32x32(tj + uint32_t(min(c, last) - tile0), v[1]); }
                    m[j] = max32_acc(v[0], m[j]);
                    if (!more1) break;
                    tmem_ld_wait();
                    const bool more0 = c < last;
                    if (more0) { c += 32; tmem_ld_32x32(tj + uint32_t(min(c, last) - tile0), v[0]); }
                    m[j] = max32_acc(v[1], m[j]);
                    if (!more0) break;
                  }
                }
              } else if constexpr (MTW == 2) {
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
              } else if constexpr (kMtSplit) {
                // one M-tile per warp: software pipeline over chunks, two register buffers
                tmem_ld_32x32(tacc + uint32_t(min(c, last) - tile0), v[0]);
                while (true) {
                  tmem_ld_wait();
                  const bool more1 = c < last;
                  if (more1) { c += 32; tmem_ld_32x32(tacc + uint32_t(min(c, last) - tile0), v[1]); }
                  m[0] = max32_acc(v[0], m[0]);
                  if (!more1) break;
                  tmem_ld_wait();
                  const bool more0 = c < last;
                  if (more0) { c += 32; tmem_ld_32x32(tacc + uint32_t(min(c, last) - tile0), v[0]); }
                  m[0] = max32_acc(v[1], m[0]);
                  if (!more0) break;
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
              need_mt1();
#pragma unroll
              for (int j = 0; j < MTW; ++j) tmem_ld_32x32(tacc + uint32_t(j * TN) + uint32_t(cc - tile0), v[j]);
              tmem_ld_wait();
              const int a = lo - cc, b = hi - cc;   // 0 <= a < b <= 32, b - a < 32
              const uint32_t bits = ((1u << (b - a)) - 1u) << a;
#pragma unroll
              for (int j = 0; j < MTW; ++j) m[j] = max32_masked_acc(v[j], bits, m[j]);
            }
          }
          if (e_tok <= tile1) finish_doc(); else break;   // else: the document continues in the next tile
        }
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) {
          if constexpr (CG == 2) mbar_arrive_cluster(&tempty[unit], 0);   // the leader's MMA issuer owns both accumulators
          else mbar_arrive(&tempty[unit]);
        }
        if (pending) emit_pending();          // after the release: off the MMA <-> epilogue critical path
        if constexpr (kMtPass) {
          m_p[ps] = m[0]; pend_m_p[ps] = pend_m[0]; pend_col_p[ps] = pend_col; pending_p[ps] = pending;
        }
      }
      if (++ts == kTileStages) { ts = 0; tphase ^= 1; }
    }
    if constexpr (kMtPass) {
      const int sv_my = my, sv_s = s_tok, sv_e = e_tok, sv_ns = ns_tok, sv_ne = ne_tok, sv_batch = batch;
      const bool sv_have = have_doc;
      const uint32_t sv_ends = ends, sv_ends_next = ends_next;
#pragma unroll
      for (int ps = 0; ps < kPasses; ++ps) {
        my = sv_my; s_tok = sv_s; e_tok = sv_e; ns_tok = sv_ns; ne_tok = sv_ne; batch = sv_batch; have_doc = sv_have;
        ends = sv_ends; ends_next = sv_ends_next;
        active[0] = active_p[ps]; out_row[0] = out_row_p[ps];
        m[0] = m_p[ps]; pend_m[0] = pend_m_p[ps]; pend_col = pend_col_p[ps]; pending = pending_p[ps];
        while (have_doc) finish_doc();
        if (pending) emit_pending();
      }
    } else {
      while (have_doc) finish_doc();        // trailing empty documents (no tokens, no tile): -inf
      if (pending) emit_pending();
    }
  }

  tc_fence_before_sync();
  if constexpr (CG == 2) {
    cluster_sync_all();             // the peer may still be reading operands / arriving on this CTA's barriers
    if (warp == 1) tmem_dealloc_cg2(tmem_base, kTmemCols);
  } else {
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
  }
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

int sm_count() {
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

template <int MT, int TN, bool TS, int ZP = 0, int CG = 1, int EPI = 0>
int launch_cfg(EncodeTiledFn encode, const void* d_tokens, const void* d_queries, TcParams p, dim3 grid,
               cudaStream_t stream) {
  constexpr int kTileBytes = (TN / CG) * HRC_DIM * 2;   // what ONE CTA stages per tile
  CUtensorMap tmap_d, tmap_q;
  {
    cuuint64_t dims[2] = {HRC_DIM, (cuuint64_t)p.total_tokens};
    cuuint64_t strides[1] = {HRC_DIM * 2};
    cuuint32_t box[2] = {64, TN / CG};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&tmap_d, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(d_tokens), dims, strides,
                        box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    HRC_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(tokens) failed: %d", int(r));
  }
  {
    cuuint64_t dims[3] = {HRC_DIM, (cuuint64_t)p.lq, (cuuint64_t)p.n_real_queries};
    cuuint64_t strides[2] = {HRC_DIM * 2, (cuuint64_t)p.lq * HRC_DIM * 2};
    cuuint32_t box[3] = {64, 32, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(&tmap_q, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(d_queries), dims, strides,
                        box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    HRC_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(queries) failed: %d", int(r));
  }
  const int q_bytes = TS ? 0 : MT * (ZP == 2 ? kQTileBytes / 2 : kQTileBytes);
  int stages = (kMaxSmem - 1024 - 512 - q_bytes) / kTileBytes;
  if (stages > 8) stages = 8;
  p.n_stages = stages;
  const int smem_bytes = 1024 + q_bytes + stages * kTileBytes + 512;
  static PerDeviceOnce once;
  int dev;
  if (once.pending(&dev)) {
    HRC_CHECK_CUDA(cudaFuncSetAttribute(maxsim_tc_kernel<MT, TN, TS, ZP, CG, EPI>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    once.mark(dev);
  }
  if (ZP == 2)   // both token halves of a query accumulate into the score
    HRC_CHECK_CUDA(cudaMemsetAsync(p.scores, 0, size_t(p.n_queries) * size_t(p.n_items) * sizeof(float), stream));
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
    HRC_CHECK_CUDA(cudaLaunchKernelEx(&cfg, maxsim_tc_kernel<MT, TN, TS, ZP, CG, EPI>, tmap_d, tmap_q, p));
  } else {
    maxsim_tc_kernel<MT, TN, TS, ZP, CG, EPI><<<grid, cta_threads(MT, ZP), smem_bytes, stream>>>(tmap_d, tmap_q, p);
  }
  count_launch();
  HRC_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

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

static int launch_tc_slots(const void* d_tokens, const int64_t* d_offsets, int64_t n_docs, int64_t total_tokens,
                           const int32_t* d_cand_ids, int64_t n_items, const void* d_queries, int n_real_queries,
                           int q_slots, int lq, float* d_scores, cudaStream_t stream);

int launch_maxsim_tc(const void* d_tokens, const int64_t* d_offsets, int64_t n_docs, int64_t total_tokens,
                     const int32_t* d_cand_ids, int64_t n_items, const void* d_queries, int n_queries,
                     int lq, float* d_scores, cudaStream_t stream) {
  if (n_items == 0 || n_queries == 0) return 0;
  HRC_REQUIRE(lq >= 1 && lq <= HRC_TC_MAX_LQ * HRC_TC_MAX_SLOTS, "tc path: lq=%d not in [1,%d]", lq,
              HRC_TC_MAX_LQ * HRC_TC_MAX_SLOTS);
  const int q_slots = (lq + HRC_TC_MAX_LQ - 1) / HRC_TC_MAX_LQ;
  if (q_slots == 1)
    return launch_tc_slots(d_tokens, d_offsets, n_docs, total_tokens, d_cand_ids, n_items, d_queries, n_queries, 1, lq,
                           d_scores, stream);
  // A query of more than 32 tokens is scored as q_slots virtual queries of <= 32 tokens (rows beyond lq arrive as
  // zeros from TMA and add max_t <0, d_t> = 0); their partial scores are summed in slot order.
  HRC_REQUIRE(int64_t(n_queries) * q_slots <= 65535, "tc path: too many query slots (%d x %d)", n_queries, q_slots);
  float* part = nullptr;
  const int64_t total = int64_t(n_queries) * n_items;
  HRC_CHECK_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&part), size_t(total) * q_slots * sizeof(float), stream));
  int rc = launch_tc_slots(d_tokens, d_offsets, n_docs, total_tokens, d_cand_ids, n_items, d_queries, n_queries, q_slots,
                           lq, part, stream);
  if (rc == 0) {
    sum_slots_kernel<<<unsigned((total + 255) / 256), 256, 0, stream>>>(part, q_slots, n_items, total, d_scores);
    count_launch();
    if (cudaGetLastError() != cudaSuccess) rc = 1;
  }
  cudaFreeAsync(part, stream);
  return rc;
}

static int launch_tc_slots(const void* d_tokens, const int64_t* d_offsets, int64_t n_docs, int64_t total_tokens,
                           const int32_t* d_cand_ids, int64_t n_items, const void* d_queries, int n_real_queries,
                           int q_slots, int lq, float* d_scores, cudaStream_t stream) {
  const int n_queries = n_real_queries * q_slots;      // virtual queries from here on
  HRC_REQUIRE(total_tokens > 0 && total_tokens < (1ll << 31), "tc path: total_tokens=%lld out of range",
              (long long)total_tokens);
  HRC_REQUIRE((reinterpret_cast<uintptr_t>(d_tokens) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_queries) & 15) == 0,
              "tc path: token / query buffers must be 16-byte aligned");
  EncodeTiledFn encode = get_encode_fn();
  HRC_REQUIRE(encode != nullptr, "cuTensorMapEncodeTiled not available from the driver");

  TcParams p;
  p.offsets = d_offsets;
  p.cand_ids = d_cand_ids;
  p.queries = static_cast<const __nv_bfloat16*>(d_queries);
  p.scores = d_scores;
  p.n_docs = n_docs;
  p.total_tokens = total_tokens;
  p.n_items = n_items;
  p.n_queries = n_queries;
  p.n_real_queries = n_real_queries;
  p.q_slots = q_slots;
  p.vq_base = 0;
  p.lq = lq;
  p.n_segments = 1;
  p.n_qgroups = 1;
  p.n_stages = 0;
  p.slots_used = 1;
  p.debug = 0;
  if (const char* e = getenv("HRC_TC_DEBUG")) p.debug = atoi(e);
  p.doc_policy = kEvictFirst;

  if (d_cand_ids != nullptr) {
    HRC_REQUIRE(n_queries <= 65535, "tc path: too many queries for a candidate launch (%d)", n_queries);
    return launch_cfg<1, 128, false, 1>(encode, d_tokens, d_queries, p,
                                        dim3((unsigned)n_items, (unsigned)n_queries), stream);
  }
  if (n_queries <= 4) {
    const int64_t tiles = (total_tokens + 127) / 128;
    p.n_segments = int(tiles < sm_count() ? tiles : sm_count());
    p.slots_used = n_queries == 1 ? 1 : (n_queries == 2 ? 2 : 4);
    const bool zp = getenv("HRC_TC_ZP") == nullptr || atoi(getenv("HRC_TC_ZP")) != 0;   // default; 0 = replicate (A/B)
    const bool m64 = n_queries <= 2 && getenv("HRC_TC_M64") != nullptr && atoi(getenv("HRC_TC_M64")) != 0;
    if (zp && m64) return launch_cfg<1, 128, false, 2>(encode, d_tokens, d_queries, p, dim3((unsigned)p.n_segments), stream);
    if (zp) return launch_cfg<1, 128, false, 1>(encode, d_tokens, d_queries, p, dim3((unsigned)p.n_segments), stream);
    return launch_cfg<1, 128, false>(encode, d_tokens, d_queries, p, dim3((unsigned)p.n_segments), stream);
  }
  p.n_qgroups = (n_queries + 7) / 8;
  p.slots_used = 4;
  p.doc_policy = kEvictNormal;  // the other query groups re-read this tile from L2
  // Batched default: SS operands, N=128.  Measured on C3 (256 queries, power-capped at ~990 W): SS 1123 TFLOP/s,
  // TS (A in TMEM, N=96) 1069 TFLOP/s; with TMA and epilogue disabled both reach the cuBLAS burst rate.
  const bool use_ts = q_slots == 1 && getenv("HRC_TC_TS") != nullptr && atoi(getenv("HRC_TC_TS")) != 0;
  // CTA pairs (cta_group::2, default from 2 query groups up; env HRC_TC_PAIR=0 disables): two query groups of the
  // same corpus segment share every document tile — each CTA stages half of it — so the L2 -> shared-memory
  // traffic and the B-operand reads per SM halve.  C3 (256 queries, 1M ragged documents, power-capped):
  // 1233 vs 1186 TFLOP/s.  An odd last query group runs on the single-CTA kernel.
  const bool use_pair = !use_ts && p.n_qgroups >= 2 && (getenv("HRC_TC_PAIR") == nullptr || atoi(getenv("HRC_TC_PAIR")) != 0);
  if (use_pair) {
    const int64_t tiles = (total_tokens + 127) / 128;
    p.n_segments = int(tiles < sm_count() ? tiles : sm_count());
    const int paired = p.n_qgroups & ~1;                // query groups handled by pairs
    TcParams pp = p;
    pp.n_qgroups = paired;
    pp.n_queries = n_queries < paired * 8 ? n_queries : paired * 8;
    const int epi = getenv("HRC_TC_EPI") != nullptr ? atoi(getenv("HRC_TC_EPI")) : 0;   // see EPI in the kernel
    const dim3 pgrid((unsigned)(pp.n_segments * paired));
    int rc = epi == 1   ? launch_cfg<2, 128, false, 0, 2, 1>(encode, d_tokens, d_queries, pp, pgrid, stream)
             : epi == 2 ? launch_cfg<2, 128, false, 0, 2, 2>(encode, d_tokens, d_queries, pp, pgrid, stream)
             : epi == 3 ? launch_cfg<2, 128, false, 0, 2, 3>(encode, d_tokens, d_queries, pp, pgrid, stream)
                        : launch_cfg<2, 128, false, 0, 2, 0>(encode, d_tokens, d_queries, pp, pgrid, stream);
    if (rc != 0 || paired == p.n_qgroups) return rc;
    TcParams pl = p;                                    // the odd group: (virtual) queries [paired * 8, n_queries)
    pl.n_qgroups = 1;
    pl.vq_base = paired * 8;
    return launch_cfg<2, 128, false>(encode, d_tokens, d_queries, pl, dim3((unsigned)pl.n_segments), stream);
  }
  if (!use_ts) {
    const int64_t tiles = (total_tokens + 127) / 128;
    p.n_segments = int(tiles < sm_count() ? tiles : sm_count());
    return launch_cfg<2, 128, false>(encode, d_tokens, d_queries, p, dim3((unsigned)(p.n_segments * p.n_qgroups)),
                                     stream);
  }
  const int64_t tiles = (total_tokens + 95) / 96;
  p.n_segments = int(tiles < sm_count() ? tiles : sm_count());
  return launch_cfg<2, 96, true>(encode, d_tokens, d_queries, p, dim3((unsigned)(p.n_segments * p.n_qgroups)), stream);
}

}  // namespace hrc
