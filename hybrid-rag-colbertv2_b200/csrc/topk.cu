// topk.cu — per-row top-k over MaxSim scores as exact radix select on 64-bit (score, id) keys.
//
// Replaces torch.topk (local_rag_complete.py:767) and torch.argsort + [:k] (:789-792).
// A key is orderable(score) << 32 | ~id, so keys of one row are unique and totally ordered
// (higher score first, then lower id): the k-th largest key is exact and the result is
// deterministic, which torch.topk's tie order is not.
//
//   stage 1  select_scores_kernel : grid (chunks, rows); each CTA turns <= 8192 scores into keys in
//            shared memory, radix-selects the chunk's top-k (11-bit digits, MSB first, early exit: 3
//            passes for distinct scores, warp-aggregated histogram atomics) and writes k unsorted candidates.  When the row is
//            a single chunk it sorts and writes the final answer itself.
//   stage 2  select_keys_kernel : per row (and per group while the candidate list is too long for
//            shared memory) the same select over candidate keys; the last level bitonic-sorts.
// The same stage-2 kernel is hrc_topk_merge, the on-device merge of all-gathered per-GPU lists.
// HBM traffic: 4 B per score, once.
#include <cstdio>

#include "hrc_common.cuh"

namespace hrc {

namespace {

constexpr int kSelThreads = 1024;
constexpr int kChunk = 8192;          // scores per stage-1 CTA (64 KB of keys)
constexpr int kMergeMax = 24576;      // keys per stage-2 CTA (192 KB)
constexpr int kSortMax = HRC_MAX_TOPK;

constexpr int kBins = 2048;           // 11-bit digits: 64-bit keys in at most 6 passes, usually 3
constexpr int kSmallMerge = 2048;     // merges of up to this many keys sort them outright

struct SelectScratch {
  uint32_t hist[kBins];
  uint64_t prefix;
  int k_rem;
  int done;
  int count;
};

__device__ __forceinline__ int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

// A threshold T such that exactly k keys of keys[0..n) are >= T or, when keys repeat, at least k are >= T and
// fewer than k are > T (n > k >= 1; all threads of the CTA call this).  MSB-first radix select with 11-bit
// digits; stops as soon as the bucket holding the k-th key is needed in full, which for distinct scores is
// after the 32 score bits (3 passes).
__device__ uint64_t radix_kth_largest(const uint64_t* keys, int n, int k, SelectScratch& sc) {
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  uint64_t prefix = 0, mask = 0;
  int k_rem = k;
  const int n_pad = (n + 31) & ~31;
#pragma unroll 1
  for (int pass = 0; pass < 6; ++pass) {
    // digits of 11, 11, 10 bits over the score half, then 11, 11, 10 over the id half
    const int half_pos = pass % 3;
    const int shift = (pass < 3 ? 32 : 0) + (half_pos == 0 ? 21 : (half_pos == 1 ? 10 : 0));
    const uint32_t dmask = half_pos == 2 ? 1023u : 2047u;
    for (int i = tid; i < kBins; i += kSelThreads) sc.hist[i] = 0;
    __syncthreads();
    for (int i = tid; i < n_pad; i += kSelThreads) {
      uint32_t digit = 0xffffffffu;
      if (i < n) {
        const uint64_t key = keys[i];
        if ((key & mask) == prefix) digit = uint32_t(key >> shift) & dmask;
      }
      const uint32_t peers = __match_any_sync(0xffffffffu, digit);
      if (digit != 0xffffffffu && lane == __ffs(peers) - 1) atomicAdd(&sc.hist[digit], __popc(peers));
    }
    __syncthreads();
    if (tid < 32) {
      // lane l owns the 64 bins [kBins-64l-64, kBins-64l): lanes walk the digits from high to low
      const int top = kBins - 1 - 64 * lane;
      uint32_t sum = 0;
      for (int j = 0; j < 64; ++j) sum += sc.hist[top - j];
      uint32_t incl = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
      }
      const uint32_t excl = incl - sum;  // keys with a digit above this lane's bins
      if (excl < uint32_t(k_rem) && incl >= uint32_t(k_rem)) {
        uint32_t above = excl;
        for (int j = 0; j < 64; ++j) {
          const uint32_t c = sc.hist[top - j];
          if (above + c >= uint32_t(k_rem)) {
            sc.prefix = prefix | (uint64_t(top - j) << shift);
            sc.k_rem = k_rem - int(above);
            sc.done = (c == uint32_t(k_rem) - above) ? 1 : 0;   // the whole bucket is needed: no need to refine
            break;
          }
          above += c;
        }
      }
    }
    __syncthreads();
    prefix = sc.prefix;
    k_rem = sc.k_rem;
    mask |= uint64_t(dmask) << shift;
    if (sc.done) break;
  }
  return prefix;
}

// In-place bitonic sort, descending, n_pad a power of two, all threads of the CTA participate.
__device__ void bitonic_sort_desc(uint64_t* a, int n_pad) {
  for (int size = 2; size <= n_pad; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int i = threadIdx.x; i < n_pad / 2; i += kSelThreads) {
        const int lo = 2 * i - (i & (stride - 1));
        const int hi = lo + stride;
        const bool desc = (lo & size) == 0;
        const uint64_t x = a[lo], y = a[hi];
        if ((x < y) == desc) { a[lo] = y; a[hi] = x; }
      }
    }
  }
  __syncthreads();
}

// Top-k of keys[0..n) in shared memory -> out[0..k).  sorted: descending order, else any order.
// Slots beyond min(k, n) are zero.  sort_buf: shared, >= next_pow2(k) entries (only if sorted).
__device__ void select_topk(const uint64_t* keys, int n, int k, uint64_t* out, bool sorted,
                            uint64_t* sort_buf, SelectScratch& sc) {
  const int tid = threadIdx.x;
  const int k_eff = min(k, n);
  uint64_t* dst = sorted ? sort_buf : out;
  const int dst_len = sorted ? next_pow2(max(k, 1)) : k;
  for (int i = tid; i < dst_len; i += kSelThreads) dst[i] = 0;
  if (tid == 0) sc.count = 0;
  __syncthreads();
  if (n <= k) {
    for (int i = tid; i < n; i += kSelThreads) dst[i] = keys[i];
  } else {
    const uint64_t kth = radix_kth_largest(keys, n, k, sc);
    for (int i = tid; i < n; i += kSelThreads) {
      const uint64_t key = keys[i];
      if (key > kth) dst[atomicAdd(&sc.count, 1)] = key;
    }
    __syncthreads();
    for (int i = tid; i < n; i += kSelThreads) {
      if (keys[i] == kth) {
        const int slot = atomicAdd(&sc.count, 1);
        if (slot < k_eff) dst[slot] = kth;
      }
    }
  }
  __syncthreads();
  if (sorted) {
    bitonic_sort_desc(sort_buf, dst_len);
    for (int i = tid; i < k; i += kSelThreads) out[i] = sort_buf[i];
  }
}

__global__ void __launch_bounds__(kSelThreads)
select_scores_kernel(const float* __restrict__ scores, const int32_t* __restrict__ ids, int64_t n,
                     int k, int32_t id_base, uint64_t* __restrict__ out, int n_chunks, int final_sorted) {
  extern __shared__ __align__(16) uint8_t smem[];
  uint64_t* keys = reinterpret_cast<uint64_t*>(smem);
  uint64_t* sort_buf = keys + kChunk;
  __shared__ SelectScratch sc;

  const int chunk = blockIdx.x;
  const int64_t row = blockIdx.y;
  const int64_t begin = int64_t(chunk) * kChunk;
  const int64_t left = n - begin;
  const int cn = int(left < kChunk ? left : kChunk);
  const float* src = scores + row * n + begin;
  const int32_t* id_src = ids ? ids + row * n + begin : nullptr;
  for (int i = threadIdx.x; i < cn; i += kSelThreads) {
    const int32_t id = id_src ? id_src[i] : int32_t(id_base + begin + i);
    keys[i] = make_key(src[i], id);
  }
  __syncthreads();
  uint64_t* dst = out + (row * n_chunks + chunk) * k;
  select_topk(keys, cn, k, dst, final_sorted != 0, sort_buf, sc);
}

// the final answer of a row: sorted keys (+ unpacked ids / scores)
__device__ __forceinline__ void emit_sorted(const uint64_t* sorted, int k, int64_t row, uint64_t* __restrict__ out,
                                            int32_t* __restrict__ ids_out, float* __restrict__ scores_out) {
  for (int i = threadIdx.x; i < k; i += kSelThreads) {
    const uint64_t key = sorted[i];
    out[row * k + i] = key;
    if (ids_out) ids_out[row * k + i] = key ? key_id(key) : -1;
    if (scores_out) scores_out[row * k + i] = key ? key_score(key) : -INFINITY;
  }
}

// Selection fused with the multi-GPU exchange (ONE query, P2P transport): the CTA that holds this GPU's sorted top-k
// stores it straight into every rank's receive slot over NVLink (peer stores, then a system-scope release of the step
// number in each rank's flag), acquires the flags of all ranks in its own buffer, merges the world x k keys and emits
// the GLOBAL top-k — no separate push or merge launch, a sharded search is the same two launches as a local one.
// `sorted` (k keys) and `buf` (>= next_pow2(world * k) keys) are shared memory and may alias.
__device__ __noinline__ void exchange_then_emit(const KeyExchange& x, const uint64_t* sorted, uint64_t* buf, int k,
                                   uint64_t* __restrict__ out, int32_t* __restrict__ ids_out,
                                   float* __restrict__ scores_out) {
  for (int i = threadIdx.x; i < k * x.world; i += kSelThreads) {
    const int dst = i / k, j = i - dst * k;
    uint64_t* slot = reinterpret_cast<uint64_t*>(x.peers[dst] + exchange_slot_offset(x.parity, x.my_rank, x.world, x.max_keys));
    slot[j] = sorted[j];
  }
  __threadfence_system();
  __syncthreads();
  if (int(threadIdx.x) < x.world) {
    uint64_t* flag = reinterpret_cast<uint64_t*>(x.peers[threadIdx.x]) + x.parity * kMaxWorld + x.my_rank;
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(x.seq) : "memory");
    const uint64_t* mine = reinterpret_cast<const uint64_t*>(x.local) + x.parity * kMaxWorld + threadIdx.x;
    uint64_t t0 = 0, v = 0;
    uint32_t spins = 0;
    while (true) {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
      if (v >= x.seq) break;
      if ((++spins & 0xffu) == 0 && x.watchdog_ns != 0) {
        uint64_t now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (t0 == 0) t0 = now;
        if (now - t0 > x.watchdog_ns) {
          printf("hrc: peer %d never published step %llu\n", int(threadIdx.x), x.seq);
          __trap();
        }
      }
    }
  }
  __syncthreads();
  const uint64_t* slots = reinterpret_cast<const uint64_t*>(x.local + exchange_slot_offset(x.parity, 0, x.world, x.max_keys));
  const int n = x.world * k;
  const int n_pad = next_pow2(n);
  for (int i = threadIdx.x; i < n_pad; i += kSelThreads) {
    uint64_t key = 0;
    if (i < n) {
      const int p = i / k, j = i - p * k;
      key = __ldcg(slots + size_t(p) * x.max_keys + j);
    }
    buf[i] = key;
  }
  bitonic_sort_desc(buf, n_pad);
  emit_sorted(buf, k, 0, out, ids_out, scores_out);
}

// ids_out / scores_out (optional, final level only): the unpacked result, so that no separate unpack launch is needed.
// list_len > 0 (final level only): the input is a concatenation of n_lists lists of list_len keys, each SORTED best
// first (what the fused MaxSim epilogue and the streaming top-k hand over).  Let j = ceil(k / n_lists) - 1 and
// m = ceil(k / (j + 1)): the m-th largest of the lists' j-th keys, T, has at least m * (j + 1) >= k keys at or above it
// (j + 1 in each of m lists), so it is a lower bound of the global k-th key — and a tight one: for 148 lists and
// k = 100 it is the 100th best list HEAD, which leaves a few hundred of the 18,944 candidates.  Gather the keys >= T,
// sort them outright.  (If more than kSmallMerge pass — adversarial ties — fall through to the radix select.)
__global__ void __launch_bounds__(kSelThreads)
select_keys_kernel(const uint64_t* __restrict__ keys_in, int n_in, int k, uint64_t* __restrict__ out,
                   int n_groups, int group_len, int final_sorted, int32_t* __restrict__ ids_out,
                   float* __restrict__ scores_out, int list_len, const KeyExchange xch) {
  extern __shared__ __align__(16) uint8_t smem[];
  uint64_t* keys = reinterpret_cast<uint64_t*>(smem);
  uint64_t* sort_buf = keys + group_len;
  __shared__ SelectScratch sc;

  const int group = blockIdx.x;
  const int64_t row = blockIdx.y;
  const int begin = group * group_len;
  const int gn = min(group_len, n_in - begin);
  const uint64_t* src = keys_in + row * n_in + begin;
  if (final_sorted && n_groups == 1 && gn <= kSmallMerge) {      // few keys: sort them outright
    const int n_pad = next_pow2(gn < k ? k : gn);
    for (int i = threadIdx.x; i < n_pad; i += kSelThreads) keys[i] = i < gn ? src[i] : 0;
    bitonic_sort_desc(keys, n_pad);
    if (xch.world > 0) exchange_then_emit(xch, keys, keys, k, out, ids_out, scores_out);
    else emit_sorted(keys, k, row, out, ids_out, scores_out);
    return;
  }
  if (list_len > 0 && final_sorted && n_groups == 1 && k <= list_len && gn > kSmallMerge) {
    __shared__ unsigned long long s_thr;
    const int n_lists = (gn + list_len - 1) / list_len;          // <= kMergeMax / 128 = 192 < kSelThreads
    const int j = (k + n_lists - 1) / n_lists - 1;
    const int m = (k + j) / (j + 1);
    uint64_t* heads = reinterpret_cast<uint64_t*>(sc.hist);       // 192 x 8 B of the 8 KB histogram
    if (threadIdx.x == 0) { s_thr = 0; sc.count = 0; }
    if (int(threadIdx.x) < n_lists) heads[threadIdx.x] = (threadIdx.x * list_len + j < unsigned(gn)) ? src[threadIdx.x * list_len + j] : 0;
    __syncthreads();
    if (int(threadIdx.x) < n_lists) {
      const uint64_t mine = heads[threadIdx.x];
      int rank = 0;
      for (int l = 0; l < n_lists; ++l) rank += (heads[l] > mine || (heads[l] == mine && l < int(threadIdx.x))) ? 1 : 0;
      if (rank == m - 1) s_thr = mine;
    }
    __syncthreads();
    const uint64_t thr = s_thr;
    if (thr != 0) {
      for (int i = threadIdx.x; i < gn; i += kSelThreads) {
        const uint64_t key = src[i];
        if (key >= thr) {
          const int slot = atomicAdd(&sc.count, 1);
          if (slot < kSmallMerge) keys[slot] = key;
        }
      }
      __syncthreads();
      const int got = sc.count;
      if (got <= kSmallMerge) {                       // (got >= k: the list that set T alone holds k keys >= T)
        const int n_pad = next_pow2(got < k ? k : got);
        for (int i = got + threadIdx.x; i < n_pad; i += kSelThreads) keys[i] = 0;
        bitonic_sort_desc(keys, n_pad);
        if (xch.world > 0) exchange_then_emit(xch, keys, keys, k, out, ids_out, scores_out);
        else emit_sorted(keys, k, row, out, ids_out, scores_out);
        return;
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < gn; i += kSelThreads) keys[i] = src[i];
  __syncthreads();
  uint64_t* dst = out + (row * n_groups + group) * k;
  select_topk(keys, gn, k, dst, final_sorted != 0, sort_buf, sc);
  if (final_sorted && xch.world > 0) {                                      // sort_buf holds the sorted keys
    __syncthreads();
    exchange_then_emit(xch, sort_buf, keys, k, out, ids_out, scores_out);
  } else if (final_sorted && (ids_out != nullptr || scores_out != nullptr)) {
    for (int i = threadIdx.x; i < k; i += kSelThreads) {
      const uint64_t key = sort_buf[i];
      if (ids_out) ids_out[row * k + i] = key ? key_id(key) : -1;
      if (scores_out) scores_out[row * k + i] = key ? key_score(key) : -INFINITY;
    }
  }
}

// ---- streaming top-k (k <= 128) --------------------------------------------------------------------------------------
// One CTA per (chunk of a row, row); each of its 8 warps streams a contiguous slice of the scores with coalesced loads,
// eight in flight per lane.  Every lane keeps the warp's threshold score (the k-th best so far), so a score that cannot
// enter the top-k costs one compare; the warp takes the slow path — form the 64-bit keys, append them to its
// shared-memory list, compact a full list with a bitonic sort — only when a ballot says some lane passed.  At the end
// the warps tree-merge their lists and the CTA hands 128 sorted keys on (or, for a single chunk, writes the answer).
// The cost is the HBM read of the scores: C3's 256 x 1M matrix (1 GB) in ~0.3 ms, where the radix select took 3.9 ms.
constexpr int kStreamThreads = 256;
constexpr int kStreamWarps = kStreamThreads / 32;
static_assert(64 * kKeyListOut <= kMergeMax, "stream_chunks' 64 lists per row must fit the merge level");

__host__ __device__ inline int stream_chunks(int64_t n, int n_rows) {
  int64_t want = (2 * 148 + n_rows - 1) / n_rows;              // ~2 CTAs per SM over all rows ...
  const int64_t by_size = (n + 2047) / 2048;                   // ... of at least 2048 scores each ...
  if (want > by_size) want = by_size;
  if (want > 64) want = 64;                                    // ... and at most 64 lists per row for the merge launch
  return int(want < 1 ? 1 : want);
}

__global__ void __launch_bounds__(kStreamThreads)
topk_stream_kernel(const float* __restrict__ scores, const int32_t* __restrict__ ids, int64_t n, int k, int32_t id_base,
                   int n_chunks, int64_t chunk_len, uint64_t* __restrict__ chunk_out, uint64_t* __restrict__ keys_out,
                   int32_t* __restrict__ ids_out, float* __restrict__ scores_out) {
  __shared__ uint64_t lists[kStreamWarps][kKeyListCap];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunk = blockIdx.x;
  const int64_t row = blockIdx.y;
  const int64_t begin = int64_t(chunk) * chunk_len;
  const int64_t end = begin + chunk_len < n ? begin + chunk_len : n;
  const int64_t per_warp = ((end - begin + kStreamWarps - 1) / kStreamWarps + 31) & ~int64_t(31);
  const int64_t w0 = begin + warp * per_warp;
  const int64_t w1 = w0 + per_warp < end ? w0 + per_warp : end;
  const float* src = scores + row * n;
  const int32_t* id_src = ids ? ids + row * n : nullptr;
  uint64_t* lst = lists[warp];
  for (int i = lane; i < kKeyListCap; i += 32) lst[i] = 0;
  __syncwarp();
  int cnt = 0;                 // warp-uniform
  uint64_t thr = 0;            // warp-uniform: the k-th best key so far (0: accept everything)
  float thr_f = -INFINITY;

  auto offer = [&](float s, int64_t col, bool valid) {     // whole warp
    bool pass = valid && !(s < thr_f);                       // (also true for NaN, which make_key orders as -inf)
    if (!__any_sync(0xffffffffu, pass)) return;
    const uint64_t key = pass ? make_key(s, id_src ? id_src[col] : int32_t(id_base + col)) : 0;
    while (true) {
      pass = pass && key > thr;
      const uint32_t b = __ballot_sync(0xffffffffu, pass);
      if (b == 0) break;
      const int add = __popc(b);
      if (cnt + add > kKeyListCap) {                         // full: keep the best k, raise the threshold, look again
        warp_sort256_desc(lst, lane);
        cnt = k;
        thr = lst[k - 1];
        thr_f = key_score(thr);
        continue;
      }
      if (pass) lst[cnt + __popc(b & ((1u << lane) - 1u))] = key;
      cnt += add;
      break;
    }
  };

  constexpr int U = 8;
  int64_t base = w0;                                         // warp-uniform: every lane runs every iteration
  for (; base + int64_t(U) * 32 <= w1; base += int64_t(U) * 32) {
    const int64_t i = base + lane;
    float v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = __ldg(src + i + u * 32);
    bool any = false;
#pragma unroll
    for (int u = 0; u < U; ++u) any |= !(v[u] < thr_f);
    if (__any_sync(0xffffffffu, any)) {
#pragma unroll
      for (int u = 0; u < U; ++u) offer(v[u], i + u * 32, true);
    }
  }
  for (; base < w1; base += 32) {
    const int64_t i = base + lane;
    offer(i < w1 ? __ldg(src + i) : 0.f, i, i < w1);
  }

  warp_sort256_desc(lst, lane);
  // tree-merge the 8 warps' lists: partner's best 128 into the upper half, sort
  for (int st = 1; st < kStreamWarps; st <<= 1) {
    __syncthreads();
    if ((warp & (2 * st - 1)) == 0) {
      const uint64_t* other = lists[warp + st];
      for (int j = lane; j < kKeyListOut; j += 32) lst[kKeyListOut + j] = other[j];
      warp_sort256_desc(lst, lane);
    }
  }
  if (warp == 0) {
    if (n_chunks > 1) {
      uint64_t* out = chunk_out + (row * n_chunks + chunk) * kKeyListOut;
      for (int j = lane; j < kKeyListOut; j += 32) out[j] = lst[j];
    } else {
      for (int j = lane; j < k; j += 32) {
        const uint64_t key = lst[j];
        keys_out[row * k + j] = key;
        if (ids_out) ids_out[row * k + j] = key ? key_id(key) : -1;
        if (scores_out) scores_out[row * k + j] = key ? key_score(key) : -INFINITY;
      }
    }
  }
}

// Merge of per-rank key lists (the multi-GPU exchange step): row r's input is parts[p * part_stride + r * k + j] for
// p < n_parts, j < k — the layout an all-gather of [n_rows][k] blocks produces.  With `flags` (P2P transport) the CTA
// first ACQUIRES flags[0..n_flags) >= seq at system scope: the peers' stores of this step's keys are then visible.
__global__ void __launch_bounds__(kSelThreads)
merge_parts_kernel(const uint64_t* __restrict__ parts, int n_parts, int64_t part_stride, int k, uint64_t* __restrict__ out,
                   int32_t* __restrict__ ids_out, float* __restrict__ scores_out, const uint64_t* flags, uint64_t seq,
                   int n_flags, uint64_t watchdog_ns) {
  extern __shared__ __align__(16) uint8_t smem[];
  uint64_t* keys = reinterpret_cast<uint64_t*>(smem);
  const int n = n_parts * k;
  uint64_t* sort_buf = keys + n;
  __shared__ SelectScratch sc;
  if (flags != nullptr) {
    if (int(threadIdx.x) < n_flags) {
      uint64_t t0 = 0, v = 0;
      uint32_t spins = 0;
      while (true) {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + threadIdx.x) : "memory");
        if (v >= seq) break;
        if ((++spins & 0xffu) == 0 && watchdog_ns != 0) {
          uint64_t now;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
          if (t0 == 0) t0 = now;
          if (now - t0 > watchdog_ns) {
            printf("hrc: peer %d never published step %llu\n", int(threadIdx.x), (unsigned long long)seq);
            __trap();
          }
        }
      }
    }
    __syncthreads();
  }
  const int64_t row = blockIdx.x;
  if (n <= kSmallMerge) {
    // few keys (2..8 ranks x k = 100: the usual case): one bitonic sort of the padded list, no radix passes
    const int n_pad = next_pow2(n < k ? k : n);
    for (int i = threadIdx.x; i < n_pad; i += kSelThreads) {
      uint64_t key = 0;
      if (i < n) {
        const int p = i / k, j = i - p * k;
        key = __ldcg(parts + int64_t(p) * part_stride + row * k + j);
      }
      keys[i] = key;
    }
    bitonic_sort_desc(keys, n_pad);
    for (int i = threadIdx.x; i < k; i += kSelThreads) {
      const uint64_t key = keys[i];
      out[row * k + i] = key;
      if (ids_out) ids_out[row * k + i] = key ? key_id(key) : -1;
      if (scores_out) scores_out[row * k + i] = key ? key_score(key) : -INFINITY;
    }
    return;
  }
  for (int i = threadIdx.x; i < n; i += kSelThreads) {
    const int p = i / k, j = i - p * k;
    keys[i] = __ldcg(parts + int64_t(p) * part_stride + row * k + j);
  }
  __syncthreads();
  select_topk(keys, n, k, out + row * k, true, sort_buf, sc);
  if (ids_out != nullptr || scores_out != nullptr) {
    for (int i = threadIdx.x; i < k; i += kSelThreads) {
      const uint64_t key = sort_buf[i];
      if (ids_out) ids_out[row * k + i] = key ? key_id(key) : -1;
      if (scores_out) scores_out[row * k + i] = key ? key_score(key) : -INFINITY;
    }
  }
}

__global__ void keys_unpack_kernel(const uint64_t* __restrict__ keys, int64_t n, int32_t* __restrict__ ids,
                                   float* __restrict__ scores) {
  const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const uint64_t key = keys[i];
  if (ids) ids[i] = key ? key_id(key) : -1;
  if (scores) scores[i] = key ? key_score(key) : -INFINITY;
}

// keys of a rerank (id = position in the candidate list) -> position, candidate doc id, score
__global__ void rerank_unpack_kernel(const uint64_t* __restrict__ keys, int k, int n_rows, const int32_t* __restrict__ cand,
                                     int n_cand, int32_t* __restrict__ pos_out, int32_t* __restrict__ ids_out,
                                     float* __restrict__ scores_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= k * n_rows) return;
  const int row = i / k;
  const uint64_t key = keys[i];
  const int32_t pos = key ? key_id(key) : -1;
  if (pos_out) pos_out[i] = pos;
  if (ids_out) ids_out[i] = pos >= 0 ? cand[int64_t(row) * n_cand + pos] : -1;
  if (scores_out) scores_out[i] = key ? key_score(key) : -INFINITY;
}

int sort_buf_bytes(int k) {
  int p = 1;
  while (p < k) p <<= 1;
  return p * 8;
}

int configure_smem() {
  static PerDeviceOnce once;
  int dev;
  if (!once.pending(&dev)) return 0;
  HRC_CHECK_CUDA(cudaFuncSetAttribute(select_scores_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      kChunk * 8 + kSortMax * 8));
  HRC_CHECK_CUDA(cudaFuncSetAttribute(select_keys_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      kMergeMax * 8 + kSortMax * 8));
  HRC_CHECK_CUDA(cudaFuncSetAttribute(merge_parts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      kMergeMax * 8 + kSortMax * 8));
  once.mark(dev);
  return 0;
}

// levels of stage 2 over n_in keys per row; writes sorted top-k to d_out.  tmp holds intermediates.
int run_key_levels(const uint64_t* d_in, int64_t n_in, int n_rows, int k, uint64_t* d_out, uint64_t* tmp0,
                   uint64_t* tmp1, cudaStream_t stream, int32_t* d_ids_out = nullptr, float* d_scores_out = nullptr,
                   int list_len = 0, const KeyExchange* xch = nullptr) {
  const uint64_t* cur = d_in;
  int64_t cur_n = n_in;
  int flip = 0;
  while (true) {
    const bool last = cur_n <= kMergeMax;
    const int group_len = last ? int(cur_n) : kMergeMax;
    const int n_groups = int((cur_n + group_len - 1) / group_len);
    uint64_t* dst = last ? d_out : (flip ? tmp1 : tmp0);
    HRC_REQUIRE(dst != nullptr, "top-k: %lld candidate keys per row need workspace", (long long)cur_n);
    size_t smem = size_t(group_len) * 8 + (last ? sort_buf_bytes(k) : 0);
    if (last && group_len <= kSmallMerge) {                  // the direct-sort path pads the keys to a power of two
      size_t n_pad = 1;
      while (n_pad < size_t(group_len) || n_pad < size_t(k)) n_pad <<= 1;
      if (n_pad * 8 > smem) smem = n_pad * 8;
    }
    KeyExchange x;                                           // world == 0: no exchange
    if (last && xch != nullptr) {
      x = *xch;
      size_t n_pad = 1;
      while (n_pad < size_t(x.world) * k) n_pad <<= 1;
      if (n_pad * 8 > smem) smem = n_pad * 8;
    }
    select_keys_kernel<<<dim3(n_groups, n_rows), kSelThreads, smem, stream>>>(
        cur, int(cur_n), k, dst, n_groups, group_len, last ? 1 : 0, last ? d_ids_out : nullptr,
        last ? d_scores_out : nullptr, (last && cur == d_in) ? list_len : 0, x);
    count_launch();
    HRC_CHECK_CUDA(cudaGetLastError());
    if (last) break;
    cur = dst;
    cur_n = int64_t(n_groups) * k;
    flip ^= 1;
  }
  return 0;
}

}  // namespace

int launch_keys_unpack(const uint64_t* d_keys, int64_t n, int32_t* d_ids, float* d_scores, cudaStream_t stream);

size_t topk_workspace_bytes(int64_t n, int n_rows, int k) {
  if (k >= 1 && k <= kKeyListOut && n_rows >= 1 && n > 0) {       // streaming top-k: 128 keys per chunk of a row
    const int c = stream_chunks(n, n_rows);
    return c > 1 ? size_t(n_rows) * size_t(c) * kKeyListOut * 8 + 256 : 0;
  }
  if (n <= kChunk) return 0;
  const int64_t n_chunks = (n + kChunk - 1) / kChunk;
  const int64_t a = n_chunks * k;                      // stage-1 candidates per row
  const int64_t b = ((a + kMergeMax - 1) / kMergeMax) * k;  // first merge level (if needed)
  return size_t(n_rows) * size_t(a + 2 * b) * 8 + 256;
}

int launch_topk(const float* d_scores, const int32_t* d_ids, int64_t n, int n_rows, int k, int32_t id_base,
                uint64_t* d_keys_out, void* d_workspace, size_t workspace_bytes, cudaStream_t stream,
                int32_t* d_ids_out, float* d_scores_out) {
  if (n_rows == 0 || k == 0) return 0;
  HRC_REQUIRE(k >= 1 && k <= HRC_MAX_TOPK, "top-k: k=%d not in [1,%d]", k, HRC_MAX_TOPK);
  HRC_REQUIRE(n_rows <= 65535, "top-k: too many rows (%d)", n_rows);
  if (configure_smem()) return 1;
  if (n <= 0) {
    HRC_CHECK_CUDA(cudaMemsetAsync(d_keys_out, 0, size_t(n_rows) * k * 8, stream));
    if (d_ids_out) HRC_CHECK_CUDA(cudaMemsetAsync(d_ids_out, 0xff, size_t(n_rows) * k * 4, stream));          // -1
    if (d_scores_out) return launch_keys_unpack(d_keys_out, int64_t(n_rows) * k, nullptr, d_scores_out, stream);
    return 0;
  }
  if (k <= kKeyListOut) {
    // streaming top-k: one pass over the scores, then (more than one chunk per row) one merge launch
    const int c = stream_chunks(n, n_rows);
    const int64_t chunk_len = (((n + c - 1) / c) + 255) & ~int64_t(255);
    uint64_t* cand = nullptr;
    if (c > 1) {
      const size_t need = topk_workspace_bytes(n, n_rows, k);
      HRC_REQUIRE(d_workspace != nullptr && workspace_bytes >= need, "top-k: workspace %zu < %zu bytes", workspace_bytes, need);
      cand = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(d_workspace) + 255) & ~uintptr_t(255));
    }
    topk_stream_kernel<<<dim3((unsigned)c, (unsigned)n_rows), kStreamThreads, 0, stream>>>(
        d_scores, d_ids, n, k, id_base, c, chunk_len, cand, d_keys_out, d_ids_out, d_scores_out);
    count_launch();
    HRC_CHECK_CUDA(cudaGetLastError());
    if (c == 1) return 0;
    return run_key_levels(cand, int64_t(c) * kKeyListOut, n_rows, k, d_keys_out, nullptr, nullptr, stream, d_ids_out,
                          d_scores_out, kKeyListOut);
  }
  const int64_t n_chunks = (n + kChunk - 1) / kChunk;
  if (n_chunks == 1) {
    select_scores_kernel<<<dim3(1, n_rows), kSelThreads, kChunk * 8 + sort_buf_bytes(k), stream>>>(
        d_scores, d_ids, n, k, id_base, d_keys_out, 1, 1);
    count_launch();
    HRC_CHECK_CUDA(cudaGetLastError());
    if (d_ids_out != nullptr || d_scores_out != nullptr)
      return launch_keys_unpack(d_keys_out, int64_t(n_rows) * k, d_ids_out, d_scores_out, stream);
    return 0;
  }
  HRC_REQUIRE(n_chunks <= 0x7fffffff, "top-k: row too long");
  const size_t need = topk_workspace_bytes(n, n_rows, k);
  HRC_REQUIRE(d_workspace != nullptr && workspace_bytes >= need, "top-k: workspace %zu < %zu bytes",
              workspace_bytes, need);
  const int64_t a = n_chunks * k;
  const int64_t b = ((a + kMergeMax - 1) / kMergeMax) * k;
  uint64_t* cand = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(d_workspace) + 255) & ~uintptr_t(255));
  uint64_t* tmp0 = cand + size_t(n_rows) * a;
  uint64_t* tmp1 = tmp0 + size_t(n_rows) * b;
  select_scores_kernel<<<dim3((unsigned)n_chunks, n_rows), kSelThreads, kChunk * 8, stream>>>(
      d_scores, d_ids, n, k, id_base, cand, int(n_chunks), 0);
  count_launch();
  HRC_CHECK_CUDA(cudaGetLastError());
  return run_key_levels(cand, a, n_rows, k, d_keys_out, tmp0, tmp1, stream, d_ids_out, d_scores_out);
}

bool topk_exchange_supported(int n_rows, int k, int world, int max_keys) {
  return n_rows == 1 && world >= 1 && world <= kMaxWorld && k >= 1 && k <= max_keys && world * k <= kSmallMerge;
}

int launch_topk_merge(const uint64_t* d_keys_in, int n_in, int n_rows, int k, uint64_t* d_keys_out,
                      cudaStream_t stream, int32_t* d_ids_out, float* d_scores_out, int sorted_list_len,
                      const KeyExchange* xch) {
  if (n_rows == 0 || k == 0) return 0;
  HRC_REQUIRE(xch == nullptr || (n_in > 0 && topk_exchange_supported(n_rows, k, xch->world, xch->max_keys)),
              "top-k merge: fused exchange needs one row and world x k <= %d keys", kSmallMerge);
  HRC_REQUIRE(k >= 1 && k <= HRC_MAX_TOPK, "top-k merge: k=%d not in [1,%d]", k, HRC_MAX_TOPK);
  HRC_REQUIRE(n_in >= 0 && n_in <= kMergeMax, "top-k merge: n_in=%d exceeds %d", n_in, kMergeMax);
  HRC_REQUIRE(n_rows <= 65535, "top-k merge: too many rows (%d)", n_rows);
  if (configure_smem()) return 1;
  if (n_in == 0) {
    HRC_CHECK_CUDA(cudaMemsetAsync(d_keys_out, 0, size_t(n_rows) * k * 8, stream));
    return 0;
  }
  return run_key_levels(d_keys_in, n_in, n_rows, k, d_keys_out, nullptr, nullptr, stream, d_ids_out, d_scores_out,
                        sorted_list_len, xch);
}

uint64_t get_watchdog_ns();

int launch_topk_merge_parts(const uint64_t* d_parts, int n_parts, int part_stride, int n_rows, int k, uint64_t* d_keys_out,
                            cudaStream_t stream, int32_t* d_ids_out, float* d_scores_out, const uint64_t* d_flags,
                            uint64_t seq, int n_flags, uint64_t) {
  if (n_rows == 0 || k == 0) return 0;
  HRC_REQUIRE(k >= 1 && k <= HRC_MAX_TOPK, "merge: k=%d not in [1,%d]", k, HRC_MAX_TOPK);
  HRC_REQUIRE(n_parts >= 1 && int64_t(n_parts) * k <= kMergeMax, "merge: %d x %d keys exceed %d", n_parts, k, kMergeMax);
  HRC_REQUIRE(n_rows <= 65535 && n_flags <= kSelThreads, "merge: too many rows (%d)", n_rows);
  if (configure_smem()) return 1;
  int n_pad = 1;
  while (n_pad < n_parts * k || n_pad < k) n_pad <<= 1;
  const size_t smem = n_parts * k <= kSmallMerge ? size_t(n_pad) * 8 : size_t(n_parts) * k * 8 + sort_buf_bytes(k);
  merge_parts_kernel<<<n_rows, kSelThreads, smem, stream>>>(d_parts, n_parts, part_stride, k, d_keys_out, d_ids_out,
                                                            d_scores_out, d_flags, seq, n_flags, get_watchdog_ns());
  count_launch();
  HRC_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_rerank_unpack(const uint64_t* d_keys, int k, int n_rows, const int32_t* d_cand, int n_cand, int32_t* d_pos,
                         int32_t* d_ids, float* d_scores, cudaStream_t stream) {
  if (k == 0 || n_rows == 0) return 0;
  const int n = k * n_rows;
  rerank_unpack_kernel<<<(n + 255) / 256, 256, 0, stream>>>(d_keys, k, n_rows, d_cand, n_cand, d_pos, d_ids, d_scores);
  count_launch();
  HRC_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_keys_unpack(const uint64_t* d_keys, int64_t n, int32_t* d_ids, float* d_scores, cudaStream_t stream) {
  if (n == 0) return 0;
  keys_unpack_kernel<<<unsigned((n + 255) / 256), 256, 0, stream>>>(d_keys, n, d_ids, d_scores);
  count_launch();
  HRC_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace hrc
