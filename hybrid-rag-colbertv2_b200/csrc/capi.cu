// capi.cu — the extern "C" surface of libhrc.so (declared in include/hrc.h).
// Argument validation, path selection and error reporting live here; kernels live in
// maxsim_tc.cu / maxsim_simt.cu / topk.cu / rrf.cu / synth.cu.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <mutex>
#include <vector>

#include "hrc_common.cuh"

namespace hrc {

static thread_local char g_error[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(uint64_t(n), std::memory_order_relaxed); }

// ---- kernel trace (hrc_trace_enable / hrc_trace_collect): CUDA events around every scoring-kernel launch ----------
struct Trace {
  std::vector<cudaEvent_t> ev;   // 2 * capacity: begin, end, begin, end ...
  int n = 0;                     // launches recorded
  bool open = false;             // a begin without its end
};
static Trace g_trace;
static std::mutex g_trace_mu;

void trace_begin(cudaStream_t stream) {
  if (g_trace.ev.empty()) return;
  std::lock_guard<std::mutex> lk(g_trace_mu);
  if (2 * g_trace.n + 1 >= int(g_trace.ev.size())) return;     // full: later launches are not recorded
  cudaEventRecord(g_trace.ev[2 * g_trace.n], stream);
  g_trace.open = true;
}
void trace_end(cudaStream_t stream) {
  if (g_trace.ev.empty()) return;
  std::lock_guard<std::mutex> lk(g_trace_mu);
  if (!g_trace.open) return;
  cudaEventRecord(g_trace.ev[2 * g_trace.n + 1], stream);
  g_trace.open = false;
  ++g_trace.n;
}

int launch_maxsim_simt(const void*, const int64_t*, int64_t, const int32_t*, int64_t, const void*, int, int, float*,
                       cudaStream_t);
int launch_maxsim_tc(const void*, const int64_t*, int64_t, int64_t, const int32_t*, int64_t, const void*, int, int,
                     float*, int, void*, size_t, cudaStream_t);
size_t maxsim_tc_workspace_bytes(int64_t, int, int);
void set_watchdog_ns(uint64_t);
int store_register(const void*, int64_t);
void store_release(const void*);
#ifdef HRC_EXPERIMENTS
void set_debug(int);
void set_stages(int);
void set_ctas(int);
void set_dyn(int, int);
void set_cta_times(unsigned long long*);
#endif
size_t topk_workspace_bytes(int64_t, int, int);
int launch_topk(const float*, const int32_t*, int64_t, int, int, int32_t, uint64_t*, void*, size_t, cudaStream_t,
                int32_t* = nullptr, float* = nullptr);
int launch_topk_merge(const uint64_t*, int, int, int, uint64_t*, cudaStream_t, int32_t* = nullptr, float* = nullptr, int = 0,
                      const KeyExchange* = nullptr);
bool topk_exchange_supported(int n_rows, int k, int world, int max_keys);
bool tc_topk_supported(int64_t, int, int, int);
bool tc_rerank_supported(int64_t, int, int, int);
int launch_maxsim_tc_rerank(const void*, const int64_t*, int64_t, int64_t, const int32_t*, int, const void*, int, int, int,
                            float*, uint32_t*, int32_t*, int32_t*, float*, cudaStream_t);
int tc_topk_segments(int64_t);
int tc_topk_list_len();
int launch_maxsim_tc_topk(const void*, const int64_t*, int64_t, int64_t, const void*, int, int, int, int32_t, float*,
                          uint64_t*, int, uint32_t*, cudaStream_t);
int launch_keys_unpack(const uint64_t*, int64_t, int32_t*, float*, cudaStream_t);
int launch_rerank_unpack(const uint64_t*, int, int, const int32_t*, int, int32_t*, int32_t*, float*, cudaStream_t);
int launch_rrf(const int32_t*, int, const int32_t*, int, int, int, int, int32_t*, double*, int32_t*, cudaStream_t);
int launch_synth(void*, int64_t, int64_t, uint64_t, cudaStream_t);
int launch_read_probe(const void*, size_t, uint32_t*, cudaStream_t);
int launch_store_validate(const void*, const int64_t*, int64_t, int64_t, int, unsigned long long*, cudaStream_t);
int launch_meanpool_cosine(const void*, const int64_t*, int64_t, const void*, int, int, float*, cudaStream_t);

static int check_device() {
  static int ok[64] = {};   // per device: 0 unknown, 1 sm_100, -1 other
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    set_error("no CUDA device available (libhrc has no CPU fallback)");
    return 3;
  }
  if (dev >= 0 && dev < 64 && ok[dev] == 1) return 0;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
    set_error("no CUDA device available (libhrc has no CPU fallback)");
    return 3;
  }
  if (prop.major != 10) {
    set_error("libhrc is built for sm_100a only; device %d is sm_%d%d", dev, prop.major, prop.minor);
    return 3;
  }
  if (dev >= 0 && dev < 64) ok[dev] = 1;
  return 0;
}

// kernel organisation for one query (maxsim_tc.cu: 0 query-major, 2 doc-major, 3 auto)
static int tc_variant(int path) { return path == HRC_PATH_TC_DM ? 2 : (path == HRC_PATH_AUTO ? 3 : 0); }

static int maxsim_dispatch(const void* d_tokens, const int64_t* d_offsets, int64_t n_docs, int64_t total_tokens,
                           const int32_t* d_cand_ids, int64_t n_items, const void* d_queries, int n_queries, int lq,
                           float* d_scores, int path, void* d_workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (int rc = check_device()) return rc;
  HRC_REQUIRE(n_docs >= 0 && total_tokens >= 0 && n_queries >= 0 && lq >= 1, "maxsim: negative size or lq < 1");
  HRC_REQUIRE(n_items == 0 || n_queries == 0 ||
                  (d_tokens != nullptr && d_offsets != nullptr && d_queries != nullptr && d_scores != nullptr),
              "maxsim: null pointer argument");
  // AUTO: the tensor-core path whenever it applies (lq <= 256).  Measured on B200 (profiles/r02_summary.md, "TC / SIMT
  // crossover"): the tcgen05 kernel is faster than the CUDA-core kernel down to a single query token and a 50-document
  // rerank, because the SIMT kernel is FMA-bound and both pay the same launch latency.
  int variant = tc_variant(path);
  if (path == HRC_PATH_AUTO) path = (lq <= HRC_TC_MAX_LQ * HRC_TC_MAX_SLOTS && total_tokens > 0) ? HRC_PATH_TC : HRC_PATH_SIMT;
  if (path == HRC_PATH_TC || path == HRC_PATH_TC_DM)
    return launch_maxsim_tc(d_tokens, d_offsets, n_docs, total_tokens, d_cand_ids, n_items, d_queries, n_queries, lq,
                            d_scores, variant, d_workspace, workspace_bytes, stream);
  HRC_REQUIRE(path == HRC_PATH_SIMT, "maxsim: unknown path %d", path);
  return launch_maxsim_simt(d_tokens, d_offsets, n_docs, d_cand_ids, n_items, d_queries, n_queries, lq, d_scores,
                            stream);
}

// fp32 -> bf16 (round to nearest even, what torch's .to(torch.bfloat16) does) for queries that arrive from the host
__global__ void f32_to_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int64_t n) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __float2bfloat16_rn(in[i]);
}

static size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

// ids[i] += delta where ids[i] >= 0 (global <-> shard-local document ids; negative = absent stays absent)
__global__ void shift_ids_kernel(int32_t* ids, int64_t n, int32_t delta) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n && ids[i] >= 0) ids[i] += delta;
}

// ---- workspace layouts (every sub-buffer 256-byte aligned) -------------------------------------------------
// hrc_search takes the FUSED route (MaxSim kernel with per-segment top-k in its epilogue, then one merge launch: the
// score matrix is never written) for a SINGLE query of <= 32 tokens with k <= 128 on the tensor-core path — the
// HBM-bound kernel, whose epilogue has slack.  Otherwise: score matrix -> streaming / radix top-k.
static bool search_is_fused(int64_t total_tokens, int nq, int lq, int k, int path) {
  return (path == HRC_PATH_AUTO || path == HRC_PATH_TC || path == HRC_PATH_TC_DM) && tc_topk_supported(total_tokens, nq, lq, k);
}

struct SearchLayout {      // fused: candidate keys, the doc-major kernel's claim counter.  staged: score matrix, slot
                           // partials (lq > 32) or that counter, top-k scratch
  size_t cand, counter, scores, part, topk, total, part_bytes, topk_bytes;
  bool fused;
  int n_seg;
};
static SearchLayout search_layout(int64_t n_docs, int64_t total_tokens, int nq, int lq, int k, int path) {
  SearchLayout L = {};
  size_t o = 0;
  L.fused = search_is_fused(total_tokens, nq, lq, k, path);
  if (L.fused) {
    L.n_seg = tc_topk_segments(total_tokens);
    L.cand = o; o += align256(size_t(nq) * size_t(L.n_seg) * size_t(tc_topk_list_len()) * sizeof(uint64_t));
    L.counter = o; o += 256;
  } else {
    L.scores = o; o += align256(size_t(nq) * size_t(n_docs) * sizeof(float));
    L.part_bytes = maxsim_tc_workspace_bytes(n_docs, nq, lq);
    L.part = o; o += align256(L.part_bytes);
    L.topk_bytes = topk_workspace_bytes(n_docs, nq, k);
    L.topk = o; o += align256(L.topk_bytes);
  }
  L.total = o;
  return L;
}

struct RerankLayout {      // hrc_rerank: candidate scores, per-query arrival counters, slot partials, top-k scratch, keys
  size_t scores, counter, part, topk, keys, total, part_bytes, topk_bytes;
};
static RerankLayout rerank_layout(int n_cand, int nq, int lq, int k) {
  RerankLayout L;
  size_t o = 0;
  L.scores = o; o += align256(size_t(nq) * size_t(n_cand) * sizeof(float));
  L.counter = o; o += align256(size_t(nq) * sizeof(uint32_t));
  L.part_bytes = maxsim_tc_workspace_bytes(n_cand, nq, lq);
  L.part = o; o += align256(L.part_bytes);
  L.topk_bytes = topk_workspace_bytes(n_cand, nq, k);
  L.topk = o; o += align256(L.topk_bytes);
  L.keys = o; o += align256(size_t(nq) * size_t(k) * sizeof(uint64_t));
  L.total = o;
  return L;
}

struct HybridLayout {
  size_t search, keys, col_ids, fused_ids, fused_scores, counts, rerank, pos, total, search_bytes, rerank_bytes;
};
static HybridLayout hybrid_layout(int64_t n_docs, int64_t total_tokens, int nq, int lq, int colbert_k, int n_cand,
                                  int final_k, int path) {
  HybridLayout L;
  size_t o = 0;
  L.search_bytes = search_layout(n_docs, total_tokens, nq, lq, colbert_k, path).total;
  L.search = o; o += align256(L.search_bytes);
  L.keys = o; o += align256(size_t(nq) * colbert_k * sizeof(uint64_t));
  L.col_ids = o; o += align256(size_t(nq) * colbert_k * sizeof(int32_t));
  L.fused_ids = o; o += align256(size_t(nq) * n_cand * sizeof(int32_t));
  L.fused_scores = o; o += align256(size_t(nq) * n_cand * sizeof(double));
  L.counts = o; o += align256(size_t(nq) * sizeof(int32_t));
  L.rerank_bytes = rerank_layout(n_cand, nq, lq, final_k).total;
  L.rerank = o; o += align256(L.rerank_bytes);
  L.pos = o; o += align256(size_t(nq) * final_k * sizeof(int32_t));
  L.total = o;
  return L;
}

struct HostSearchLayout {
  size_t q32, q16, search, keys, ids, out_scores, total, search_bytes;
};
static HostSearchLayout host_search_layout(int64_t n_docs, int64_t total_tokens, int n_queries, int lq, int k, int path) {
  HostSearchLayout L;
  size_t o = 0;
  L.q32 = o; o += align256(size_t(n_queries) * lq * HRC_DIM * sizeof(float));
  L.q16 = o; o += align256(size_t(n_queries) * lq * HRC_DIM * 2);
  L.search_bytes = search_layout(n_docs, total_tokens, n_queries, lq, k, path).total;
  L.search = o; o += align256(L.search_bytes);
  L.keys = o; o += align256(size_t(n_queries) * k * sizeof(uint64_t));
  L.ids = o; o += size_t(n_queries) * k * sizeof(int32_t);          // ids and scores back to back: ONE device -> host
  L.out_scores = o; o += align256(size_t(n_queries) * k * sizeof(float));   // copy when the host buffers are adjacent too
  L.total = align256(o);
  return L;
}

static bool aligned256(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 255) == 0; }

}  // namespace hrc

using namespace hrc;

extern "C" {

int hrc_version(void) { return 200; }

const char* hrc_last_error(void) { return g_error; }

uint64_t hrc_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

void hrc_set_watchdog_ms(uint64_t ms) { set_watchdog_ns(ms * 1000000ull); }

int hrc_trace_enable(int capacity) {
  std::lock_guard<std::mutex> lk(g_trace_mu);
  for (cudaEvent_t e : g_trace.ev) cudaEventDestroy(e);
  g_trace.ev.clear();
  g_trace.n = 0;
  g_trace.open = false;
  HRC_REQUIRE(capacity >= 0 && capacity <= (1 << 20), "trace_enable: capacity %d out of range", capacity);
  g_trace.ev.resize(size_t(2) * capacity);
  for (auto& e : g_trace.ev) HRC_CHECK_CUDA(cudaEventCreate(&e));
  return 0;
}

int hrc_trace_collect(float* ms_out, int max_n) {
  std::lock_guard<std::mutex> lk(g_trace_mu);
  const int n = g_trace.n < max_n ? g_trace.n : max_n;
  for (int i = 0; i < n; ++i) {
    if (cudaEventSynchronize(g_trace.ev[2 * i + 1]) != cudaSuccess ||
        cudaEventElapsedTime(&ms_out[i], g_trace.ev[2 * i], g_trace.ev[2 * i + 1]) != cudaSuccess) {
      set_error("trace_collect: event %d not readable", i);
      return -1;
    }
  }
  g_trace.n = 0;
  g_trace.open = false;
  return n;
}

#ifdef HRC_EXPERIMENTS
void hrc_exp_set_debug(int bits) { set_debug(bits); }
void hrc_exp_set_stages(int n) { set_stages(n); }
void hrc_exp_set_ctas(int n) { set_ctas(n); }
void hrc_exp_set_dyn(int share, int per_cta) { set_dyn(share, per_cta); }
void hrc_exp_set_cta_times(void* d) { set_cta_times(static_cast<unsigned long long*>(d)); }
#endif

int hrc_store_register(const void* d_tokens, int64_t total_tokens) {
  if (int rc = check_device()) return rc;
  return store_register(d_tokens, total_tokens);
}

void hrc_store_release(const void* d_tokens) { store_release(d_tokens); }

size_t hrc_maxsim_workspace_bytes(int64_t n_items, int n_queries, int lq) {
  if (n_items < 0 || n_queries < 0 || lq < 1) return 0;
  return maxsim_tc_workspace_bytes(n_items, n_queries, lq);
}

int hrc_maxsim_scores(const void* d_tokens, const int64_t* d_offsets, int64_t n_docs, int64_t total_tokens,
                      const void* d_queries, int n_queries, int lq, float* d_scores, int path, void* d_workspace,
                      size_t workspace_bytes, void* stream) {
  return maxsim_dispatch(d_tokens, d_offsets, n_docs, total_tokens, nullptr, n_docs, d_queries, n_queries, lq,
                         d_scores, path, d_workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

int hrc_maxsim_scores_ids(const void* d_tokens, const int64_t* d_offsets, int64_t n_docs, int64_t total_tokens,
                          const int32_t* d_cand_ids, int n_cand, const void* d_queries, int n_queries, int lq,
                          float* d_scores, int path, void* d_workspace, size_t workspace_bytes, void* stream) {
  HRC_REQUIRE(n_cand >= 0, "maxsim_ids: n_cand < 0");
  HRC_REQUIRE(n_cand == 0 || d_cand_ids != nullptr, "maxsim_ids: null candidate list");
  return maxsim_dispatch(d_tokens, d_offsets, n_docs, total_tokens, d_cand_ids, n_cand, d_queries, n_queries, lq,
                         d_scores, path, d_workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

int hrc_meanpool_cosine_scores(const void* d_tokens, const int64_t* d_offsets, int64_t n_docs, int64_t total_tokens,
                               const void* d_queries, int n_queries, int lq, float* d_scores, void* stream) {
  if (int rc = check_device()) return rc;
  HRC_REQUIRE(n_docs >= 0 && total_tokens >= 0 && n_queries >= 0 && lq >= 1, "meanpool_cosine: negative size or lq < 1");
  HRC_REQUIRE(n_docs == 0 || n_queries == 0 ||
                  (d_offsets != nullptr && d_queries != nullptr && d_scores != nullptr && (total_tokens == 0 || d_tokens != nullptr)),
              "meanpool_cosine: null pointer argument");
  HRC_REQUIRE((reinterpret_cast<uintptr_t>(d_tokens) & 15) == 0, "meanpool_cosine: token buffer must be 16-byte aligned");
  return launch_meanpool_cosine(d_tokens, d_offsets, n_docs, d_queries, n_queries, lq, d_scores,
                                static_cast<cudaStream_t>(stream));
}

size_t hrc_search_workspace_bytes(int64_t n_docs, int64_t total_tokens, int n_queries, int lq, int k, int path) {
  if (n_docs < 0 || total_tokens < 0 || n_queries < 0 || lq < 1 || k < 0) return 0;
  return search_layout(n_docs, total_tokens, n_queries, lq, k, path).total;
}

int hrc_search(const void* d_tokens, const int64_t* d_offsets, int64_t n_docs, int64_t total_tokens, const void* d_queries,
               int n_queries, int lq, int k, int32_t id_base, void* d_workspace, size_t workspace_bytes,
               uint64_t* d_keys_out, int32_t* d_ids_out, float* d_scores_out, int path, void* stream) {
  return hrc::search_with_exchange(d_tokens, d_offsets, n_docs, total_tokens, d_queries, n_queries, lq, k, id_base,
                                   d_workspace, workspace_bytes, d_keys_out, d_ids_out, d_scores_out, path, stream, nullptr);
}

}  // extern "C"

namespace hrc {

// Can a sharded search hand the exchange to the search's own final kernel?  (one query on the fused-top-k route)
bool search_exchange_supported(int64_t n_docs, int64_t total_tokens, int n_queries, int lq, int k, int path, int world,
                               int max_keys) {
  return k >= 1 && k <= n_docs && search_is_fused(total_tokens, n_queries, lq, k, path) &&
         topk_exchange_supported(n_queries, k, world, max_keys);
}

// hrc_search; with `xch` the final selection kernel also exchanges the keys with the other ranks over peer memory and
// the outputs are the GLOBAL top-k (comm.cu: hrc_sharded_search, P2P transport).
int search_with_exchange(const void* d_tokens, const int64_t* d_offsets, int64_t n_docs, int64_t total_tokens,
                         const void* d_queries, int n_queries, int lq, int k, int32_t id_base, void* d_workspace,
                         size_t workspace_bytes, uint64_t* d_keys_out, int32_t* d_ids_out, float* d_scores_out, int path,
                         void* stream, const KeyExchange* xch) {
  if (int rc = check_device()) return rc;
  HRC_REQUIRE(n_queries >= 0 && lq >= 1 && k >= 0 && k <= n_docs, "search: k=%d must be in [0, n_docs]", k);
  if (n_queries == 0 || k == 0) return 0;
  HRC_REQUIRE(d_keys_out != nullptr && d_workspace != nullptr && d_tokens != nullptr && d_offsets != nullptr &&
                  d_queries != nullptr, "search: null buffer");
  const SearchLayout L = search_layout(n_docs, total_tokens, n_queries, lq, k, path);
  HRC_REQUIRE(workspace_bytes >= L.total, "search: workspace too small (%zu < %zu)", workspace_bytes, L.total);
  HRC_REQUIRE(aligned256(d_workspace), "search: workspace must be 256-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(d_workspace);
  if (L.fused) {
    // MaxSim with the per-segment top-k fused into its epilogue, then ONE launch that merges the segments' lists,
    // sorts and unpacks: two launches per search, no [n_queries x n_docs] matrix in HBM.
    HRC_REQUIRE(n_queries <= 65535, "search: too many queries (%d)", n_queries);
    uint64_t* cand = reinterpret_cast<uint64_t*>(ws + L.cand);
    if (int rc = launch_maxsim_tc_topk(d_tokens, d_offsets, n_docs, total_tokens, d_queries, n_queries, lq, k, id_base,
                                       nullptr, cand, tc_variant(path), reinterpret_cast<uint32_t*>(ws + L.counter), st))
      return rc;
    return launch_topk_merge(cand, L.n_seg * tc_topk_list_len(), n_queries, k, d_keys_out, st, d_ids_out, d_scores_out,
                             tc_topk_list_len(), xch);
  }
  HRC_REQUIRE(xch == nullptr, "search: the fused exchange needs the fused top-k route");
  float* scores = reinterpret_cast<float*>(ws + L.scores);
  if (int rc = maxsim_dispatch(d_tokens, d_offsets, n_docs, total_tokens, nullptr, n_docs, d_queries, n_queries, lq,
                               scores, path, ws + L.part, L.part_bytes, st))
    return rc;
  return launch_topk(scores, nullptr, n_docs, n_queries, k, id_base, d_keys_out, ws + L.topk, L.topk_bytes, st, d_ids_out,
                     d_scores_out);
}

}  // namespace hrc

extern "C" {

size_t hrc_search_host_workspace_bytes(int64_t n_docs, int64_t total_tokens, int n_queries, int lq, int k, int path) {
  if (n_docs < 0 || total_tokens < 0 || n_queries < 0 || lq < 1 || k < 0) return 0;
  return host_search_layout(n_docs, total_tokens, n_queries, lq, k, path).total;
}

int hrc_search_host(const void* d_tokens, const int64_t* d_offsets, int64_t n_docs, int64_t total_tokens,
                    const float* h_queries, int n_queries, int lq, int k, int32_t id_base, void* d_workspace,
                    size_t workspace_bytes, int32_t* h_ids_out, float* h_scores_out, int path, void* stream) {
  if (int rc = check_device()) return rc;
  HRC_REQUIRE(n_queries >= 0 && lq >= 1 && k >= 0 && k <= n_docs, "search_host: bad sizes (k=%d must be in [0, n_docs])", k);
  if (n_queries == 0 || k == 0) return 0;
  HRC_REQUIRE(h_queries != nullptr && h_ids_out != nullptr && h_scores_out != nullptr && d_workspace != nullptr,
              "search_host: null buffer");
  const HostSearchLayout L = host_search_layout(n_docs, total_tokens, n_queries, lq, k, path);
  HRC_REQUIRE(workspace_bytes >= L.total, "search_host: workspace too small (%zu < %zu)", workspace_bytes, L.total);
  HRC_REQUIRE(aligned256(d_workspace), "search_host: workspace must be 256-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(d_workspace);
  float* q32 = reinterpret_cast<float*>(ws + L.q32);
  __nv_bfloat16* q16 = reinterpret_cast<__nv_bfloat16*>(ws + L.q16);
  const int64_t nq_elems = int64_t(n_queries) * lq * HRC_DIM;
  HRC_CHECK_CUDA(cudaMemcpyAsync(q32, h_queries, size_t(nq_elems) * sizeof(float), cudaMemcpyHostToDevice, st));
  f32_to_bf16_kernel<<<unsigned((nq_elems + 255) / 256), 256, 0, st>>>(q32, q16, nq_elems);
  count_launch();
  HRC_CHECK_CUDA(cudaGetLastError());
  int32_t* d_ids = reinterpret_cast<int32_t*>(ws + L.ids);
  float* d_sc = reinterpret_cast<float*>(ws + L.out_scores);
  if (int rc = hrc_search(d_tokens, d_offsets, n_docs, total_tokens, q16, n_queries, lq, k, id_base, ws + L.search,
                          L.search_bytes, reinterpret_cast<uint64_t*>(ws + L.keys), d_ids, d_sc, path, stream))
    return rc;
  const size_t half = size_t(n_queries) * k * sizeof(int32_t);
  if (reinterpret_cast<const uint8_t*>(h_scores_out) == reinterpret_cast<const uint8_t*>(h_ids_out) + half) {
    HRC_CHECK_CUDA(cudaMemcpyAsync(h_ids_out, d_ids, 2 * half, cudaMemcpyDeviceToHost, st));
  } else {
    HRC_CHECK_CUDA(cudaMemcpyAsync(h_ids_out, d_ids, half, cudaMemcpyDeviceToHost, st));
    HRC_CHECK_CUDA(cudaMemcpyAsync(h_scores_out, d_sc, half, cudaMemcpyDeviceToHost, st));
  }
  return 0;
}

size_t hrc_rerank_workspace_bytes(int n_cand, int n_queries, int lq, int k) {
  if (n_cand < 0 || n_queries < 0 || lq < 1 || k < 0) return 0;
  return rerank_layout(n_cand, n_queries, lq, k).total;
}

int hrc_rerank(const void* d_tokens, const int64_t* d_offsets, int64_t n_docs, int64_t total_tokens,
               const int32_t* d_cand_ids, int n_cand, const void* d_queries, int n_queries, int lq, int k,
               void* d_workspace, size_t workspace_bytes, int32_t* d_pos_out, int32_t* d_ids_out, float* d_scores_out,
               float* d_cand_scores_out, int path, void* stream) {
  HRC_REQUIRE(n_cand >= 0 && n_queries >= 0 && lq >= 1, "rerank: negative size or lq < 1");
  HRC_REQUIRE(k >= 0 && k <= n_cand, "rerank: k=%d must be in [0, n_cand]", k);
  if (n_queries == 0 || k == 0) return 0;
  HRC_REQUIRE(d_cand_ids != nullptr && d_workspace != nullptr, "rerank: null buffer");
  const RerankLayout L = rerank_layout(n_cand, n_queries, lq, k);
  HRC_REQUIRE(workspace_bytes >= L.total, "rerank: workspace too small (%zu < %zu)", workspace_bytes, L.total);
  HRC_REQUIRE(aligned256(d_workspace), "rerank: workspace must be 256-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(d_workspace);
  float* scores = d_cand_scores_out != nullptr ? d_cand_scores_out : reinterpret_cast<float*>(ws + L.scores);
  uint64_t* keys = reinterpret_cast<uint64_t*>(ws + L.keys);
  if ((path == HRC_PATH_AUTO || path == HRC_PATH_TC) && tc_rerank_supported(total_tokens, lq, n_cand, k) &&
      n_queries <= 65535) {
    // ONE launch: every CTA scores one candidate, the last CTA of a query ranks them and writes the sorted top-k
    if (int rc = check_device()) return rc;
    HRC_REQUIRE(d_tokens != nullptr && d_offsets != nullptr && d_queries != nullptr && d_pos_out != nullptr &&
                    d_scores_out != nullptr, "rerank: null buffer");
    return launch_maxsim_tc_rerank(d_tokens, d_offsets, n_docs, total_tokens, d_cand_ids, n_cand, d_queries, n_queries, lq,
                                   k, scores, reinterpret_cast<uint32_t*>(ws + L.counter), d_pos_out, d_ids_out,
                                   d_scores_out, st);
  }
  if (int rc = maxsim_dispatch(d_tokens, d_offsets, n_docs, total_tokens, d_cand_ids, n_cand, d_queries, n_queries, lq,
                               scores, path, ws + L.part, L.part_bytes, st))
    return rc;
  if (int rc = launch_topk(scores, nullptr, n_cand, n_queries, k, 0, keys, ws + L.topk, L.topk_bytes, st)) return rc;
  return launch_rerank_unpack(keys, k, n_queries, d_cand_ids, n_cand, d_pos_out, d_ids_out, d_scores_out, st);
}

size_t hrc_hybrid_retrieve_workspace_bytes(int64_t n_docs, int64_t total_tokens, int n_queries, int lq, int colbert_k,
                                           int n_candidates, int final_k, int path) {
  if (n_docs < 0 || total_tokens < 0 || n_queries < 0 || lq < 1 || colbert_k < 0 || n_candidates < 0 || final_k < 0) return 0;
  return hybrid_layout(n_docs, total_tokens, n_queries, lq, colbert_k, n_candidates, final_k, path).total;
}

int hrc_hybrid_retrieve(const void* d_tokens, const int64_t* d_offsets, int64_t n_docs, int64_t total_tokens,
                        const void* d_queries, int n_queries, int lq, const int32_t* d_bm25_ids, int n_bm25, int colbert_k,
                        int rrf_k, int n_candidates, int final_k, int32_t id_base, void* d_workspace,
                        size_t workspace_bytes, int32_t* d_ids_out, float* d_scores_out, int path, void* stream) {
  if (int rc = check_device()) return rc;
  HRC_REQUIRE(n_queries >= 0 && lq >= 1 && n_bm25 >= 0 && colbert_k >= 1 && colbert_k <= n_docs && n_candidates >= 1 &&
                  final_k >= 1 && final_k <= n_candidates,
              "hybrid_retrieve: need 1 <= colbert_k <= n_docs and 1 <= final_k <= n_candidates");
  if (n_queries == 0) return 0;
  HRC_REQUIRE(d_workspace != nullptr && d_ids_out != nullptr && d_scores_out != nullptr && (n_bm25 == 0 || d_bm25_ids != nullptr),
              "hybrid_retrieve: null buffer");
  const HybridLayout L = hybrid_layout(n_docs, total_tokens, n_queries, lq, colbert_k, n_candidates, final_k, path);
  HRC_REQUIRE(workspace_bytes >= L.total, "hybrid_retrieve: workspace too small (%zu < %zu)", workspace_bytes, L.total);
  HRC_REQUIRE(aligned256(d_workspace), "hybrid_retrieve: workspace must be 256-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(d_workspace);
  int32_t* col_ids = reinterpret_cast<int32_t*>(ws + L.col_ids);
  int32_t* fused = reinterpret_cast<int32_t*>(ws + L.fused_ids);
  // stage 2 of retrieve (:908-911): ColBERT first stage over the whole store, GLOBAL ids
  if (int rc = hrc_search(d_tokens, d_offsets, n_docs, total_tokens, d_queries, n_queries, lq, colbert_k, id_base,
                          ws + L.search, L.search_bytes, reinterpret_cast<uint64_t*>(ws + L.keys), col_ids, nullptr, path,
                          stream))
    return rc;
  // stage 3 (:914-916): RRF of the BM25 list with the ColBERT list, top n_candidates
  if (int rc = launch_rrf(d_bm25_ids, n_bm25, col_ids, colbert_k, n_queries, rrf_k, n_candidates, fused,
                          reinterpret_cast<double*>(ws + L.fused_scores), reinterpret_cast<int32_t*>(ws + L.counts), st))
    return rc;
  const int64_t n_fused = int64_t(n_queries) * n_candidates;
  if (id_base != 0) {   // candidates are looked up by shard-local id
    shift_ids_kernel<<<unsigned((n_fused + 255) / 256), 256, 0, st>>>(fused, n_fused, -id_base);
    count_launch();
  }
  // stage 5 (:926-929): rerank the candidates' STORED token embeddings, sorted top final_k
  if (int rc = hrc_rerank(d_tokens, d_offsets, n_docs, total_tokens, fused, n_candidates, d_queries, n_queries, lq, final_k,
                          ws + L.rerank, L.rerank_bytes, reinterpret_cast<int32_t*>(ws + L.pos), d_ids_out, d_scores_out,
                          nullptr, path, stream))
    return rc;
  if (id_base != 0) {
    const int64_t n_out = int64_t(n_queries) * final_k;
    shift_ids_kernel<<<unsigned((n_out + 255) / 256), 256, 0, st>>>(d_ids_out, n_out, id_base);
    count_launch();
  }
  HRC_CHECK_CUDA(cudaGetLastError());
  return 0;
}

size_t hrc_topk_workspace_bytes(int64_t n, int n_rows, int k) { return topk_workspace_bytes(n, n_rows, k); }

int hrc_topk(const float* d_scores, const int32_t* d_ids, int64_t n, int n_rows, int k, int32_t id_base,
             uint64_t* d_keys_out, void* d_workspace, size_t workspace_bytes, void* stream) {
  if (int rc = check_device()) return rc;
  HRC_REQUIRE(n >= 0 && n_rows >= 0 && k >= 0, "topk: negative size");
  HRC_REQUIRE(n_rows == 0 || k == 0 || d_keys_out != nullptr, "topk: null output");
  HRC_REQUIRE(n == 0 || n_rows == 0 || k == 0 || d_scores != nullptr, "topk: null scores");
  return launch_topk(d_scores, d_ids, n, n_rows, k, id_base, d_keys_out, d_workspace, workspace_bytes,
                     static_cast<cudaStream_t>(stream));
}

int hrc_topk_merge(const uint64_t* d_keys_in, int n_in, int n_rows, int k, uint64_t* d_keys_out, void* stream) {
  if (int rc = check_device()) return rc;
  HRC_REQUIRE(n_in >= 0 && n_rows >= 0 && k >= 0, "topk_merge: negative size");
  HRC_REQUIRE(n_rows == 0 || k == 0 || d_keys_out != nullptr, "topk_merge: null output");
  return launch_topk_merge(d_keys_in, n_in, n_rows, k, d_keys_out, static_cast<cudaStream_t>(stream));
}

int hrc_keys_unpack(const uint64_t* d_keys, int64_t n, int32_t* d_ids_out, float* d_scores_out, void* stream) {
  if (int rc = check_device()) return rc;
  HRC_REQUIRE(n >= 0, "keys_unpack: negative size");
  return launch_keys_unpack(d_keys, n, d_ids_out, d_scores_out, static_cast<cudaStream_t>(stream));
}

int hrc_rrf_fuse(const int32_t* d_ids_a, int n_a, const int32_t* d_ids_b, int n_b, int n_rows, int rrf_k, int top_n,
                 int32_t* d_ids_out, double* d_scores_out, int32_t* d_counts_out, void* stream) {
  if (int rc = check_device()) return rc;
  HRC_REQUIRE(n_rows >= 0 && top_n >= 0, "rrf: negative size");
  HRC_REQUIRE(n_rows == 0 || top_n == 0 || (d_ids_out != nullptr && d_scores_out != nullptr), "rrf: null output");
  return launch_rrf(d_ids_a, n_a, d_ids_b, n_b, n_rows, rrf_k, top_n, d_ids_out, d_scores_out, d_counts_out,
                    static_cast<cudaStream_t>(stream));
}

int hrc_synth_tokens(void* d_tokens_out, int64_t token_begin, int64_t n_tokens, uint64_t seed, void* stream) {
  if (int rc = check_device()) return rc;
  return launch_synth(d_tokens_out, token_begin, n_tokens, seed, static_cast<cudaStream_t>(stream));
}

int hrc_store_validate(const void* d_tokens, const int64_t* d_offsets, int64_t n_docs, int64_t total_tokens,
                       int check_values, void* d_workspace, size_t workspace_bytes, int64_t* report_out, void* stream) {
  if (int rc = check_device()) return rc;
  HRC_REQUIRE(d_offsets != nullptr && n_docs >= 0 && total_tokens >= 0 && (total_tokens == 0 || d_tokens != nullptr),
              "store_validate: null store");
  HRC_REQUIRE(d_workspace != nullptr && workspace_bytes >= 32 && (reinterpret_cast<uintptr_t>(d_workspace) & 7) == 0,
              "store_validate: workspace of 32 bytes, 8-byte aligned, required");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  auto* d_report = static_cast<unsigned long long*>(d_workspace);
  if (int rc = launch_store_validate(d_tokens, d_offsets, n_docs, total_tokens, check_values, d_report, s)) return rc;
  unsigned long long rep[4];
  HRC_CHECK_CUDA(cudaMemcpyAsync(rep, d_report, sizeof(rep), cudaMemcpyDeviceToHost, s));
  HRC_CHECK_CUDA(cudaStreamSynchronize(s));
  const int64_t first_bad = rep[1] ? n_docs + 1 - int64_t(rep[1]) : -1;
  if (report_out) {
    report_out[0] = int64_t(rep[0]);
    report_out[1] = first_bad;
    report_out[2] = int64_t(rep[2]);
    report_out[3] = int64_t(rep[3]);
  }
  if (rep[0]) {
    set_error("store_validate: %llu offsets entries break the CSR contract (offsets[0] = 0, non-decreasing, "
              "offsets[n_docs] = total_tokens = %lld); first at entry %lld", rep[0], (long long)total_tokens, (long long)first_bad);
    return 3;
  }
  if (rep[2]) {
    set_error("store_validate: %llu token values are NaN or infinite", rep[2]);
    return 4;
  }
  return 0;
}

int hrc_read_probe(const void* d_buf, size_t bytes, uint32_t* d_out, void* stream) {
  if (int rc = check_device()) return rc;
  return launch_read_probe(d_buf, bytes, d_out, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
