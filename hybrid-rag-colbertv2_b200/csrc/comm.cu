// comm.cu — the multi-GPU exchange step of the document-sharded search, inside libhrc.so.
//
// The corpus shards by document: each rank (one process per GPU) computes a local top-k, the
// ranks exchange k 64-bit (score, global doc id) keys each, and every rank merges world*k keys on its device.  The
// reference is single-process (local_rag_complete.py has no distributed code), so this step has no counterpart
// there; it exists so that the N>1 search is ONE C call like the N=1 search.  Two transports:
//   NCCL  ncclAllGather of k*8 bytes per rank (libnccl.so.2 is dlopen'ed at hrc_comm_init: libhrc.so has no link-time
//         dependency on it and picks up the copy the host process has already loaded, e.g. torch's);
//   P2P   each rank STORES its keys straight into every peer's receive buffer over NVLink (CUDA IPC mapped memory,
//         set up once in hrc_comm_enable_p2p) followed by a system-scope release of a sequence flag; the merge kernel
//         acquires the world flags and merges — the collective is fused into the producer and the consumer kernels,
//         no collective launch, no host involvement.  Receive slots are double-buffered by sequence parity: a rank
//         can be at most one step ahead of a peer, because its own merge needs that peer's keys of the same step.
#include <dlfcn.h>

// The handful of NCCL declarations this file needs, restated from nccl.h (stable since NCCL 2.0) so that neither
// building nor loading libhrc.so depends on an NCCL installation: the library is dlopen'ed in hrc_comm_init.
extern "C" {
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef enum { ncclSuccess = 0 } ncclResult_t;
typedef enum { ncclInt8 = 0, ncclUint8 = 1, ncclInt32 = 2, ncclUint32 = 3, ncclInt64 = 4, ncclUint64 = 5 } ncclDataType_t;
}

#include <cstdio>
#include <cstring>

#include "hrc_common.cuh"

namespace hrc {

int launch_topk_merge_parts(const uint64_t*, int, int, int, int, uint64_t*, cudaStream_t, int32_t*, float*, const uint64_t*,
                            uint64_t, int, uint64_t);
int launch_keys_unpack(const uint64_t*, int64_t, int32_t*, float*, cudaStream_t);
uint64_t get_watchdog_ns();
int launch_rerank_unpack(const uint64_t*, int, int, const int32_t*, int, int32_t*, int32_t*, float*, cudaStream_t);
int launch_rrf(const int32_t*, int, const int32_t*, int, int, int, int, int32_t*, double*, int32_t*, cudaStream_t);

namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;

int load_nccl() {
  if (g_nccl.handle != nullptr) return 0;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  HRC_REQUIRE(h != nullptr, "comm: cannot load libnccl.so.2 (%s)", dlerror());
  NcclApi a;
  a.handle = h;
  a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
  a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
  a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
  a.AllGather = reinterpret_cast<decltype(a.AllGather)>(dlsym(h, "ncclAllGather"));
  a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
  HRC_REQUIRE(a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllGather && a.GetErrorString,
              "comm: libnccl.so.2 lacks a required symbol");
  g_nccl = a;
  return 0;
}

#define HRC_CHECK_NCCL(expr)                                                                           \
  do {                                                                                                 \
    ncclResult_t _r = (expr);                                                                          \
    if (_r != ncclSuccess) {                                                                           \
      hrc::set_error("%s failed: %s (%s:%d)", #expr, g_nccl.GetErrorString(_r), __FILE__, __LINE__);   \
      return 1;                                                                                        \
    }                                                                                                  \
  } while (0)

}  // namespace

struct Comm {
  ncclComm_t nccl = nullptr;
  int world = 0, rank = 0, device = 0;
  // P2P transport (hrc_comm_enable_p2p)
  bool p2p = false;
  int max_keys = 0;                       // keys per rank and step (n_rows * k) the slots can hold
  uint8_t* local = nullptr;               // this rank's receive buffer: flags[2][kMaxWorld] then slots[2][world][max_keys]
  uint8_t* peer[kMaxWorld] = {};          // every rank's receive buffer as mapped here (peer[rank] == local)
  uint8_t** d_peer = nullptr;             // the same table on the device
  uint64_t seq = 0;                       // steps done
};

namespace {

constexpr size_t kFlagBytes = kExchangeFlagBytes;
__host__ __device__ inline size_t slot_offset(int parity, int src_rank, int world, int max_keys) {
  return exchange_slot_offset(parity, src_rank, world, max_keys);
}

// One CTA per destination rank: copy this rank's keys into slot[parity][my_rank] of the destination's receive buffer,
// make them visible system-wide, then publish the step's sequence number in the destination's flag[parity][my_rank].
__global__ void __launch_bounds__(256)
p2p_push_kernel(uint8_t* const* __restrict__ peers, const uint64_t* __restrict__ keys, int n_keys, int world, int my_rank,
                int max_keys, int parity, uint64_t seq) {
  const int dst = blockIdx.x;
  uint8_t* base = peers[dst];
  uint64_t* slot = reinterpret_cast<uint64_t*>(base + slot_offset(parity, my_rank, world, max_keys));
  for (int i = threadIdx.x; i < n_keys; i += blockDim.x) slot[i] = keys[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    uint64_t* flag = reinterpret_cast<uint64_t*>(base) + parity * kMaxWorld + my_rank;
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(seq) : "memory");
  }
}

// global candidate ids -> ids local to this shard (-1: not mine)
__global__ void localize_ids_kernel(const int32_t* __restrict__ ids, int64_t n, int32_t id_base, int64_t n_docs,
                                    int32_t* __restrict__ local) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t l = int64_t(ids[i]) - id_base;
  local[i] = (ids[i] >= 0 && l >= 0 && l < n_docs) ? int32_t(l) : -1;
}

// (score, candidate position) keys of the candidates THIS rank answers for: the ones its shard holds, and — on rank 0 —
// the ones no shard holds (absent / out of range: -inf, exactly what the single-GPU rerank gives them).  Others: 0.
__global__ void owned_rerank_keys_kernel(const float* __restrict__ scores, const int32_t* __restrict__ global_ids,
                                         int n_cand, int64_t n, int32_t id_base, int64_t n_docs, int64_t n_docs_global,
                                         int rank, uint64_t* __restrict__ keys) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t id = global_ids[i];
  const int pos = int(i % n_cand);
  const int64_t l = int64_t(id) - id_base;
  uint64_t key = 0;
  if (id >= 0 && l >= 0 && l < n_docs) key = make_key(scores[i], pos);
  else if (rank == 0 && (id < 0 || id >= n_docs_global)) key = make_key(-INFINITY, pos);
  keys[i] = key;
}

size_t al256(size_t x) { return (x + 255) & ~size_t(255); }

struct ShardedSearchLayout { size_t search, local, gather, total, search_bytes, gather_bytes; };
ShardedSearchLayout sharded_search_layout(int world, int64_t n_docs, int64_t total_tokens, int nq, int lq, int k, int path) {
  ShardedSearchLayout L;
  size_t o = 0;
  // (a shard with fewer than k documents searches with k_local = n_docs, possibly on the other route)
  const size_t a = hrc_search_workspace_bytes(n_docs, total_tokens, nq, lq, k, path);
  const size_t b = hrc_search_workspace_bytes(n_docs, total_tokens, nq, lq, int(n_docs < k ? n_docs : k), path);
  L.search_bytes = a > b ? a : b;
  L.search = o; o += al256(L.search_bytes);
  L.local = o; o += al256(size_t(nq) * k * sizeof(uint64_t));
  L.gather_bytes = hrc_allgather_merge_workspace_bytes(world, nq, k);
  L.gather = o; o += al256(L.gather_bytes);
  L.total = o;
  return L;
}

struct ShardedHostLayout { size_t q32, q16, inner, ids, scores, total, inner_bytes; };
ShardedHostLayout sharded_host_layout(int world, int64_t n_docs, int64_t total_tokens, int nq, int lq, int k, int path) {
  ShardedHostLayout L;
  size_t o = 0;
  L.q32 = o; o += al256(size_t(nq) * lq * HRC_DIM * sizeof(float));
  L.q16 = o; o += al256(size_t(nq) * lq * HRC_DIM * 2);
  L.inner_bytes = sharded_search_layout(world, n_docs, total_tokens, nq, lq, k, path).total + al256(size_t(nq) * k * 8);
  L.inner = o; o += al256(L.inner_bytes);
  L.ids = o; o += size_t(nq) * k * sizeof(int32_t);                 // ids and scores back to back: one D2H copy
  L.scores = o; o += al256(size_t(nq) * k * sizeof(float));
  L.total = al256(o);
  return L;
}

struct ShardedHybridLayout {
  size_t search, local, gather, gkeys, col_ids, fused, fused_scores, counts, local_cand, cand_scores, part, rr_keys, gather2,
      fkeys, pos, total, search_bytes, gather_bytes, part_bytes, gather2_bytes;
};
ShardedHybridLayout sharded_hybrid_layout(int world, int64_t n_docs, int64_t total_tokens, int nq, int lq, int ck, int nc,
                                          int fk, int path) {
  ShardedHybridLayout L;
  size_t o = 0;
  L.search_bytes = sharded_search_layout(world, n_docs, total_tokens, nq, lq, ck, path).search_bytes;
  L.search = o; o += al256(L.search_bytes);
  L.local = o; o += al256(size_t(nq) * ck * 8);
  L.gather_bytes = hrc_allgather_merge_workspace_bytes(world, nq, ck);
  L.gather = o; o += al256(L.gather_bytes);
  L.gkeys = o; o += al256(size_t(nq) * ck * 8);
  L.col_ids = o; o += al256(size_t(nq) * ck * 4);
  L.fused = o; o += al256(size_t(nq) * nc * 4);
  L.fused_scores = o; o += al256(size_t(nq) * nc * 8);
  L.counts = o; o += al256(size_t(nq) * 4);
  L.local_cand = o; o += al256(size_t(nq) * nc * 4);
  L.cand_scores = o; o += al256(size_t(nq) * nc * 4);
  L.part_bytes = hrc_maxsim_workspace_bytes(nc, nq, lq);
  L.part = o; o += al256(L.part_bytes);
  L.rr_keys = o; o += al256(size_t(nq) * nc * 8);
  L.gather2_bytes = hrc_allgather_merge_workspace_bytes(world, nq, nc);
  L.gather2 = o; o += al256(L.gather2_bytes);
  L.fkeys = o; o += al256(size_t(nq) * fk * 8);
  L.pos = o; o += al256(size_t(nq) * fk * 4);
  L.total = o;
  return L;
}

__global__ void f32_to_bf16_rows_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int64_t n) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __float2bfloat16_rn(in[i]);
}

}  // namespace

}  // namespace hrc

using namespace hrc;

#ifdef HRC_EXPERIMENTS
// libhrc_exp.so only: make hrc_comm_enable_p2p fail locally on one rank, to test that every rank then gets the same verdict
static int g_exp_fail_p2p_rank = -1;
extern "C" void hrc_exp_fail_p2p(int rank) { g_exp_fail_p2p_rank = rank; }
#endif

extern "C" {

int hrc_comm_unique_id(void* id_out) {
  if (int rc = load_nccl()) return rc;
  HRC_REQUIRE(id_out != nullptr, "comm_unique_id: null output");
  ncclUniqueId id;
  HRC_CHECK_NCCL(g_nccl.GetUniqueId(&id));
  static_assert(sizeof(id) == HRC_COMM_ID_BYTES, "HRC_COMM_ID_BYTES must equal NCCL_UNIQUE_ID_BYTES");
  memcpy(id_out, &id, sizeof(id));
  return 0;
}

int hrc_comm_init(const void* unique_id, int world, int rank, hrc_comm_t** out) {
  if (int rc = load_nccl()) return rc;
  HRC_REQUIRE(unique_id != nullptr && out != nullptr, "comm_init: null argument");
  HRC_REQUIRE(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, "comm_init: bad world %d / rank %d", world, rank);
  ncclUniqueId id;
  memcpy(&id, unique_id, sizeof(id));
  Comm* c = new Comm();
  c->world = world;
  c->rank = rank;
  cudaGetDevice(&c->device);
  ncclResult_t r = g_nccl.CommInitRank(&c->nccl, world, id, rank);
  if (r != ncclSuccess) {
    set_error("ncclCommInitRank failed: %s", g_nccl.GetErrorString(r));
    delete c;
    return 1;
  }
  // warm the communicator up (NCCL sets up its channels lazily on the first collectives): three tiny all-gathers now,
  // not inside somebody's first searches
  uint64_t* d_tmp = nullptr;
  if (cudaMalloc(reinterpret_cast<void**>(&d_tmp), sizeof(uint64_t) * 128 * (world + 1)) == cudaSuccess) {
    cudaMemset(d_tmp, 0, sizeof(uint64_t) * 128 * (world + 1));
    for (int i = 0; i < 3; ++i) g_nccl.AllGather(d_tmp + 128 * world, d_tmp, 128, ncclUint64, c->nccl, nullptr);
    cudaDeviceSynchronize();
    cudaFree(d_tmp);
  }
  *out = reinterpret_cast<hrc_comm_t*>(c);
  return 0;
}

int hrc_comm_world(const hrc_comm_t* comm) { return comm ? reinterpret_cast<const Comm*>(comm)->world : 0; }
int hrc_comm_rank(const hrc_comm_t* comm) { return comm ? reinterpret_cast<const Comm*>(comm)->rank : -1; }

int hrc_comm_enable_p2p(hrc_comm_t* comm, int max_keys, void* stream) {
  Comm* c = reinterpret_cast<Comm*>(comm);
  HRC_REQUIRE(c != nullptr && max_keys >= 1, "comm_enable_p2p: bad argument");
  if (c->p2p && c->max_keys >= max_keys) return 0;
  HRC_REQUIRE(!c->p2p, "comm_enable_p2p: already enabled with a smaller capacity (%d < %d)", c->max_keys, max_keys);
  // Collective: every rank of the communicator calls this.  A failure on ANY rank (no peer access, IPC not permitted
  // in this container, out of memory) must fail on EVERY rank, or the others would wait for it in the next collective:
  // every rank takes part in both all-gathers whatever happened locally, and the second one carries its verdict.
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t bytes = kFlagBytes + size_t(2) * c->world * size_t(max_keys) * sizeof(uint64_t);
  bool ok = true;
  char why[256] = "";
  auto fail = [&](const char* what, cudaError_t e) {
    if (ok) snprintf(why, sizeof(why), "%s: %s", what, cudaGetErrorString(e));
    ok = false;
    cudaGetLastError();
  };
  cudaError_t e;
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof(mine));
  uint8_t* local = nullptr;
#ifdef HRC_EXPERIMENTS
  if (c->rank == g_exp_fail_p2p_rank) fail("forced by hrc_exp_fail_p2p", cudaErrorNotSupported);
#endif
  if (ok && (e = cudaMalloc(reinterpret_cast<void**>(&local), bytes)) != cudaSuccess) fail("cudaMalloc(receive buffer)", e);
  if (ok && (e = cudaMemsetAsync(local, 0, bytes, st)) != cudaSuccess) fail("cudaMemsetAsync", e);
  if (ok && (e = cudaIpcGetMemHandle(&mine, local)) != cudaSuccess) fail("cudaIpcGetMemHandle", e);
  // scratch for the two all-gathers: world + 1 handles, world + 1 verdicts
  uint8_t* d_tmp = nullptr;
  const size_t hbytes = sizeof(mine) * size_t(c->world + 1), vbytes = sizeof(uint64_t) * size_t(c->world + 1);
  HRC_CHECK_CUDA(cudaMalloc(reinterpret_cast<void**>(&d_tmp), hbytes + vbytes));     // (failing here fails every later call too)
  uint8_t* d_handles = d_tmp;
  uint64_t* d_verdict = reinterpret_cast<uint64_t*>(d_tmp + hbytes);
  HRC_CHECK_CUDA(cudaMemcpyAsync(d_handles + sizeof(mine) * c->world, &mine, sizeof(mine), cudaMemcpyHostToDevice, st));
  HRC_CHECK_NCCL(g_nccl.AllGather(d_handles + sizeof(mine) * c->world, d_handles, sizeof(mine), ncclUint8, c->nccl, st));
  cudaIpcMemHandle_t all[kMaxWorld];
  HRC_CHECK_CUDA(cudaMemcpyAsync(all, d_handles, sizeof(mine) * c->world, cudaMemcpyDeviceToHost, st));
  HRC_CHECK_CUDA(cudaStreamSynchronize(st));
  uint8_t* peer[kMaxWorld] = {};
  for (int r = 0; ok && r < c->world; ++r) {
    if (r == c->rank) { peer[r] = local; continue; }
    void* p = nullptr;
    if ((e = cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess)) != cudaSuccess) fail("cudaIpcOpenMemHandle", e);
    peer[r] = static_cast<uint8_t*>(p);
  }
  uint8_t** d_peer = nullptr;
  if (ok && (e = cudaMalloc(reinterpret_cast<void**>(&d_peer), sizeof(uint8_t*) * kMaxWorld)) != cudaSuccess) fail("cudaMalloc(peer table)", e);
  if (ok && (e = cudaMemcpy(d_peer, peer, sizeof(uint8_t*) * kMaxWorld, cudaMemcpyHostToDevice)) != cudaSuccess) fail("cudaMemcpy(peer table)", e);
  // the verdicts — and the barrier: every rank's buffer is zeroed and mapped before anybody pushes
  const uint64_t verdict = ok ? 1 : 0;
  uint64_t verdicts[kMaxWorld] = {};
  HRC_CHECK_CUDA(cudaMemcpyAsync(d_verdict + c->world, &verdict, sizeof(verdict), cudaMemcpyHostToDevice, st));
  HRC_CHECK_NCCL(g_nccl.AllGather(d_verdict + c->world, d_verdict, 1, ncclUint64, c->nccl, st));
  HRC_CHECK_CUDA(cudaMemcpyAsync(verdicts, d_verdict, sizeof(uint64_t) * c->world, cudaMemcpyDeviceToHost, st));
  HRC_CHECK_CUDA(cudaStreamSynchronize(st));
  cudaFree(d_tmp);
  int first_bad = -1;
  for (int r = 0; r < c->world; ++r)
    if (verdicts[r] != 1 && first_bad < 0) first_bad = r;
  if (first_bad >= 0) {                       // the same answer on every rank: undo, report, leave the NCCL transport usable
    for (int r = 0; r < c->world; ++r)
      if (r != c->rank && peer[r] != nullptr) cudaIpcCloseMemHandle(peer[r]);
    if (d_peer) cudaFree(d_peer);
    if (local) cudaFree(local);
    cudaGetLastError();
    set_error("comm_enable_p2p: peer-memory transport unavailable (first failing rank %d%s%s)", first_bad, ok ? "" : "; here: ",
              ok ? "" : why);
    return 5;
  }
  c->local = local;
  for (int r = 0; r < kMaxWorld; ++r) c->peer[r] = peer[r];
  c->d_peer = d_peer;
  c->max_keys = max_keys;
  c->p2p = true;
  c->seq = 0;
  return 0;
}

int hrc_comm_destroy(hrc_comm_t* comm) {
  Comm* c = reinterpret_cast<Comm*>(comm);
  if (c == nullptr) return 0;
  if (c->p2p) {
    cudaDeviceSynchronize();
    for (int r = 0; r < c->world; ++r)
      if (r != c->rank && c->peer[r] != nullptr) cudaIpcCloseMemHandle(c->peer[r]);
    cudaFree(c->d_peer);
    cudaFree(c->local);
  }
  if (c->nccl != nullptr && g_nccl.CommDestroy != nullptr) g_nccl.CommDestroy(c->nccl);
  delete c;
  return 0;
}

size_t hrc_allgather_merge_workspace_bytes(int world, int n_rows, int k) {
  if (world < 1 || n_rows < 0 || k < 0) return 0;
  return (size_t(world) * size_t(n_rows) * size_t(k) * sizeof(uint64_t) + 255) & ~size_t(255);
}

int hrc_allgather_merge_topk(hrc_comm_t* comm, const uint64_t* d_local_keys, int n_rows, int k, int transport,
                             void* d_workspace, size_t workspace_bytes, uint64_t* d_keys_out, int32_t* d_ids_out,
                             float* d_scores_out, void* stream) {
  Comm* c = reinterpret_cast<Comm*>(comm);
  HRC_REQUIRE(c != nullptr, "allgather_merge: null communicator");
  HRC_REQUIRE(n_rows >= 0 && k >= 0 && k <= HRC_MAX_TOPK, "allgather_merge: bad sizes");
  if (n_rows == 0 || k == 0) return 0;
  HRC_REQUIRE(d_local_keys != nullptr && d_keys_out != nullptr, "allgather_merge: null buffer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int n_keys = n_rows * k;
  if (transport == HRC_TRANSPORT_P2P) {
    HRC_REQUIRE(c->p2p && n_keys <= c->max_keys, "allgather_merge: P2P transport not enabled for %d keys (hrc_comm_enable_p2p)",
                n_keys);
    const uint64_t seq = ++c->seq;
    const int parity = int(seq & 1);
    p2p_push_kernel<<<c->world, 256, 0, st>>>(c->d_peer, d_local_keys, n_keys, c->world, c->rank, c->max_keys, parity, seq);
    count_launch();
    HRC_CHECK_CUDA(cudaGetLastError());
    // the merge kernel acquires the world flags of this parity (>= seq), then reads slots [parity][0..world)
    const uint64_t* flags = reinterpret_cast<const uint64_t*>(c->local) + parity * kMaxWorld;
    const uint64_t* slots = reinterpret_cast<const uint64_t*>(c->local + slot_offset(parity, 0, c->world, c->max_keys));
    return launch_topk_merge_parts(slots, c->world, c->max_keys, n_rows, k, d_keys_out, st, d_ids_out, d_scores_out, flags,
                                   seq, c->world, 0);
  }
  HRC_REQUIRE(transport == HRC_TRANSPORT_NCCL, "allgather_merge: unknown transport %d", transport);
  const size_t need = hrc_allgather_merge_workspace_bytes(c->world, n_rows, k);
  HRC_REQUIRE(d_workspace != nullptr && workspace_bytes >= need, "allgather_merge: workspace too small (%zu < %zu)",
              workspace_bytes, need);
  uint64_t* gathered = static_cast<uint64_t*>(d_workspace);                  // [world][n_rows][k]
  HRC_CHECK_NCCL(g_nccl.AllGather(d_local_keys, gathered, size_t(n_keys), ncclUint64, c->nccl, st));
  return launch_topk_merge_parts(gathered, c->world, n_keys, n_rows, k, d_keys_out, st, d_ids_out, d_scores_out, nullptr, 0,
                                 0, 0);
}

size_t hrc_sharded_search_workspace_bytes(int world, int64_t n_docs, int64_t total_tokens, int n_queries, int lq, int k,
                                          int path) {
  if (world < 1 || n_docs < 0 || total_tokens < 0 || n_queries < 0 || lq < 1 || k < 0) return 0;
  return sharded_search_layout(world, n_docs, total_tokens, n_queries, lq, k, path).total;
}

int hrc_sharded_search(hrc_comm_t* comm, int transport, const void* d_tokens, const int64_t* d_offsets, int64_t n_docs,
                       int64_t total_tokens, const void* d_queries, int n_queries, int lq, int k, int32_t id_base,
                       void* d_workspace, size_t workspace_bytes, uint64_t* d_keys_out, int32_t* d_ids_out,
                       float* d_scores_out, int path, void* stream) {
  Comm* c = reinterpret_cast<Comm*>(comm);
  HRC_REQUIRE(c != nullptr, "sharded_search: null communicator");
  HRC_REQUIRE(n_queries >= 0 && lq >= 1 && k >= 0 && k <= HRC_MAX_TOPK, "sharded_search: bad sizes");
  if (n_queries == 0 || k == 0) return 0;
  const ShardedSearchLayout L = sharded_search_layout(c->world, n_docs, total_tokens, n_queries, lq, k, path);
  HRC_REQUIRE(d_workspace != nullptr && workspace_bytes >= L.total, "sharded_search: workspace too small (%zu < %zu)",
              workspace_bytes, L.total);
  HRC_REQUIRE((reinterpret_cast<uintptr_t>(d_workspace) & 255) == 0, "sharded_search: workspace must be 256-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(d_workspace);
  uint64_t* local = reinterpret_cast<uint64_t*>(ws + L.local);
  const int k_local = int(n_docs < k ? n_docs : k);       // a shard with fewer than k documents pads with empty slots
  if (transport == HRC_TRANSPORT_P2P && c->p2p && k_local == k &&
      search_exchange_supported(n_docs, total_tokens, n_queries, lq, k, path, c->world, c->max_keys)) {
    // ONE query over peer memory: the search's final selection kernel stores this GPU's top-k into every rank's slot,
    // waits for the others' and emits the global top-k itself — two launches, like a single-GPU search.  A rank that
    // takes the separate push + merge kernels below (small shard, other route) speaks the same slot / flag protocol.
    KeyExchange x;
    x.peers = c->d_peer;
    x.local = c->local;
    x.world = c->world;
    x.my_rank = c->rank;
    x.max_keys = c->max_keys;
    x.seq = ++c->seq;
    x.parity = int(x.seq & 1);
    x.watchdog_ns = get_watchdog_ns();
    return search_with_exchange(d_tokens, d_offsets, n_docs, total_tokens, d_queries, n_queries, lq, k, id_base, ws + L.search,
                                L.search_bytes, d_keys_out, d_ids_out, d_scores_out, path, stream, &x);
  }
  if (k_local < k) HRC_CHECK_CUDA(cudaMemsetAsync(local, 0, size_t(n_queries) * k * sizeof(uint64_t), st));
  if (k_local > 0) {
    // (a shard smaller than k writes rows of k_local keys; re-spread them to rows of k below)
    uint64_t* dst = k_local == k ? local : reinterpret_cast<uint64_t*>(ws + L.gather);
    if (int rc = hrc_search(d_tokens, d_offsets, n_docs, total_tokens, d_queries, n_queries, lq, k_local, id_base, ws + L.search,
                            L.search_bytes, dst, nullptr, nullptr, path, stream))
      return rc;
    if (k_local != k)
      HRC_CHECK_CUDA(cudaMemcpy2DAsync(local, size_t(k) * 8, dst, size_t(k_local) * 8, size_t(k_local) * 8, n_queries,
                                       cudaMemcpyDeviceToDevice, st));
  }
  return hrc_allgather_merge_topk(comm, local, n_queries, k, transport, ws + L.gather, L.gather_bytes, d_keys_out, d_ids_out,
                                  d_scores_out, stream);
}

size_t hrc_sharded_search_host_workspace_bytes(int world, int64_t n_docs, int64_t total_tokens, int n_queries, int lq,
                                               int k, int path) {
  if (world < 1 || n_docs < 0 || total_tokens < 0 || n_queries < 0 || lq < 1 || k < 0) return 0;
  return sharded_host_layout(world, n_docs, total_tokens, n_queries, lq, k, path).total;
}

int hrc_sharded_search_host(hrc_comm_t* comm, int transport, const void* d_tokens, const int64_t* d_offsets,
                            int64_t n_docs, int64_t total_tokens, const float* h_queries, int n_queries, int lq, int k,
                            int32_t id_base, void* d_workspace, size_t workspace_bytes, int32_t* h_ids_out,
                            float* h_scores_out, int path, void* stream) {
  Comm* c = reinterpret_cast<Comm*>(comm);
  HRC_REQUIRE(c != nullptr, "sharded_search_host: null communicator");
  HRC_REQUIRE(n_queries >= 0 && lq >= 1 && k >= 0 && k <= HRC_MAX_TOPK, "sharded_search_host: bad sizes");
  if (n_queries == 0 || k == 0) return 0;
  HRC_REQUIRE(h_queries != nullptr && h_ids_out != nullptr && h_scores_out != nullptr, "sharded_search_host: null buffer");
  const ShardedHostLayout L = sharded_host_layout(c->world, n_docs, total_tokens, n_queries, lq, k, path);
  HRC_REQUIRE(d_workspace != nullptr && workspace_bytes >= L.total, "sharded_search_host: workspace too small (%zu < %zu)",
              workspace_bytes, L.total);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(d_workspace);
  float* q32 = reinterpret_cast<float*>(ws + L.q32);
  __nv_bfloat16* q16 = reinterpret_cast<__nv_bfloat16*>(ws + L.q16);
  const int64_t n = int64_t(n_queries) * lq * HRC_DIM;
  HRC_CHECK_CUDA(cudaMemcpyAsync(q32, h_queries, size_t(n) * sizeof(float), cudaMemcpyHostToDevice, st));
  f32_to_bf16_rows_kernel<<<unsigned((n + 255) / 256), 256, 0, st>>>(q32, q16, n);
  count_launch();
  const size_t inner = sharded_search_layout(c->world, n_docs, total_tokens, n_queries, lq, k, path).total;
  uint64_t* keys = reinterpret_cast<uint64_t*>(ws + L.inner + inner);
  int32_t* d_ids = reinterpret_cast<int32_t*>(ws + L.ids);
  float* d_sc = reinterpret_cast<float*>(ws + L.scores);
  if (int rc = hrc_sharded_search(comm, transport, d_tokens, d_offsets, n_docs, total_tokens, q16, n_queries, lq, k, id_base,
                                  ws + L.inner, inner, keys, d_ids, d_sc, path, stream))
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

size_t hrc_sharded_hybrid_workspace_bytes(int world, int64_t n_docs, int64_t total_tokens, int n_queries, int lq,
                                          int colbert_k, int n_candidates, int final_k, int path) {
  if (world < 1 || n_docs < 0 || total_tokens < 0 || n_queries < 0 || lq < 1 || colbert_k < 0 || n_candidates < 0 || final_k < 0)
    return 0;
  return sharded_hybrid_layout(world, n_docs, total_tokens, n_queries, lq, colbert_k, n_candidates, final_k, path).total;
}

int hrc_sharded_hybrid_retrieve(hrc_comm_t* comm, int transport, const void* d_tokens, const int64_t* d_offsets,
                                int64_t n_docs, int64_t total_tokens, int64_t n_docs_global, const void* d_queries,
                                int n_queries, int lq, const int32_t* d_bm25_ids, int n_bm25, int colbert_k, int rrf_k,
                                int n_candidates, int final_k, int32_t id_base, void* d_workspace, size_t workspace_bytes,
                                int32_t* d_ids_out, float* d_scores_out, int path, void* stream) {
  Comm* c = reinterpret_cast<Comm*>(comm);
  HRC_REQUIRE(c != nullptr, "sharded_hybrid: null communicator");
  HRC_REQUIRE(n_queries >= 0 && lq >= 1 && n_bm25 >= 0 && colbert_k >= 1 && colbert_k <= n_docs_global && n_candidates >= 1 &&
                  final_k >= 1 && final_k <= n_candidates && colbert_k <= HRC_MAX_TOPK,
              "sharded_hybrid: need 1 <= colbert_k <= n_docs_global and 1 <= final_k <= n_candidates");
  if (n_queries == 0) return 0;
  HRC_REQUIRE(d_ids_out != nullptr && d_scores_out != nullptr && (n_bm25 == 0 || d_bm25_ids != nullptr), "sharded_hybrid: null buffer");
  const ShardedHybridLayout L =
      sharded_hybrid_layout(c->world, n_docs, total_tokens, n_queries, lq, colbert_k, n_candidates, final_k, path);
  HRC_REQUIRE(d_workspace != nullptr && workspace_bytes >= L.total, "sharded_hybrid: workspace too small (%zu < %zu)",
              workspace_bytes, L.total);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(d_workspace);
  uint64_t* gkeys = reinterpret_cast<uint64_t*>(ws + L.gkeys);
  int32_t* col_ids = reinterpret_cast<int32_t*>(ws + L.col_ids);
  int32_t* fused = reinterpret_cast<int32_t*>(ws + L.fused);
  int32_t* local_cand = reinterpret_cast<int32_t*>(ws + L.local_cand);
  float* cand_scores = reinterpret_cast<float*>(ws + L.cand_scores);
  uint64_t* rr_keys = reinterpret_cast<uint64_t*>(ws + L.rr_keys);
  uint64_t* fkeys = reinterpret_cast<uint64_t*>(ws + L.fkeys);
  // stage 2 (:908-911): the GLOBAL ColBERT list — local top-k, exchange, merge; ids unpacked by the merge kernel
  const size_t inner = sharded_search_layout(c->world, n_docs, total_tokens, n_queries, lq, colbert_k, path).total;
  HRC_REQUIRE(inner <= L.gkeys, "sharded_hybrid: internal layout error");
  if (int rc = hrc_sharded_search(comm, transport, d_tokens, d_offsets, n_docs, total_tokens, d_queries, n_queries, lq,
                                  colbert_k, id_base, ws, inner, gkeys, col_ids, nullptr, path, stream))
    return rc;
  // stage 3 (:914-916): RRF on global ids; every rank computes the same fusion
  if (int rc = launch_rrf(d_bm25_ids, n_bm25, col_ids, colbert_k, n_queries, rrf_k, n_candidates, fused,
                          reinterpret_cast<double*>(ws + L.fused_scores), reinterpret_cast<int32_t*>(ws + L.counts), st))
    return rc;
  // stage 5 (:926-929): every rank scores the candidates it owns, the (score, position) keys are exchanged and merged
  const int64_t n_fused = int64_t(n_queries) * n_candidates;
  localize_ids_kernel<<<unsigned((n_fused + 255) / 256), 256, 0, st>>>(fused, n_fused, id_base, n_docs, local_cand);
  count_launch();
  if (total_tokens > 0) {
    if (int rc = hrc_maxsim_scores_ids(d_tokens, d_offsets, n_docs, total_tokens, local_cand, n_candidates, d_queries, n_queries,
                                       lq, cand_scores, path, ws + L.part, L.part_bytes, stream))
      return rc;
  }
  owned_rerank_keys_kernel<<<unsigned((n_fused + 255) / 256), 256, 0, st>>>(cand_scores, fused, n_candidates, n_fused, id_base,
                                                                           total_tokens > 0 ? n_docs : 0, n_docs_global,
                                                                           c->rank, rr_keys);
  count_launch();
  HRC_CHECK_CUDA(cudaGetLastError());
  // merge to final_k: hrc_allgather_merge_topk merges world x n_candidates keys per row down to n_candidates; the top
  // final_k of that sorted list are the result
  if (int rc = hrc_allgather_merge_topk(comm, rr_keys, n_queries, n_candidates, transport, ws + L.gather2, L.gather2_bytes,
                                        rr_keys, nullptr, nullptr, stream))
    return rc;
  if (final_k == n_candidates) {
    return launch_rerank_unpack(rr_keys, final_k, n_queries, fused, n_candidates, reinterpret_cast<int32_t*>(ws + L.pos),
                                d_ids_out, d_scores_out, st);
  }
  HRC_CHECK_CUDA(cudaMemcpy2DAsync(fkeys, size_t(final_k) * 8, rr_keys, size_t(n_candidates) * 8, size_t(final_k) * 8,
                                   n_queries, cudaMemcpyDeviceToDevice, st));
  return launch_rerank_unpack(fkeys, final_k, n_queries, fused, n_candidates, reinterpret_cast<int32_t*>(ws + L.pos), d_ids_out,
                              d_scores_out, st);
}

}  // extern "C"
