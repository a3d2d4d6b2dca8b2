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
#include <nccl.h>

#include <cstring>

#include "hrc_common.cuh"

namespace hrc {

int launch_topk_merge_parts(const uint64_t*, int, int, int, int, uint64_t*, cudaStream_t, int32_t*, float*, const uint64_t*,
                            uint64_t, int, uint64_t);

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

constexpr int kMaxWorld = 16;

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

constexpr size_t kFlagBytes = 2 * kMaxWorld * sizeof(uint64_t);
__host__ __device__ inline size_t slot_offset(int parity, int src_rank, int world, int max_keys) {
  return kFlagBytes + (size_t(parity) * world + src_rank) * size_t(max_keys) * sizeof(uint64_t);
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

}  // namespace

}  // namespace hrc

using namespace hrc;

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
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t bytes = kFlagBytes + size_t(2) * c->world * size_t(max_keys) * sizeof(uint64_t);
  HRC_CHECK_CUDA(cudaMalloc(reinterpret_cast<void**>(&c->local), bytes));     // set-up time, never on the query path
  HRC_CHECK_CUDA(cudaMemsetAsync(c->local, 0, bytes, st));
  // exchange the IPC handles of the receive buffers with the communicator itself
  cudaIpcMemHandle_t mine;
  HRC_CHECK_CUDA(cudaIpcGetMemHandle(&mine, c->local));
  uint8_t* d_handles = nullptr;
  HRC_CHECK_CUDA(cudaMalloc(reinterpret_cast<void**>(&d_handles), sizeof(mine) * (c->world + 1)));
  HRC_CHECK_CUDA(cudaMemcpyAsync(d_handles + sizeof(mine) * c->world, &mine, sizeof(mine), cudaMemcpyHostToDevice, st));
  HRC_CHECK_NCCL(g_nccl.AllGather(d_handles + sizeof(mine) * c->world, d_handles, sizeof(mine), ncclUint8, c->nccl, st));
  cudaIpcMemHandle_t all[kMaxWorld];
  HRC_CHECK_CUDA(cudaMemcpyAsync(all, d_handles, sizeof(mine) * c->world, cudaMemcpyDeviceToHost, st));
  HRC_CHECK_CUDA(cudaStreamSynchronize(st));
  cudaFree(d_handles);
  for (int r = 0; r < c->world; ++r) {
    if (r == c->rank) { c->peer[r] = c->local; continue; }
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      set_error("comm_enable_p2p: cudaIpcOpenMemHandle(rank %d) failed: %s", r, cudaGetErrorString(e));
      return 1;
    }
    c->peer[r] = static_cast<uint8_t*>(p);
  }
  HRC_CHECK_CUDA(cudaMalloc(reinterpret_cast<void**>(&c->d_peer), sizeof(uint8_t*) * kMaxWorld));
  HRC_CHECK_CUDA(cudaMemcpy(c->d_peer, c->peer, sizeof(uint8_t*) * kMaxWorld, cudaMemcpyHostToDevice));
  // every rank's buffer is zeroed and mapped before anybody pushes: one more (tiny) collective as the barrier
  uint64_t* d_tmp = nullptr;
  HRC_CHECK_CUDA(cudaMalloc(reinterpret_cast<void**>(&d_tmp), sizeof(uint64_t) * (c->world + 1)));
  HRC_CHECK_NCCL(g_nccl.AllGather(d_tmp + c->world, d_tmp, 1, ncclUint64, c->nccl, st));
  HRC_CHECK_CUDA(cudaStreamSynchronize(st));
  cudaFree(d_tmp);
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

}  // extern "C"
