// rrf.cu — reciprocal-rank fusion of two ranked id lists, bit-compatible with the reference.
//
// Replaces HybridRetriever._reciprocal_rank_fusion (local_rag_complete.py:960-978) and the
// [:50] slice at :916.  The reference accumulates Python floats (IEEE fp64) in a dict:
//     for rank, r in enumerate(list_a, 1): scores[id] = scores.get(id, 0) + 1 / (k + rank)
//     for rank, r in enumerate(list_b, 1): scores[id] = scores.get(id, 0) + 1 / (k + rank)
//     sorted(scores.items(), key=score, reverse=True)          # stable: ties keep insertion order
// so a fused score is the left-to-right fp64 sum of 1/(k+rank) over the id's occurrences in
// concat(a, b), and equal scores keep first-occurrence order (an fp32 or re-ordered
// version flips adjacent results).  One CTA per row; positions are compared all-pairs in shared
// memory (lists are <= 16384 long, 200 in the reference), which keeps the arithmetic order exact.
#include "hrc_common.cuh"

namespace hrc {

namespace {

constexpr int kRrfThreads = 256;
constexpr int kRrfMax = 16384;   // 13 bytes of shared memory per entry

__global__ void __launch_bounds__(kRrfThreads)
rrf_fuse_kernel(const int32_t* __restrict__ ids_a, int n_a, const int32_t* __restrict__ ids_b, int n_b,
                int rrf_k, int top_n, int32_t* __restrict__ ids_out, double* __restrict__ scores_out,
                int32_t* __restrict__ counts_out) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int len = n_a + n_b;
  double* score = reinterpret_cast<double*>(smem);              // [len] fused score of a first occurrence
  int32_t* ids = reinterpret_cast<int32_t*>(score + len);       // [len]
  uint8_t* first = reinterpret_cast<uint8_t*>(ids + len);       // [len] 1 = first occurrence of a valid id
  __shared__ int n_unique;

  const int64_t row = blockIdx.x;
  if (threadIdx.x == 0) n_unique = 0;
  for (int p = threadIdx.x; p < len; p += kRrfThreads)
    ids[p] = (p < n_a) ? ids_a[row * n_a + p] : ids_b[row * n_b + (p - n_a)];
  for (int i = threadIdx.x; i < top_n; i += kRrfThreads) {
    ids_out[row * top_n + i] = -1;
    scores_out[row * top_n + i] = 0.0;
  }
  __syncthreads();

  for (int p = threadIdx.x; p < len; p += kRrfThreads) {
    const int32_t id = ids[p];
    bool is_first = id >= 0;
    double s = 0.0;
    if (is_first) {
      for (int p2 = 0; p2 < len; ++p2) {
        if (ids[p2] == id) {
          if (p2 < p) { is_first = false; break; }
          const int rank = (p2 < n_a) ? (p2 + 1) : (p2 - n_a + 1);
          s = s + 1.0 / double(rrf_k + rank);
        }
      }
    }
    first[p] = is_first ? 1 : 0;
    score[p] = s;
    if (is_first) atomicAdd(&n_unique, 1);
  }
  __syncthreads();

  for (int p = threadIdx.x; p < len; p += kRrfThreads) {
    if (!first[p]) continue;
    const double s = score[p];
    int pos = 0;
    for (int p2 = 0; p2 < len; ++p2) {
      if (!first[p2]) continue;
      const double s2 = score[p2];
      pos += (s2 > s || (s2 == s && p2 < p)) ? 1 : 0;
    }
    if (pos < top_n) {
      ids_out[row * top_n + pos] = ids[p];
      scores_out[row * top_n + pos] = s;
    }
  }
  if (counts_out != nullptr && threadIdx.x == 0) counts_out[row] = n_unique;
}

}  // namespace

int launch_rrf(const int32_t* d_ids_a, int n_a, const int32_t* d_ids_b, int n_b, int n_rows, int rrf_k, int top_n,
               int32_t* d_ids_out, double* d_scores_out, int32_t* d_counts_out, cudaStream_t stream) {
  if (n_rows == 0 || top_n == 0) return 0;
  HRC_REQUIRE(n_a >= 0 && n_b >= 0 && n_a + n_b <= kRrfMax, "rrf: n_a + n_b = %d exceeds %d", n_a + n_b, kRrfMax);
  HRC_REQUIRE(rrf_k + 1 > 0, "rrf: k=%d must keep k + rank positive", rrf_k);
  const int len = n_a + n_b;
  const size_t smem = size_t(len) * (8 + 4 + 1) + 16;
  static PerDeviceOnce once;
  int dev;
  if (once.pending(&dev)) {
    HRC_CHECK_CUDA(cudaFuncSetAttribute(rrf_fuse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        kRrfMax * 13 + 16));
    once.mark(dev);
  }
  rrf_fuse_kernel<<<n_rows, kRrfThreads, smem, stream>>>(d_ids_a, n_a, d_ids_b, n_b, rrf_k, top_n, d_ids_out,
                                                         d_scores_out, d_counts_out);
  count_launch();
  HRC_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace hrc
