// meanpool.cu — the reference's `_maxsim_score` EXACTLY AS CODED (local_rag_complete.py:821-829):
//     query_vec = query_embedding.mean(dim=1); doc_vec = doc_embeddings.mean(dim=1)
//     scores    = cosine_similarity(query_vec.unsqueeze(1), doc_vec.unsqueeze(0), dim=2)
// i.e. the cosine of the mean-pooled token vectors — not MaxSim.  It exists so that the reference's
// literal behaviour is reproducible on the GPU and so that ONE device path is pinned by vectors the unmodified
// reference produced (tests/golden/literal_maxsim.npz).  The retriever uses it only with
// RAGConfig.score_mode == "reference_literal".
//
// HBM-bound: every document token is read once (256 B), fp32 accumulation.  One warp per document: a half-warp
// reads one 256-byte token row per instruction (lane s holds dims [8s, 8s+8)), four rows in flight per half-warp;
// the two halves are combined with one shuffle, the document mean is dotted with every query mean (shared memory).
#include "hrc_common.cuh"

namespace hrc {

namespace {

constexpr int kMpThreads = 256;
constexpr int kMpWarps = kMpThreads / 32;
constexpr int kMpMaxQueries = 64;       // query means held in shared memory per launch
constexpr float kCosEps = 1e-8f;        // torch.nn.functional.cosine_similarity default eps

__device__ __forceinline__ void add_row(float (&acc)[8], const uint4 r) {
  acc[0] += bf16lo_to_f32(r.x); acc[1] += bf16hi_to_f32(r.x);
  acc[2] += bf16lo_to_f32(r.y); acc[3] += bf16hi_to_f32(r.y);
  acc[4] += bf16lo_to_f32(r.z); acc[5] += bf16hi_to_f32(r.z);
  acc[6] += bf16lo_to_f32(r.w); acc[7] += bf16hi_to_f32(r.w);
}

__global__ void __launch_bounds__(kMpThreads)
meanpool_cosine_kernel(const __nv_bfloat16* __restrict__ tokens, const int64_t* __restrict__ offsets, int64_t n_docs,
                       const __nv_bfloat16* __restrict__ queries, int n_queries, int lq, float* __restrict__ scores,
                       int64_t score_stride) {
  __shared__ float qmean[kMpMaxQueries][HRC_DIM];
  __shared__ float qnorm[kMpMaxQueries];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int half = lane >> 4;
  const int s = lane & 15;

  // query means (:821): thread -> (query, dim)
  for (int e = threadIdx.x; e < n_queries * HRC_DIM; e += kMpThreads) {
    const int q = e / HRC_DIM, d = e % HRC_DIM;
    float a = 0.f;
    for (int t = 0; t < lq; ++t) a += __bfloat162float(queries[(int64_t(q) * lq + t) * HRC_DIM + d]);
    qmean[q][d] = a / float(lq);
  }
  __syncthreads();
  for (int q = warp; q < n_queries; q += kMpWarps) {
    float a = 0.f;
    for (int d = lane; d < HRC_DIM; d += 32) a += qmean[q][d] * qmean[q][d];
    a = warp_sum(a);
    if (lane == 0) qnorm[q] = fmaxf(sqrtf(a), kCosEps);
  }
  __syncthreads();

  const int64_t warps_total = int64_t(gridDim.x) * kMpWarps;
  for (int64_t doc = int64_t(blockIdx.x) * kMpWarps + warp; doc < n_docs; doc += warps_total) {
    const int64_t t0 = offsets[doc];
    const int len = int(offsets[doc + 1] - t0);
    const uint4* rows = reinterpret_cast<const uint4*>(tokens + t0 * HRC_DIM);   // 16 uint4 per token row
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    int t = half;
    for (; t + 6 < len; t += 8) {        // four rows in flight per half-warp
      const uint4 r0 = ldg_stream16(rows + int64_t(t) * 16 + s);
      const uint4 r1 = ldg_stream16(rows + int64_t(t + 2) * 16 + s);
      const uint4 r2 = ldg_stream16(rows + int64_t(t + 4) * 16 + s);
      const uint4 r3 = ldg_stream16(rows + int64_t(t + 6) * 16 + s);
      add_row(acc, r0); add_row(acc, r1); add_row(acc, r2); add_row(acc, r3);
    }
    for (; t < len; t += 2) add_row(acc, ldg_stream16(rows + int64_t(t) * 16 + s));
    const float inv = 1.f / float(len);   // len == 0: 0 * inf = NaN, like torch's mean over an empty axis
    float nrm = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 16);     // both halves now hold the sum of dims [8s, 8s+8)
      acc[i] *= inv;                                           // doc_vec (:822)
      nrm += acc[i] * acc[i];
    }
    // reduce over the 16 lanes of a half (the halves hold the same values)
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) nrm += __shfl_xor_sync(0xffffffffu, nrm, o);
    const float dnorm = fmaxf(sqrtf(nrm), kCosEps);
    for (int q = 0; q < n_queries; ++q) {
      float dot = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) dot += acc[i] * qmean[q][8 * s + i];
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
      if (lane == 0) scores[int64_t(q) * score_stride + doc] = dot / (dnorm * qnorm[q]);   // (:825-829)
    }
  }
}

}  // namespace

int launch_meanpool_cosine(const void* d_tokens, const int64_t* d_offsets, int64_t n_docs, const void* d_queries,
                           int n_queries, int lq, float* d_scores, cudaStream_t stream) {
  if (n_docs == 0 || n_queries == 0) return 0;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t want = (n_docs + kMpWarps - 1) / kMpWarps;
  const int grid = int(want < int64_t(sms) * 8 ? want : int64_t(sms) * 8);   // 8 CTAs x 8 warps resident per SM
  for (int q0 = 0; q0 < n_queries; q0 += kMpMaxQueries) {
    const int nq = n_queries - q0 < kMpMaxQueries ? n_queries - q0 : kMpMaxQueries;
    meanpool_cosine_kernel<<<grid, kMpThreads, 0, stream>>>(
        static_cast<const __nv_bfloat16*>(d_tokens), d_offsets, n_docs,
        static_cast<const __nv_bfloat16*>(d_queries) + size_t(q0) * lq * HRC_DIM, nq, lq,
        d_scores + size_t(q0) * size_t(n_docs), n_docs);
    count_launch();
    HRC_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}

}  // namespace hrc
