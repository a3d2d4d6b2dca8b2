// maxsim_simt.cu — CUDA-core MaxSim path: coalesced 16-byte loads + warp-shuffle reductions.
//
// Used when the problem is too small (or the query too long: lq > 32) for the tcgen05 contraction
// to pay, and as the on-GPU cross-check of the tensor-core path.  One CTA scores one
// (query, document) pair:
//   - the query's tokens are staged in shared memory as fp32, 32 tokens at a time;
//   - each half-warp owns one document token per iteration: lane s (0..15) loads dims [8s, 8s+8)
//     with one 16-byte streaming load, so a warp instruction reads two whole 256-byte token rows;
//   - every lane forms 32 partial dot products (one per staged query token) over its 8 dims, and a
//     halving butterfly over the 16 lanes (16+8+4+2 = 30 shuffles instead of 32 x 4) leaves lane s
//     holding the complete dot products for query tokens 2s and 2s+1;
//   - running maxima stay in registers; halves, warps and finally query tokens are combined once
//     per document.
// Reference semantics: local_rag_complete.py:807-812 (docstring of _maxsim_score), sum-reduced over
// query tokens as BASELINE.json's north_star states.
#include "hrc_common.cuh"

namespace hrc {

namespace {

constexpr int kSimtThreads = 128;
constexpr int kSimtWarps = kSimtThreads / 32;
constexpr int kQChunk = 32;

// One butterfly level: n live values per lane -> n/2, exchanging with lane ^ off.
template <int N, int OFF>
__device__ __forceinline__ void butterfly_halve(float (&a)[32], bool upper) {
#pragma unroll
  for (int i = 0; i < N / 2; ++i) {
    const float keep = upper ? a[i + N / 2] : a[i];
    const float send = upper ? a[i] : a[i + N / 2];
    a[i] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
  }
}

__global__ void __launch_bounds__(kSimtThreads)
maxsim_simt_kernel(const __nv_bfloat16* __restrict__ tokens, const int64_t* __restrict__ offsets,
                   int64_t n_docs, const int32_t* __restrict__ cand_ids, int n_items,
                   const __nv_bfloat16* __restrict__ queries, int lq, float* __restrict__ scores) {
  // Qs[q][h][s] (float4): dims 8s+4h .. 8s+4h+3 of staged query token q.  A half-warp reads 256
  // contiguous bytes per access (conflict-free); the two halves read the same address (broadcast).
  __shared__ float4 Qs[kQChunk * 32];
  __shared__ float warp_m[kSimtWarps][kQChunk];

  const int item = blockIdx.x;
  const int q_idx = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int half = lane >> 4;
  const int s = lane & 15;

  int64_t doc;
  if (cand_ids != nullptr) {
    doc = cand_ids[int64_t(q_idx) * n_items + item];
  } else {
    doc = item;
  }
  float* out = scores + int64_t(q_idx) * n_items + item;
  if (doc < 0 || doc >= n_docs) {
    if (threadIdx.x == 0) *out = -INFINITY;
    return;
  }
  const int64_t tok0 = offsets[doc];
  const int len = int(offsets[doc + 1] - tok0);
  const uint4* doc_rows = reinterpret_cast<const uint4*>(tokens + tok0 * HRC_DIM);
  const __nv_bfloat16* qbase = queries + int64_t(q_idx) * lq * HRC_DIM;

  float total = 0.f;  // meaningful on thread 0 only
  for (int q0 = 0; q0 < lq; q0 += kQChunk) {
    const int nq = min(kQChunk, lq - q0);
    __syncthreads();  // previous chunk's readers are done with Qs / warp_m
    // stage the query chunk: thread -> (q, 4-dim group)
    for (int e = threadIdx.x; e < kQChunk * 32; e += kSimtThreads) {
      const int q = e >> 5;
      const int g = e & 31;  // dims 4g..4g+3
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (q < nq) {
        const uint2 raw = *reinterpret_cast<const uint2*>(qbase + int64_t(q0 + q) * HRC_DIM + 4 * g);
        v = make_float4(bf16lo_to_f32(raw.x), bf16hi_to_f32(raw.x), bf16lo_to_f32(raw.y),
                        bf16hi_to_f32(raw.y));
      }
      Qs[q * 32 + (g & 1) * 16 + (g >> 1)] = v;
    }
    __syncthreads();

    float m0 = -INFINITY, m1 = -INFINITY;  // running max for query tokens 2s, 2s+1 (this half's tokens)
    for (int p = warp; 2 * p < len; p += kSimtWarps) {
      const int t = 2 * p + half;
      const bool valid = t < len;
      uint4 raw = make_uint4(0u, 0u, 0u, 0u);
      if (valid) raw = ldg_stream16(doc_rows + int64_t(t) * 16 + s);
      const float d0 = bf16lo_to_f32(raw.x), d1 = bf16hi_to_f32(raw.x);
      const float d2 = bf16lo_to_f32(raw.y), d3 = bf16hi_to_f32(raw.y);
      const float d4 = bf16lo_to_f32(raw.z), d5 = bf16hi_to_f32(raw.z);
      const float d6 = bf16lo_to_f32(raw.w), d7 = bf16hi_to_f32(raw.w);
      float a[32];
#pragma unroll
      for (int q = 0; q < kQChunk; ++q) {
        const float4 qa = Qs[q * 32 + s];
        const float4 qb = Qs[q * 32 + 16 + s];
        float acc = d0 * qa.x;
        acc = fmaf(d1, qa.y, acc);
        acc = fmaf(d2, qa.z, acc);
        acc = fmaf(d3, qa.w, acc);
        acc = fmaf(d4, qb.x, acc);
        acc = fmaf(d5, qb.y, acc);
        acc = fmaf(d6, qb.z, acc);
        acc = fmaf(d7, qb.w, acc);
        a[q] = acc;
      }
      butterfly_halve<32, 8>(a, (s & 8) != 0);
      butterfly_halve<16, 4>(a, (s & 4) != 0);
      butterfly_halve<8, 2>(a, (s & 2) != 0);
      butterfly_halve<4, 1>(a, (s & 1) != 0);
      if (valid) {
        m0 = fmaxf(m0, a[0]);
        m1 = fmaxf(m1, a[1]);
      }
    }
    // halves -> warp
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 16));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 16));
    if (half == 0) {
      warp_m[warp][2 * s] = m0;
      warp_m[warp][2 * s + 1] = m1;
    }
    __syncthreads();
    if (warp == 0) {
      float m = warp_m[0][lane];
#pragma unroll
      for (int w = 1; w < kSimtWarps; ++w) m = fmaxf(m, warp_m[w][lane]);
      const float contrib = (lane < nq) ? m : 0.f;
      total += warp_sum(contrib);
    }
  }
  if (threadIdx.x == 0) *out = total;
}

}  // namespace

int launch_maxsim_simt(const void* d_tokens, const int64_t* d_offsets, int64_t n_docs,
                       const int32_t* d_cand_ids, int64_t n_items, const void* d_queries,
                       int n_queries, int lq, float* d_scores, cudaStream_t stream) {
  if (n_items == 0 || n_queries == 0) return 0;
  HRC_REQUIRE(n_items <= 0x7fffffffLL, "simt path: too many items (%lld)", (long long)n_items);
  HRC_REQUIRE(n_queries <= 65535, "simt path: too many queries per launch (%d)", n_queries);
  dim3 grid((unsigned)n_items, (unsigned)n_queries);
  trace_begin(stream);
  maxsim_simt_kernel<<<grid, kSimtThreads, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(d_tokens), d_offsets, n_docs, d_cand_ids, int(n_items),
      static_cast<const __nv_bfloat16*>(d_queries), lq, d_scores);
  trace_end(stream);
  count_launch();
  HRC_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace hrc
