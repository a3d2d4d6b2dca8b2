// validate.cu — integrity check of a packed store before it is searched (hrc_store_validate).
// The scoring kernels trust the CSR offsets (TMA never reads outside the token buffer, but a broken offsets array
// silently scores the wrong tokens) and the finiteness of the token rows (one NaN row poisons the max of its
// document).  An index file read from disk is checked once, at load time: one pass over the offsets, optionally
// one streaming pass over the tokens (HBM-bound, 32.8 GB in ~4.6 ms).
#include "hrc_common.cuh"

namespace hrc {
namespace {

// report[0] = offsets entries that break the CSR contract, report[1] = n_docs + 1 - (first such entry) (0 = none),
// report[2] = token VALUES that are NaN or +-inf, report[3] = longest document (tokens)
__global__ void __launch_bounds__(256)
validate_offsets_kernel(const int64_t* __restrict__ off, int64_t n_docs, int64_t total_tokens,
                        unsigned long long* __restrict__ report) {
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  unsigned long long bad = 0, first = 0, longest = 0;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i <= n_docs; i += stride) {
    const int64_t o = off[i];
    bool ok = o >= 0 && o <= total_tokens;
    if (i == 0) ok = ok && o == 0;
    if (i == n_docs) ok = ok && o == total_tokens;
    if (i < n_docs) {
      const int64_t nxt = off[i + 1];
      ok = ok && nxt >= o;
      if (nxt >= o && (unsigned long long)(nxt - o) > longest) longest = (unsigned long long)(nxt - o);
    }
    if (!ok) {
      ++bad;
      const unsigned long long tag = (unsigned long long)(n_docs + 1 - i);
      if (tag > first) first = tag;
    }
  }
  if (bad) atomicAdd(report + 0, bad);
  if (first) atomicMax(report + 1, first);
  if (longest) atomicMax(report + 3, longest);
}

__global__ void __launch_bounds__(512)
validate_values_kernel(const uint4* __restrict__ src, int64_t n_vec, unsigned long long* __restrict__ report) {
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  unsigned bad = 0;
  auto count = [&](uint32_t w) {   // two bf16 per word: exponent all ones = NaN or inf
    bad += ((w & 0x7f800000u) == 0x7f800000u) + ((w & 0x00007f80u) == 0x00007f80u);
  };
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
    const uint4 v = ldg_stream16(src + i);
    count(v.x); count(v.y); count(v.z); count(v.w);
  }
  bad = __reduce_add_sync(0xffffffffu, bad);
  if ((threadIdx.x & 31) == 0 && bad) atomicAdd(report + 2, (unsigned long long)bad);
}

}  // namespace

int launch_store_validate(const void* d_tokens, const int64_t* d_offsets, int64_t n_docs, int64_t total_tokens,
                          int check_values, unsigned long long* d_report, cudaStream_t stream) {
  HRC_CHECK_CUDA(cudaMemsetAsync(d_report, 0, 4 * sizeof(unsigned long long), stream));
  const int64_t blocks = (n_docs + 1 + 255) / 256;
  validate_offsets_kernel<<<unsigned(blocks < 148 * 8 ? blocks : 148 * 8), 256, 0, stream>>>(d_offsets, n_docs, total_tokens,
                                                                                           d_report);
  count_launch();
  if (check_values && total_tokens > 0) {
    HRC_REQUIRE((reinterpret_cast<uintptr_t>(d_tokens) & 15) == 0, "store_validate: token buffer must be 16-byte aligned");
    validate_values_kernel<<<148 * 8, 512, 0, stream>>>(static_cast<const uint4*>(d_tokens), total_tokens * (HRC_DIM * 2 / 16),
                                                        d_report);
    count_launch();
  }
  HRC_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace hrc
