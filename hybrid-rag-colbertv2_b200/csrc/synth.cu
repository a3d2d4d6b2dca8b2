// synth.cu — deterministic synthetic token embeddings (bench / test utility, not on the query path).
//
// The encoder weights of jinaai/jina-colbert-v2 are not available offline, so corpora are synthetic:
// row t = L2-normalised N(0,1)^128, rounded to bf16, drawn from a counter-based hash of
// (seed, GLOBAL token index, dim).  Because the stream is indexed by the global token number, any
// document sharding (1, 2, 4, 8 GPUs) materialises exactly the same corpus.
#include "hrc_common.cuh"

namespace hrc {

namespace {

__device__ __forceinline__ uint64_t mix64(uint64_t z) {  // splitmix64 finaliser
  z += 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& n0, float& n1) {
  const float u1 = (float(a >> 8) + 0.5f) * (1.0f / 16777216.0f);  // (0,1)
  const float u2 = (float(b >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float r = sqrtf(-2.0f * logf(u1));
  float s, c;
  sincospif(2.0f * u2, &s, &c);
  n0 = r * c;
  n1 = r * s;
}

// one warp per token; lane l produces dims 4l .. 4l+3
__global__ void __launch_bounds__(256)
synth_tokens_kernel(__nv_bfloat16* __restrict__ out, int64_t token_begin, int64_t n_tokens, uint64_t seed) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (blockIdx.x * int64_t(blockDim.x) + threadIdx.x) >> 5;
  const int64_t warps_total = (gridDim.x * int64_t(blockDim.x)) >> 5;
  for (int64_t t = warp_global; t < n_tokens; t += warps_total) {
    const uint64_t g = uint64_t(token_begin + t);
    const uint64_t h0 = mix64(mix64(seed ^ (g * 0xd1342543de82ef95ull)) + uint64_t(lane) * 2 + 1);
    const uint64_t h1 = mix64(h0 ^ 0xa0761d6478bd642full);
    float v0, v1, v2, v3;
    box_muller(uint32_t(h0), uint32_t(h0 >> 32), v0, v1);
    box_muller(uint32_t(h1), uint32_t(h1 >> 32), v2, v3);
    const float ss = warp_sum(v0 * v0 + v1 * v1 + v2 * v2 + v3 * v3);
    const float inv = rsqrtf(fmaxf(ss, 1e-20f));
    __nv_bfloat162 lo = __floats2bfloat162_rn(v0 * inv, v1 * inv);
    __nv_bfloat162 hi = __floats2bfloat162_rn(v2 * inv, v3 * inv);
    uint2 packed;
    packed.x = *reinterpret_cast<uint32_t*>(&lo);
    packed.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(out + t * HRC_DIM + 4 * lane) = packed;
  }
}

// Read-bandwidth probe (bench utility): every thread streams 16-byte loads, four in flight, and folds them into one
// word so that nothing is optimised away.  Calibrates "what can a pure read reach on this GPU" next to the
// read+write copy figure in MEASURED_PEAKS.json (BASELINE.md §2 asks for both).
__global__ void __launch_bounds__(512)
read_probe_kernel(const uint4* __restrict__ src, int64_t n_vec, uint32_t* __restrict__ out) {
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  uint32_t acc = 0;
  for (; i + 3 * stride < n_vec; i += 4 * stride) {
    const uint4 a = ldg_stream16(src + i), b = ldg_stream16(src + i + stride), c = ldg_stream16(src + i + 2 * stride),
                d = ldg_stream16(src + i + 3 * stride);
    acc ^= a.x ^ a.y ^ a.z ^ a.w ^ b.x ^ b.y ^ b.z ^ b.w ^ c.x ^ c.y ^ c.z ^ c.w ^ d.x ^ d.y ^ d.z ^ d.w;
  }
  for (; i < n_vec; i += stride) {
    const uint4 a = ldg_stream16(src + i);
    acc ^= a.x ^ a.y ^ a.z ^ a.w;
  }
  acc = __reduce_xor_sync(0xffffffffu, acc);
  if ((threadIdx.x & 31) == 0 && acc == 0x9e3779b9u) atomicXor(out, acc);   // practically never taken; keeps the loads live
}

}  // namespace

int launch_read_probe(const void* d_buf, size_t bytes, uint32_t* d_out, cudaStream_t stream) {
  HRC_REQUIRE((reinterpret_cast<uintptr_t>(d_buf) & 15) == 0 && d_out != nullptr, "read_probe: buffer must be 16-byte aligned");
  if (bytes < 16) return 0;
  read_probe_kernel<<<148 * 8, 512, 0, stream>>>(static_cast<const uint4*>(d_buf), int64_t(bytes / 16), d_out);
  count_launch();
  HRC_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_synth(void* d_out, int64_t token_begin, int64_t n_tokens, uint64_t seed, cudaStream_t stream) {
  if (n_tokens <= 0) return 0;
  const int64_t blocks_needed = (n_tokens + 7) / 8;
  const unsigned grid = unsigned(blocks_needed < 148 * 16 ? blocks_needed : 148 * 16);
  synth_tokens_kernel<<<grid, 256, 0, stream>>>(static_cast<__nv_bfloat16*>(d_out), token_begin, n_tokens, seed);
  count_launch();
  HRC_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace hrc
