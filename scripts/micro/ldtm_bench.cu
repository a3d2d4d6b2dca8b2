// Microbenchmark: tcgen05.ld (TMEM -> registers) throughput on sm_100a, alone and mixed with a max tree,
// for 1..8 warps on the same / different TMEM lane groups.  Informs the MaxSim epilogue design.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o ldtm_bench ldtm_bench.cu && ./ldtm_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../hybrid-rag-colbertv2_b200/csrc/hrc_common.cuh"
namespace hrc { void set_error(const char*, ...) {} void count_launch(int) {} }
using namespace hrc;

__device__ __forceinline__ float max32f(const uint32_t (&v)[32]) {
  float t[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) t[i] = fmaxf(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
#pragma unroll
  for (int w = 8; w > 0; w >>= 1)
#pragma unroll
    for (int i = 0; i < w; ++i) t[i] = fmaxf(t[i], t[i + w]);
  return t[0];
}

// mode 0: ld, wait            (latency + throughput of one load at a time)
// mode 1: ld, ld, wait        (two loads in flight)
// mode 2: ld, wait, max32     (the serial epilogue step)
// mode 3: pipelined: ld(next) ; max32(cur) ; wait
// warp_mask: which of the 8 warps take part; warp w reads lane group w % 4
__global__ void __launch_bounds__(256, 1) ldtm_kernel(int mode, int warp_mask, int iters, long long* cycles, float* sink, int coloff) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t base = slot + (uint32_t((warp & 3) * 32) << 16) + coloff;
  float acc = 0.f;
  long long t0 = 0, t1 = 0;
  if ((warp_mask >> warp) & 1) {
    uint32_t a[32], b[32];
    __syncwarp();
    t0 = clock64();
    if (mode == 0) {
      for (int i = 0; i < iters; ++i) { tmem_ld_32x32(base + (i & 7) * 32, a); tmem_ld_wait(); acc += __uint_as_float(a[i & 31]); }
    } else if (mode == 1) {
      for (int i = 0; i < iters; i += 2) {
        tmem_ld_32x32(base + (i & 7) * 32, a); tmem_ld_32x32(base + ((i + 1) & 7) * 32, b); tmem_ld_wait();
        acc += __uint_as_float(a[i & 31]) + __uint_as_float(b[i & 31]);
      }
    } else if (mode == 2) {
      for (int i = 0; i < iters; ++i) { tmem_ld_32x32(base + (i & 7) * 32, a); tmem_ld_wait(); acc = fmaxf(acc, max32f(a)); }
    } else {
      tmem_ld_32x32(base, a); tmem_ld_wait();
      for (int i = 0; i < iters; i += 2) {
        tmem_ld_32x32(base + ((i + 1) & 7) * 32, b); acc = fmaxf(acc, max32f(a)); tmem_ld_wait();
        tmem_ld_32x32(base + ((i + 2) & 7) * 32, a); acc = fmaxf(acc, max32f(b)); tmem_ld_wait();
      }
    }
    t1 = clock64();
  }
  if ((threadIdx.x & 31) == 0) cycles[blockIdx.x * 8 + warp] = t1 - t0;
  sink[blockIdx.x * 256 + threadIdx.x] = acc;
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(slot, 512);
}

int main() {
  long long* d_cyc; float* d_sink;
  cudaMalloc(&d_cyc, 148 * 8 * sizeof(long long)); cudaMalloc(&d_sink, 148 * 256 * sizeof(float));
  const int iters = 4096;
  const int masks[] = {0x01, 0x11, 0x03, 0x0f, 0xff};
  const char* names[] = {"1 warp", "2 warps, same lane group", "2 warps, two lane groups", "4 warps, four lane groups", "8 warps"};
  for (int coloff = 0; coloff <= 17; coloff += (coloff == 0 ? 5 : 12))
  for (int mode = 2; mode < 4; ++mode)
    for (int m = 0; m < 5; ++m) {
      long long h[8];
      ldtm_kernel<<<1, 256>>>(mode, masks[m], iters, d_cyc, d_sink, coloff);
      if (cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
      cudaMemcpy(h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost);
      long long mx = 0; for (int w = 0; w < 8; ++w) mx = h[w] > mx ? h[w] : mx;
      int nw = __builtin_popcount(masks[m]);
      printf("coloff %2d mode %d  %-28s cycles/iter/warp %7.1f   SM bytes/cycle %7.1f\n", coloff, mode, names[m], double(mx) / iters,
             double(nw) * iters * 4096.0 / double(mx));
    }
  return 0;
}
