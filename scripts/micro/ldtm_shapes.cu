// Microbenchmark: tcgen05.ld throughput by SHAPE (4 KB per instruction each): 32x32b.x32, 16x256b.x8, 16x128b.x16,
// 16x64b.x32, one warp and four warps (one per lane group).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o ldtm_shapes ldtm_shapes.cu && ./ldtm_shapes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../hybrid-rag-colbertv2_b200/csrc/hrc_common.cuh"
namespace hrc { void set_error(const char*, ...) {} void count_launch(int) {} }
using namespace hrc;

#define REGS32(v) "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), \
  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), \
  "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), \
  "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
#define LIST32 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}"

template <int SHAPE>
__device__ __forceinline__ void ld(uint32_t taddr, uint32_t (&v)[32]) {
  if constexpr (SHAPE == 0) asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 " LIST32 ", [%32];" : REGS32(v) : "r"(taddr) : "memory");
  if constexpr (SHAPE == 1) asm volatile("tcgen05.ld.sync.aligned.16x256b.x8.b32 " LIST32 ", [%32];" : REGS32(v) : "r"(taddr) : "memory");
  if constexpr (SHAPE == 2) asm volatile("tcgen05.ld.sync.aligned.16x128b.x16.b32 " LIST32 ", [%32];" : REGS32(v) : "r"(taddr) : "memory");
  if constexpr (SHAPE == 3) asm volatile("tcgen05.ld.sync.aligned.16x64b.x32.b32 " LIST32 ", [%32];" : REGS32(v) : "r"(taddr) : "memory");
}

template <int SHAPE>
__global__ void __launch_bounds__(128, 1) k(int warp_mask, int iters, long long* cycles, float* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t base = slot + (uint32_t((warp & 3) * 32) << 16);
  float acc = 0.f;
  long long t0 = 0, t1 = 0;
  if ((warp_mask >> warp) & 1) {
    uint32_t a[32];
    __syncwarp();
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      ld<SHAPE>(base + (i & 3) * 64, a);
      tmem_ld_wait();
      float m = __uint_as_float(a[0]);
#pragma unroll
      for (int j = 1; j < 32; ++j) m = fmaxf(m, __uint_as_float(a[j]));
      acc = fmaxf(acc, m);
    }
    t1 = clock64();
  }
  if ((threadIdx.x & 31) == 0) cycles[warp] = t1 - t0;
  sink[threadIdx.x] = acc;
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(slot, 512);
}

template <int SHAPE>
void run(const char* name, long long* d_cyc, float* d_sink) {
  const int iters = 4096;
  for (int mask : {0x1, 0xf}) {
    long long h[4];
    k<SHAPE><<<1, 128>>>(mask, iters, d_cyc, d_sink);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("%s failed: %s\n", name, cudaGetErrorString(cudaGetLastError())); return; }
    cudaMemcpy(h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0; for (int w = 0; w < 4; ++w) mx = h[w] > mx ? h[w] : mx;
    printf("%-14s %d warp(s): %6.1f cycles per (4 KB load + wait + 32-way max)\n", name, __builtin_popcount(mask), double(mx) / iters);
  }
}

int main() {
  long long* d_cyc; float* d_sink;
  cudaMalloc(&d_cyc, 8 * sizeof(long long)); cudaMalloc(&d_sink, 128 * sizeof(float));
  run<0>("32x32b.x32", d_cyc, d_sink);
  run<1>("16x256b.x8", d_cyc, d_sink);
  run<2>("16x128b.x16", d_cyc, d_sink);
  run<3>("16x64b.x32", d_cyc, d_sink);
  return 0;
}
