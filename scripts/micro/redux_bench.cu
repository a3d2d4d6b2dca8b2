// Cross-lane column max of a 32 x 32 block held one row per lane (the doc-major MaxSim epilogue):
//   (a) 5-level max reduce-scatter butterfly (31 SHFL + 31 FMNMX + selects): lane j ends with column j's max
//   (b) redux.sync.max.f32 per column (CREDUX.MAX.F32 -> uniform register) + FMNMX into a replicated running max
// cycles per 32 x 32 block for 1, 2, 4 warps per SM sub-partition.   nvcc -arch... -o redux_bench redux_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int N, int OFF>
__device__ __forceinline__ void max_halve(float (&a)[32], bool upper) {
#pragma unroll
  for (int i = 0; i < N / 2; ++i) {
    const float keep = upper ? a[i + N / 2] : a[i];
    const float send = upper ? a[i] : a[i + N / 2];
    a[i] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, OFF));
  }
}
__device__ __forceinline__ float redux_max(float v) {
  float r;
  asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
  return r;
}

__global__ void bench(const float* in, float* out, long long* cycles, int iters, int mode) {
  const int lane = threadIdx.x & 31;
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = in[(threadIdx.x * 32 + j) & 1023];
  float m = -1e30f;
  float mr[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) mr[j] = -1e30f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (mode == 0) {
      float w[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) w[j] = v[j] + float(it);
      max_halve<32, 16>(w, (lane & 16) != 0);
      max_halve<16, 8>(w, (lane & 8) != 0);
      max_halve<8, 4>(w, (lane & 4) != 0);
      max_halve<4, 2>(w, (lane & 2) != 0);
      max_halve<2, 1>(w, (lane & 1) != 0);
      m = fmaxf(m, w[0]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) mr[j] = fmaxf(mr[j], redux_max(v[j] + float(it)));
    }
  }
  const long long t1 = clock64();
  float s = m;
#pragma unroll
  for (int j = 0; j < 32; ++j) s += mr[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
  float *in, *out;
  long long* cyc;
  cudaMalloc(&in, 4096);
  cudaMemset(in, 0, 4096);
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&cyc, 148 * 8);
  const int iters = 2000;
  for (int mode = 0; mode < 2; ++mode)
    for (int warps : {4, 8, 16}) {
      bench<<<148, warps * 32>>>(in, out, cyc, iters, mode);
      bench<<<148, warps * 32>>>(in, out, cyc, iters, mode);
      cudaDeviceSynchronize();
      long long h[148];
      cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      printf("%s  %2d warps/SM: %.1f cycles per 32x32 block per warp (%.1f per SM sub-partition slot)\n",
             mode == 0 ? "shuffle butterfly" : "redux.sync.max.f32", warps, double(h[0]) / iters, double(h[0]) / iters / (warps / 4.0));
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
