/*
 * A plain C client of libhrc.so: no Python, no torch, no C++ — only include/hrc.h and the CUDA runtime for device
 * memory.  It shows that the drop-in boundary really is a C ABI, and cross-checks the tensor-core path against the
 * CUDA-core path and against a host loop through that ABI.
 *   gcc -O2 -std=c99 -I include -I /usr/local/cuda/include tests/c_abi/client.c -o client \
 *       -L hybrid-rag-colbertv2_b200 -l:libhrc.so -L /usr/local/cuda/lib64 -lcudart -lm -Wl,-rpath,$PWD/hybrid-rag-colbertv2_b200
 */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "hrc.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)
#define HK(x) do { int r_ = (x); if (r_ != 0) { printf("hrc error %d: %s (%s:%d)\n", r_, hrc_last_error(), __FILE__, __LINE__); return 1; } } while (0)

static float bf16_to_f32(uint16_t h) { uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return f; }

int main(void) {
  enum { N_DOCS = 3000, NQ = 3, LQ = 32, K = 50 };
  int64_t* off = (int64_t*)malloc(sizeof(int64_t) * (N_DOCS + 1));
  uint32_t lcg = 12345u;
  off[0] = 0;
  for (int d = 0; d < N_DOCS; ++d) { lcg = lcg * 1664525u + 1013904223u; off[d + 1] = off[d] + 1 + (lcg >> 16) % 90; }
  const int64_t T = off[N_DOCS];
  void *d_tok, *d_q; int64_t* d_off; float *d_tc, *d_simt; uint64_t* d_keys; int32_t* d_ids; float* d_top; void* d_ws;
  CK(cudaMalloc(&d_tok, (size_t)T * HRC_DIM * 2));
  CK(cudaMalloc(&d_q, (size_t)NQ * LQ * HRC_DIM * 2));
  CK(cudaMalloc((void**)&d_off, sizeof(int64_t) * (N_DOCS + 1)));
  CK(cudaMalloc((void**)&d_tc, sizeof(float) * NQ * N_DOCS));
  CK(cudaMalloc((void**)&d_simt, sizeof(float) * NQ * N_DOCS));
  CK(cudaMalloc((void**)&d_keys, sizeof(uint64_t) * NQ * K));
  CK(cudaMalloc((void**)&d_ids, sizeof(int32_t) * NQ * K));
  CK(cudaMalloc((void**)&d_top, sizeof(float) * NQ * K));
  size_t ws_bytes = hrc_topk_workspace_bytes(N_DOCS, NQ, K);
  CK(cudaMalloc(&d_ws, ws_bytes ? ws_bytes : 256));
  CK(cudaMemcpy(d_off, off, sizeof(int64_t) * (N_DOCS + 1), cudaMemcpyHostToDevice));
  if (hrc_version() != 200) { printf("unexpected version %d\n", hrc_version()); return 1; }
  HK(hrc_synth_tokens(d_tok, 0, T, 7, NULL));                       /* unit-norm bf16 rows */
  HK(hrc_synth_tokens(d_q, 1000000000ll, (int64_t)NQ * LQ, 7, NULL)); /* queries: other rows of the same generator */
  HK(hrc_maxsim_scores(d_tok, d_off, N_DOCS, T, d_q, NQ, LQ, d_tc, HRC_PATH_TC, NULL, 0, NULL));
  HK(hrc_maxsim_scores(d_tok, d_off, N_DOCS, T, d_q, NQ, LQ, d_simt, HRC_PATH_SIMT, NULL, 0, NULL));
  HK(hrc_topk(d_tc, NULL, N_DOCS, NQ, K, 0, d_keys, d_ws, ws_bytes, NULL));
  HK(hrc_keys_unpack(d_keys, (int64_t)NQ * K, d_ids, d_top, NULL));
  CK(cudaDeviceSynchronize());

  float* tc = (float*)malloc(sizeof(float) * NQ * N_DOCS);
  float* simt = (float*)malloc(sizeof(float) * NQ * N_DOCS);
  uint16_t* tok = (uint16_t*)malloc((size_t)T * HRC_DIM * 2);
  uint16_t* q = (uint16_t*)malloc((size_t)NQ * LQ * HRC_DIM * 2);
  int32_t* ids = (int32_t*)malloc(sizeof(int32_t) * NQ * K);
  float* top = (float*)malloc(sizeof(float) * NQ * K);
  CK(cudaMemcpy(tc, d_tc, sizeof(float) * NQ * N_DOCS, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(simt, d_simt, sizeof(float) * NQ * N_DOCS, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(tok, d_tok, (size_t)T * HRC_DIM * 2, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(q, d_q, (size_t)NQ * LQ * HRC_DIM * 2, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(ids, d_ids, sizeof(int32_t) * NQ * K, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(top, d_top, sizeof(float) * NQ * K, cudaMemcpyDeviceToHost));

  /* 1. the two kernels agree */
  double worst = 0.0;
  for (int i = 0; i < NQ * N_DOCS; ++i) { double e = fabs(tc[i] - simt[i]) / fmax(fabs(simt[i]), 1e-6); if (e > worst) worst = e; }
  if (worst > 1e-3) { printf("tc vs simt: relative error %g\n", worst); return 1; }
  /* 2. a host loop (local_rag_complete.py:807-812, summed over query tokens) agrees on every 37th document */
  for (int b = 0; b < NQ; ++b)
    for (int d = 0; d < N_DOCS; d += 37) {
      float total = 0.f;
      for (int i = 0; i < LQ; ++i) {
        float best = -INFINITY;
        for (int64_t t = off[d]; t < off[d + 1]; ++t) {
          float dot = 0.f;
          for (int c = 0; c < HRC_DIM; ++c)
            dot += bf16_to_f32(q[((size_t)b * LQ + i) * HRC_DIM + c]) * bf16_to_f32(tok[(size_t)t * HRC_DIM + c]);
          if (dot > best) best = dot;
        }
        total += best;
      }
      if (fabs(total - tc[b * N_DOCS + d]) > 1e-3 * fabs(total)) { printf("host vs tc: query %d doc %d: %g vs %g\n", b, d, total, tc[b * N_DOCS + d]); return 1; }
    }
  /* 3. top-k: sorted, scores equal the score row, first entry is the row's maximum */
  for (int b = 0; b < NQ; ++b) {
    float mx = -INFINITY;
    for (int d = 0; d < N_DOCS; ++d) if (tc[b * N_DOCS + d] > mx) mx = tc[b * N_DOCS + d];
    if (top[b * K] != mx) { printf("top-1 of query %d is not the maximum\n", b); return 1; }
    for (int j = 0; j < K; ++j) {
      if (j && top[b * K + j] > top[b * K + j - 1]) { printf("top-k not sorted\n"); return 1; }
      if (top[b * K + j] != tc[b * N_DOCS + ids[b * K + j]]) { printf("key score != score row\n"); return 1; }
    }
  }
  /* 4. the one-call entry points with caller-owned workspaces: hrc_search (one query: 2 launches, the top-k fused
   *    into the doc-major MaxSim epilogue) must return the keys of step 3, and hrc_rerank (ONE launch) must rank the
   *    first 40 of them in the same order */
  {
    enum { NC = 40, RK = 10 };
    size_t sws = hrc_search_workspace_bytes(N_DOCS, T, 1, LQ, K, HRC_PATH_AUTO);
    size_t rws = hrc_rerank_workspace_bytes(NC, 1, LQ, RK);
    void *d_sws, *d_rws; uint64_t* d_k1; int32_t *d_i1, *d_pos, *d_rid; float *d_s1, *d_rs;
    CK(cudaMalloc(&d_sws, sws)); CK(cudaMalloc(&d_rws, rws));
    CK(cudaMalloc((void**)&d_k1, sizeof(uint64_t) * K)); CK(cudaMalloc((void**)&d_i1, sizeof(int32_t) * K));
    CK(cudaMalloc((void**)&d_s1, sizeof(float) * K)); CK(cudaMalloc((void**)&d_pos, sizeof(int32_t) * RK));
    CK(cudaMalloc((void**)&d_rid, sizeof(int32_t) * RK)); CK(cudaMalloc((void**)&d_rs, sizeof(float) * RK));
    const unsigned long long before = (unsigned long long)hrc_launch_count();
    HK(hrc_search(d_tok, d_off, N_DOCS, T, d_q, 1, LQ, K, 0, d_sws, sws, d_k1, d_i1, d_s1, HRC_PATH_AUTO, NULL));
    if ((unsigned long long)hrc_launch_count() - before != 2) { printf("hrc_search for one query should be 2 launches\n"); return 1; }
    HK(hrc_rerank(d_tok, d_off, N_DOCS, T, d_i1, NC, d_q, 1, LQ, RK, d_rws, rws, d_pos, d_rid, d_rs, NULL, HRC_PATH_AUTO, NULL));
    if ((unsigned long long)hrc_launch_count() - before != 3) { printf("hrc_rerank should be 1 launch\n"); return 1; }
    CK(cudaDeviceSynchronize());
    int32_t i1[K], pos[RK], rid[RK]; float s1[K], rs[RK];
    CK(cudaMemcpy(i1, d_i1, sizeof(i1), cudaMemcpyDeviceToHost)); CK(cudaMemcpy(s1, d_s1, sizeof(s1), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(pos, d_pos, sizeof(pos), cudaMemcpyDeviceToHost)); CK(cudaMemcpy(rid, d_rid, sizeof(rid), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(rs, d_rs, sizeof(rs), cudaMemcpyDeviceToHost));
    for (int j = 0; j < K; ++j)
      if (i1[j] != ids[j] || fabsf(s1[j] - top[j]) > 4e-6f * fabsf(top[j])) { printf("hrc_search differs from scores + top-k at rank %d\n", j); return 1; }
    for (int j = 0; j < RK; ++j)      /* candidates were handed over best first: the rerank keeps that order */
      if (pos[j] != j || rid[j] != ids[j] || rs[j] != s1[j]) { printf("hrc_rerank: rank %d is position %d\n", j, pos[j]); return 1; }
    if (hrc_search(d_tok, d_off, N_DOCS, T, d_q, 1, LQ, K, 0, d_sws, sws - 1, d_k1, d_i1, d_s1, HRC_PATH_AUTO, NULL) == 0) {
      printf("a workspace one byte short was accepted\n"); return 1;
    }
  }
  /* 5. errors are reported, not swallowed */
  if (hrc_topk(d_tc, NULL, N_DOCS, NQ, HRC_MAX_TOPK + 1, 0, d_keys, d_ws, ws_bytes, NULL) == 0 || hrc_last_error()[0] == 0) {
    printf("k > HRC_MAX_TOPK was accepted\n"); return 1;
  }
  printf("c_abi_client ok: %d docs, %lld tokens, %llu kernels launched, tc vs simt rel err %.2e\n", N_DOCS, (long long)T,
         (unsigned long long)hrc_launch_count(), worst);
  return 0;
}
