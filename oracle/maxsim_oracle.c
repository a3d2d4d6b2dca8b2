/*
 * maxsim_oracle.c — plain-C restatement of the MaxSim hot path.  TEST INFRASTRUCTURE ONLY: compiled by
 * __graft_entry__.build() into oracle/liboracle.so and loaded only by tests/ (as a second, independent
 * checker next to oracle/maxsim_oracle.py).  The product never links or loads it.
 *
 * Restated from /root/reference/local_rag_complete.py:
 *   oracle_maxsim_scores   docstring :807-812 ("for each query token, max similarity over the document's
 *                          tokens"), summed over query tokens (BASELINE.json north_star); fp32, index order.
 *                          PARITY ONLY PARTLY PINNED BY THE REFERENCE (its body :819-829 is a mean-pool cosine; it
 *                          ships no tests): bit-equal to the unmodified reference where that cosine IS MaxSim
 *                          (tests/golden/maxsim_pin.npz), within 2e-6 of float64 known answers
 *                          (maxsim_kat_f64.npz) and of scores vLLM 0.22.0's own MaxSim functions produced
 *                          (maxsim_vllm_pin.npz, tests/golden/make_vllm_pin.py), cross-checked against the Python oracle.
 *   oracle_literal_scores  the BODY of _maxsim_score as coded, :821-829: cosine of the mean-pooled token vectors
 *                          (fp32 accumulation in index order; torch reduces in a different order, so this agrees
 *                          with the reference's vectors to ~1e-6, not bit for bit).  PINNED on
 *                          tests/golden/literal_maxsim.npz and literal_bf16.npz.
 *   oracle_rrf             _reciprocal_rank_fusion :960-978: dict in insertion order, fp64 `0 + 1/(k+rank)`
 *                          accumulated list a then list b, stable descending sort.  PINNED on tests/golden/rrf.json.
 *   oracle_topk_keys       torch.topk :767 / argsort :789 expressed on the 64-bit (score, id) keys of
 *                          include/hrc.h (score descending, then id ascending).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define DIM 128

void oracle_maxsim_scores(const float* q, int n_queries, int lq, const float* tokens, const int64_t* offsets,
                          int64_t n_docs, float* out) {
  for (int b = 0; b < n_queries; ++b) {
    for (int64_t d = 0; d < n_docs; ++d) {
      float total = 0.0f;
      int empty = offsets[d + 1] == offsets[d];
      for (int i = 0; i < lq; ++i) {
        const float* qi = q + ((int64_t)b * lq + i) * DIM;
        float best = -INFINITY;
        for (int64_t t = offsets[d]; t < offsets[d + 1]; ++t) {
          const float* tk = tokens + t * DIM;
          float dot = 0.0f;
          for (int c = 0; c < DIM; ++c) dot += qi[c] * tk[c];
          if (dot > best) best = dot;
        }
        total += best;
      }
      out[(int64_t)b * n_docs + d] = empty ? -INFINITY : total;
    }
  }
}

/* the reference's function as coded (:821-829); dense documents [n_docs][ld][128], one query [lq][128] */
void oracle_literal_scores(const float* q, int lq, const float* docs, int64_t n_docs, int ld, float* out) {
  float qv[DIM];
  float qn = 0.0f;
  for (int c = 0; c < DIM; ++c) {
    float a = 0.0f;
    for (int i = 0; i < lq; ++i) a += q[(int64_t)i * DIM + c];
    qv[c] = a / (float)lq;                                      /* :821 */
    qn += qv[c] * qv[c];
  }
  qn = sqrtf(qn);
  if (qn < 1e-8f) qn = 1e-8f;                                   /* cosine_similarity eps */
  for (int64_t d = 0; d < n_docs; ++d) {
    float dot = 0.0f, dn = 0.0f;
    for (int c = 0; c < DIM; ++c) {
      float a = 0.0f;
      for (int t = 0; t < ld; ++t) a += docs[((int64_t)d * ld + t) * DIM + c];
      a /= (float)ld;                                           /* :822 */
      dot += a * qv[c];
      dn += a * a;
    }
    dn = sqrtf(dn);
    if (dn < 1e-8f) dn = 1e-8f;
    out[d] = dot / (dn * qn);                                   /* :825-829 */
  }
}

/* returns the number of distinct ids; ids_out / scores_out need room for na + nb entries */
int oracle_rrf(const int32_t* a, int na, const int32_t* b, int nb, int k, int32_t* ids_out, double* scores_out) {
  int n = 0;
  for (int pass = 0; pass < 2; ++pass) {
    const int32_t* list = pass ? b : a;
    int len = pass ? nb : na;
    for (int r = 0; r < len; ++r) {
      int32_t id = list[r];
      if (id < 0) continue;                                   /* absent slot (padding), keeps its rank */
      int j = 0;
      while (j < n && ids_out[j] != id) ++j;                  /* dict lookup, insertion order = index */
      if (j == n) { ids_out[n] = id; scores_out[n] = 0.0; ++n; }
      scores_out[j] = scores_out[j] + (1.0 / (double)(k + r + 1));      /* :971 / :975 */
    }
  }
  /* stable insertion sort, descending by score (sorted(..., reverse=True) keeps ties in original order) */
  for (int i = 1; i < n; ++i) {
    int32_t id = ids_out[i];
    double s = scores_out[i];
    int j = i - 1;
    while (j >= 0 && scores_out[j] < s) { ids_out[j + 1] = ids_out[j]; scores_out[j + 1] = scores_out[j]; --j; }
    ids_out[j + 1] = id;
    scores_out[j + 1] = s;
  }
  return n;
}

static uint32_t orderable(float s) {
  uint32_t u;
  if (s != s) s = -INFINITY;
  memcpy(&u, &s, 4);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

static int cmp_desc(const void* x, const void* y) {
  uint64_t a = *(const uint64_t*)x, b = *(const uint64_t*)y;
  return a < b ? 1 : (a > b ? -1 : 0);
}

/* out[0..k): the k largest keys in descending order, zero padded */
void oracle_topk_keys(const float* scores, const int32_t* ids, int64_t n, int k, int32_t id_base, uint64_t* out) {
  uint64_t* keys = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(n > 0 ? n : 1));
  for (int64_t i = 0; i < n; ++i) {
    int32_t id = ids ? ids[i] : (int32_t)(id_base + i);
    keys[i] = ((uint64_t)orderable(scores[i]) << 32) | (uint64_t)(uint32_t)(~(uint32_t)id);
  }
  qsort(keys, (size_t)n, sizeof(uint64_t), cmp_desc);
  for (int i = 0; i < k; ++i) out[i] = i < n ? keys[i] : 0;
  free(keys);
}
