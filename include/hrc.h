/*
 * hrc.h — C ABI of libhrc.so: B200 (sm_100a) late-interaction MaxSim scoring, top-k and RRF fusion.
 *
 * This is the drop-in boundary for the ONE hot path of techmum21p/hybrid-rag-ColBERTv2:
 * JinaColBERTRetriever's MaxSim scoring + top-k (first stage and post-RRF rerank) and the RRF step
 * between them.  The reference has no FFI of its own (it is a single Python script that calls torch
 * on CPU/MPS), so each entry point below cites the reference Python statement(s) it replaces
 * (paths relative to /root/reference).  The Python classes in hybrid-rag-colbertv2_b200/retriever.py
 * bind these symbols with ctypes.
 *
 * Conventions
 *   - plain C symbols, plain pointers and sizes; no torch / pybind types.
 *   - every pointer named d_* is DEVICE memory owned by the caller, valid on the current device.
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*; NULL = legacy stream).
 *   - return value: 0 = ok, non-zero = error; hrc_last_error() returns a thread-local message.
 *   - threads: any number of host threads may call concurrently as long as calls that share a workspace (or output
 *     buffers) are ordered on one stream; the library's own state (descriptor cache, trace, launch counter) is locked
 *     or atomic.  Calls that take device pointers neither allocate nor synchronise, so they can be captured in a
 *     CUDA graph; the *_host entry points, hrc_store_validate, the file IO and the trace collector synchronise.
 *   - token embeddings are bf16, HRC_DIM (=128) wide, rows L2-normalised by the caller.
 *   - corpus layout: packed, padding-free  d_tokens[total_tokens][128]  plus CSR  d_offsets[n_docs+1]
 *     (int64, d_offsets[0] == 0, d_offsets[n_docs] == total_tokens).
 *   - a "key" is a 64-bit totally ordered (score, doc_id) pair:
 *         key = (uint64)orderable(score) << 32 | (uint32)~doc_id
 *     so that larger key == higher score, and among equal scores the LOWER doc_id wins.
 *     key 0 is the "empty slot" sentinel.
 */
#ifndef HRC_H_
#define HRC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HRC_DIM 128            /* embedding width (jina-colbert-v2 projects to 128)            */
#define HRC_MAX_TOPK 2048      /* largest k the selection kernels accept                        */
#define HRC_TC_MAX_LQ 32       /* query tokens per query slot on the tcgen05 path               */
#define HRC_TC_MAX_SLOTS 8     /* a longer query is scored as up to 8 slots: lq <= 256 on that path */
#define HRC_FUSED_TOPK_MAX_QUERIES 1   /* hrc_search fuses the top-k into the MaxSim epilogue up to this many queries */

/* scoring path selector */
#define HRC_PATH_AUTO 0
#define HRC_PATH_SIMT 1        /* coalesced 16-byte loads + warp-shuffle reduction (CUDA cores) */
#define HRC_PATH_TC   2        /* TMA + tcgen05.mma + TMEM, fused segmented max/sum epilogue     */
#define HRC_PATH_TC_DM 3       /* as TC, but ONE query runs doc-major (documents on M, query on N = 32): exactly the useful tensor work */

/* Library / ABI version (major*10000 + minor*100 + patch). */
int hrc_version(void);

/* Thread-local message of the last failing call on this thread ("" if none). */
const char* hrc_last_error(void);

/* Number of kernels this library has launched since load (all threads). Used by bench.py's
 * gpu_launches claim. */
uint64_t hrc_launch_count(void);

/*
 * Every mbarrier wait inside the tensor-core kernels traps after this many milliseconds without progress, so that a
 * protocol bug faults instead of hanging the GPU (default 20000; 0 disables, e.g. under a debugger, compute-sanitizer
 * or GPU time-slicing).  Process-wide.  The library reads no environment variables.
 */
void hrc_set_watchdog_ms(uint64_t ms);

/*
 * Kernel trace (the reference's only observability is wall-clock stage timing, local_rag_complete.py:902-933; this
 * is the device-side counterpart).  After hrc_trace_enable(capacity) every launch of a scoring kernel (tensor-core or
 * CUDA-core MaxSim) is bracketed by CUDA events on its stream, up to `capacity` launches; hrc_trace_collect waits
 * for them, writes the launches' device times in milliseconds (launch order) and returns how many it wrote (-1 on
 * error), resetting the trace.  hrc_trace_enable(0) switches tracing off.  Process-wide; bench.py uses it to time the
 * dominant kernel INSIDE the timed steps.
 */
int hrc_trace_enable(int capacity);
int hrc_trace_collect(float* ms_out, int max_n);

/*
 * TMA descriptors (CUtensorMap) depend only on a buffer's address and extents, so the library keeps the ones it has
 * encoded in a small cache and never encodes on the launch path twice for the same store / query buffer.
 * hrc_store_register pre-encodes the maps of a token store (optional: the first scoring call does it otherwise);
 * hrc_store_release drops them (call it before freeing the store).  Replaces: the lifetime of
 * `self.corpus_embeddings`, local_rag_complete.py:725,735,752.
 */
int hrc_store_register(const void* d_tokens, int64_t total_tokens);
void hrc_store_release(const void* d_tokens);

/*
 * Integrity check of a packed store, meant to run ONCE when an index is loaded or built (not on the search path): the
 * scoring kernels trust the CSR offsets and the finiteness of the token rows.  Checks offsets[0] = 0, offsets
 * non-decreasing, offsets[n_docs] = total_tokens, and (check_values != 0) streams the tokens once looking for NaN /
 * infinite values.  d_workspace: >= 32 bytes of device memory, 8-byte aligned.  Synchronises the stream.
 * report_out (host, optional, 4 x int64): [0] offending offsets entries, [1] index of the first one (-1 = none),
 * [2] non-finite token values, [3] longest document in tokens.  Returns 0 = sound, 3 = offsets broken, 4 = non-finite
 * values (hrc_last_error says which); the reference has no counterpart — torch.load at local_rag_complete.py:751-753
 * trusts the file.
 */
int hrc_store_validate(const void* d_tokens, const int64_t* d_offsets, int64_t n_docs, int64_t total_tokens,
                       int check_values, void* d_workspace, size_t workspace_bytes, int64_t* report_out, void* stream);

/*
 * MaxSim scores of every query against every document of the packed store.
 *   d_scores[q * n_docs + d] = sum_{i < lq} max_{t in doc d} <Q[q][i], tokens[t]>      (fp32 accumulate)
 * Replaces: JinaColBERTRetriever._maxsim_score, local_rag_complete.py:802-831 (as its docstring
 * :807-812 and BASELINE.json's north_star define it), called from search :764.
 *   d_queries : bf16 [n_queries][lq][128]
 *   path      : HRC_PATH_*; AUTO picks the tensor cores when lq <= HRC_TC_MAX_LQ * HRC_TC_MAX_SLOTS — the doc-major
 *               kernel for ONE query of <= 32 tokens, the query-major kernels otherwise (a query longer than 32 tokens
 *               is scored as ceil(lq / 32) slots whose partial scores are summed in slot order) — and the CUDA cores
 *               beyond.  All tensor-core organisations return bit-identical scores.
 *   d_workspace : hrc_maxsim_workspace_bytes(n_docs, n_queries, lq) bytes of scratch: the per-slot partial scores of
 *               queries longer than 32 tokens (required), or — for ONE query of <= 32 tokens — 256 bytes holding the
 *               doc-major kernel's claim counter (OPTIONAL: NULL / 0 is accepted and selects equal token ranges per
 *               CTA instead of run-time work units; same scores, ~2 % slower on large corpora of short documents);
 *               0 bytes otherwise.  No call allocates device memory.
 * An empty document (length 0) scores -inf.
 */
size_t hrc_maxsim_workspace_bytes(int64_t n_items, int n_queries, int lq);
int hrc_maxsim_scores(const void* d_tokens, const int64_t* d_offsets, int64_t n_docs,
                      int64_t total_tokens, const void* d_queries, int n_queries, int lq,
                      float* d_scores, int path, void* d_workspace, size_t workspace_bytes, void* stream);

/*
 * MaxSim scores of query q against its own candidate list (the rerank shape):
 *   d_scores[q * n_cand + j] = maxsim(Q[q], doc d_cand_ids[q * n_cand + j])
 * Replaces: JinaColBERTRetriever.rerank's scoring step, local_rag_complete.py:782-786 (the
 * reference re-encodes the candidate texts; here the stored token embeddings are gathered by id).
 * Candidate ids < 0 or >= n_docs score -inf.  d_workspace: hrc_maxsim_workspace_bytes(n_cand, n_queries, lq).
 */
int hrc_maxsim_scores_ids(const void* d_tokens, const int64_t* d_offsets, int64_t n_docs,
                          int64_t total_tokens, const int32_t* d_cand_ids, int n_cand,
                          const void* d_queries, int n_queries, int lq, float* d_scores,
                          int path, void* d_workspace, size_t workspace_bytes, void* stream);

/*
 * The reference's `_maxsim_score` EXACTLY AS CODED, local_rag_complete.py:821-829 — the cosine of the
 * mean-pooled token vectors (its "simplified" stand-in for MaxSim):
 *   d_scores[q * n_docs + d] = cos( mean_i Q[q][i] , mean_{t in doc d} tokens[t] ),  eps = 1e-8 (torch default)
 * Same layouts as hrc_maxsim_scores; an empty document scores NaN (torch's mean over an empty axis).
 * Pinned by vectors the unmodified reference produced (tests/golden/literal_maxsim.npz).
 */
int hrc_meanpool_cosine_scores(const void* d_tokens, const int64_t* d_offsets, int64_t n_docs,
                               int64_t total_tokens, const void* d_queries, int n_queries, int lq,
                               float* d_scores, void* stream);

/*
 * Fused search: MaxSim of every query against the whole store, per-query top-k, optional unpacking —
 * the body of JinaColBERTRetriever.search (local_rag_complete.py:764-775) in one call.
 * A single query (the HBM-bound case) with lq <= 32 and k <= 128 on the tensor-core path is TWO launches: the MaxSim
 * kernel keeps a per-warp top-k of the scores its epilogue emits (the score row is never written) and hands 128 keys per
 * corpus segment to one merge-sort-unpack launch.  Otherwise: score matrix -> streaming top-k (k <= 128: every warp
 * streams its slice with a running threshold and a small sorted list) or radix select (larger k) -> merge + unpack.
 *   d_workspace  : hrc_search_workspace_bytes(n_docs, total_tokens, n_queries, lq, k, path) bytes of scratch, 256-B aligned
 *   d_keys_out   : uint64 [n_queries][k]; d_ids_out / d_scores_out optional int32 / fp32 [n_queries][k]
 * k <= HRC_MAX_TOPK and k <= n_docs.
 */
size_t hrc_search_workspace_bytes(int64_t n_docs, int64_t total_tokens, int n_queries, int lq, int k, int path);
int hrc_search(const void* d_tokens, const int64_t* d_offsets, int64_t n_docs, int64_t total_tokens,
               const void* d_queries, int n_queries, int lq, int k, int32_t id_base, void* d_workspace,
               size_t workspace_bytes, uint64_t* d_keys_out, int32_t* d_ids_out, float* d_scores_out, int path,
               void* stream);

/*
 * End-to-end search with HOST buffers — what JinaColBERTRetriever.search does between "the encoder returned the
 * query embedding" (local_rag_complete.py:758-761) and "the result list is built" (:769-775), in one call:
 * H2D copy of the fp32 query embeddings, fp32 -> bf16, MaxSim over the whole store, per-query top-k, unpack, and
 * D2H copies of ids and scores.  Everything is enqueued on `stream`; the caller synchronises the stream before
 * reading h_ids_out / h_scores_out.  Host buffers should be pinned (page-locked) for the copies to be asynchronous.
 *   h_queries    : fp32 [n_queries][lq][128] host
 *   d_workspace  : hrc_search_host_workspace_bytes(n_docs, total_tokens, n_queries, lq, k, path) bytes of device scratch, 256-B aligned
 *   h_ids_out    : int32 [n_queries][k] host;  h_scores_out : fp32 [n_queries][k] host
 */
size_t hrc_search_host_workspace_bytes(int64_t n_docs, int64_t total_tokens, int n_queries, int lq, int k, int path);
int hrc_search_host(const void* d_tokens, const int64_t* d_offsets, int64_t n_docs, int64_t total_tokens,
                    const float* h_queries, int n_queries, int lq, int k, int32_t id_base, void* d_workspace,
                    size_t workspace_bytes, int32_t* h_ids_out, float* h_scores_out, int path, void* stream);

/*
 * Fused rerank: MaxSim of query q against its candidate list, sorted top-k — the body of
 * JinaColBERTRetriever.rerank (local_rag_complete.py:786-798) on stored embeddings.
 *   d_cand_ids        : int32 [n_queries][n_cand]
 *   d_workspace       : hrc_rerank_workspace_bytes(n_cand, n_queries, lq, k) bytes of scratch, 256-B aligned
 *   d_pos_out         : int32 [n_queries][k] position in the candidate list ("result_index"), -1 = empty
 *   d_ids_out         : int32 [n_queries][k] the candidate's document id, optional
 *   d_scores_out      : fp32  [n_queries][k]
 *   d_cand_scores_out : optional fp32 [n_queries][n_cand], the score of every candidate
 */
size_t hrc_rerank_workspace_bytes(int n_cand, int n_queries, int lq, int k);
int hrc_rerank(const void* d_tokens, const int64_t* d_offsets, int64_t n_docs, int64_t total_tokens,
               const int32_t* d_cand_ids, int n_cand, const void* d_queries, int n_queries, int lq, int k,
               void* d_workspace, size_t workspace_bytes, int32_t* d_pos_out, int32_t* d_ids_out,
               float* d_scores_out, float* d_cand_scores_out, int path, void* stream);

/*
 * The device part of HybridRetriever.retrieve (local_rag_complete.py:894-935) in one call, for a batch of queries:
 * ColBERT first stage over the whole store (top colbert_k, :908-911) -> reciprocal-rank fusion with the given BM25
 * lists (:914-916, bit-compatible fp64, top n_candidates) -> rerank of the candidates' STORED token embeddings
 * (:926-929, sorted top final_k).  Nothing returns to the host in between.
 *   d_bm25_ids   : int32 [n_queries][n_bm25] ranked GLOBAL document ids as bm25s.retrieve returns them (:945-949), < 0 = absent
 *   id_base      : global id of this store's first document (document-sharded stores)
 *   d_workspace  : hrc_hybrid_retrieve_workspace_bytes(...) bytes of device scratch, 256-B aligned
 *   d_ids_out    : int32 [n_queries][final_k] GLOBAL ids, best first (-1 = fewer candidates than final_k)
 *   d_scores_out : fp32  [n_queries][final_k] MaxSim scores
 * Requires 1 <= colbert_k <= n_docs, 1 <= final_k <= n_candidates, n_bm25 + colbert_k <= 16384.
 */
size_t hrc_hybrid_retrieve_workspace_bytes(int64_t n_docs, int64_t total_tokens, int n_queries, int lq, int colbert_k,
                                           int n_candidates, int final_k, int path);
int hrc_hybrid_retrieve(const void* d_tokens, const int64_t* d_offsets, int64_t n_docs, int64_t total_tokens,
                        const void* d_queries, int n_queries, int lq, const int32_t* d_bm25_ids, int n_bm25,
                        int colbert_k, int rrf_k, int n_candidates, int final_k, int32_t id_base, void* d_workspace,
                        size_t workspace_bytes, int32_t* d_ids_out, float* d_scores_out, int path, void* stream);

/* Bytes of scratch hrc_topk needs for these sizes. */
size_t hrc_topk_workspace_bytes(int64_t n, int n_rows, int k);

/*
 * Per-row top-k of a score matrix, sorted (descending score, ascending id).
 *   d_scores   : fp32 [n_rows][n]
 *   d_ids      : optional int32 [n_rows][n] ids of the columns (NULL: id = id_base + column)
 *   d_keys_out : uint64 [n_rows][k]; slots beyond min(k, n) are 0
 * Replaces: torch.topk at local_rag_complete.py:767 and torch.argsort at :789 (+ the [:k] at :792).
 * NaN scores order as -inf.
 */
int hrc_topk(const float* d_scores, const int32_t* d_ids, int64_t n, int n_rows, int k,
             int32_t id_base, uint64_t* d_keys_out, void* d_workspace, size_t workspace_bytes,
             void* stream);

/*
 * Merge: per-row top-k of n_in keys (e.g. the all-gathered per-GPU top-k lists), sorted.
 *   d_keys_in  : uint64 [n_rows][n_in] (0 = empty);  d_keys_out : uint64 [n_rows][k]
 * New step (document-sharded search); the single-GPU reference has no counterpart.
 * n_in * 8 bytes must fit the kernel's shared memory (n_in <= 24576).
 */
int hrc_topk_merge(const uint64_t* d_keys_in, int n_in, int n_rows, int k, uint64_t* d_keys_out,
                   void* stream);

/* Unpack keys into ids (int32, -1 for empty) and scores (fp32, -inf for empty). */
int hrc_keys_unpack(const uint64_t* d_keys, int64_t n, int32_t* d_ids_out, float* d_scores_out,
                    void* stream);

/*
 * Reciprocal-rank fusion of two ranked id lists per row, bit-compatible with the reference's
 * Python-float (IEEE fp64) arithmetic, accumulation order and stable tie order.
 * Replaces: HybridRetriever._reciprocal_rank_fusion, local_rag_complete.py:960-978, and the
 * [:50] slice at :916.
 *   d_ids_a : int32 [n_rows][n_a]  first list (BM25 order), rank = position+1; id < 0 = absent
 *   d_ids_b : int32 [n_rows][n_b]  second list (ColBERT order)
 *   rrf_k   : the constant (60 in the reference)
 *   d_ids_out    : int32  [n_rows][top_n]  fused ids, best first (-1 padded)
 *   d_scores_out : double [n_rows][top_n]  fused scores (0 padded)
 *   d_counts_out : int32  [n_rows]         number of distinct ids (may exceed top_n), optional
 * n_a + n_b <= 16384.
 */
int hrc_rrf_fuse(const int32_t* d_ids_a, int n_a, const int32_t* d_ids_b, int n_b, int n_rows,
                 int rrf_k, int top_n, int32_t* d_ids_out, double* d_scores_out,
                 int32_t* d_counts_out, void* stream);

/*
 * Deterministic synthetic token embeddings (test/bench utility, not on the query path):
 * row t (global token index token_begin + t) = L2-normalised N(0,1)^128 drawn from a counter-based
 * hash of (seed, global token index, dim), rounded to bf16.  Any sharding reproduces the same corpus.
 */
int hrc_synth_tokens(void* d_tokens_out, int64_t token_begin, int64_t n_tokens, uint64_t seed,
                     void* stream);

/*
 * ---- multi-GPU: the exchange step of the document-sharded search ------------------------------------------------
 * The corpus shards by document over the GPUs of one box, ONE PROCESS PER GPU; each rank computes a
 * local top-k and the ranks exchange k keys each.  The single-process reference has no counterpart
 * (local_rag_complete.py contains no distributed code); these calls make the N>1 search one C call like the N=1 one.
 *   hrc_comm_unique_id  rank 0 creates a 128-byte id and hands it to every rank (any host channel)
 *   hrc_comm_init       collective over the ranks: joins the communicator on the CURRENT device.  libnccl.so.2 is
 *                       loaded at run time (dlopen); libhrc.so does not link against it
 *   hrc_comm_enable_p2p collective: maps every rank's receive buffer into every other rank (CUDA IPC over NVLink) for
 *                       HRC_TRANSPORT_P2P; max_keys = the largest n_rows * k a later call will exchange.  Returns 5
 *                       ON EVERY RANK when any rank cannot (no peer access, IPC not permitted): the ranks agree on the
 *                       verdict inside the call, and the NCCL transport stays usable
 * Transports of the exchange:
 *   HRC_TRANSPORT_NCCL  ncclAllGather of n_rows * k keys per rank, then the merge kernel
 *   HRC_TRANSPORT_P2P   a push kernel STORES this rank's keys into every peer's receive buffer and releases a sequence
 *                       flag (system scope); the merge kernel acquires the world's flags and merges.  No collective
 *                       launch; the exchange is part of the producer and the consumer kernels.  For ONE query,
 *                       hrc_sharded_search goes further: the search's own final selection kernel stores its top-k into
 *                       the peers, waits for theirs and emits the global top-k — two launches, like a local search.
 */
#define HRC_COMM_ID_BYTES 128
#define HRC_TRANSPORT_NCCL 0
#define HRC_TRANSPORT_P2P 1
typedef struct hrc_comm hrc_comm_t;
int hrc_comm_unique_id(void* id_out);
int hrc_comm_init(const void* unique_id, int world, int rank, hrc_comm_t** comm_out);
int hrc_comm_enable_p2p(hrc_comm_t* comm, int max_keys, void* stream);
int hrc_comm_world(const hrc_comm_t* comm);
int hrc_comm_rank(const hrc_comm_t* comm);
int hrc_comm_destroy(hrc_comm_t* comm);

/*
 * All-gather of every rank's sorted local keys + merge: d_keys_out[r] = the k best of the world * k keys of row r,
 * identical on every rank (keys are totally ordered, so this equals the single-GPU result bit for bit).
 *   d_local_keys : uint64 [n_rows][k] (0 = empty slot)
 *   d_workspace  : hrc_allgather_merge_workspace_bytes(world, n_rows, k) bytes (NCCL transport only)
 *   d_ids_out / d_scores_out : optional unpacked result
 */
size_t hrc_allgather_merge_workspace_bytes(int world, int n_rows, int k);
int hrc_allgather_merge_topk(hrc_comm_t* comm, const uint64_t* d_local_keys, int n_rows, int k, int transport,
                             void* d_workspace, size_t workspace_bytes, uint64_t* d_keys_out, int32_t* d_ids_out,
                             float* d_scores_out, void* stream);

/*
 * Document-sharded search in one call per rank: hrc_search over this rank's shard (global ids through id_base), the
 * exchange above, the merge.  hrc_sharded_search_host is its host-buffer form (H2D of the fp32 queries, fp32 -> bf16,
 * the sharded search, D2H of ids and scores; the caller synchronises the stream).
 */
size_t hrc_sharded_search_workspace_bytes(int world, int64_t n_docs, int64_t total_tokens, int n_queries, int lq, int k,
                                          int path);
int hrc_sharded_search(hrc_comm_t* comm, int transport, const void* d_tokens, const int64_t* d_offsets, int64_t n_docs,
                       int64_t total_tokens, const void* d_queries, int n_queries, int lq, int k, int32_t id_base,
                       void* d_workspace, size_t workspace_bytes, uint64_t* d_keys_out, int32_t* d_ids_out,
                       float* d_scores_out, int path, void* stream);
size_t hrc_sharded_search_host_workspace_bytes(int world, int64_t n_docs, int64_t total_tokens, int n_queries, int lq,
                                               int k, int path);
int hrc_sharded_search_host(hrc_comm_t* comm, int transport, const void* d_tokens, const int64_t* d_offsets,
                            int64_t n_docs, int64_t total_tokens, const float* h_queries, int n_queries, int lq, int k,
                            int32_t id_base, void* d_workspace, size_t workspace_bytes, int32_t* h_ids_out,
                            float* h_scores_out, int path, void* stream);

/*
 * Document-sharded HybridRetriever.retrieve (local_rag_complete.py:894-935) for a batch of queries, one call per rank:
 * local ColBERT top-colbert_k -> exchange + merge (the GLOBAL ColBERT list, :908-911) -> RRF with the BM25 lists on
 * global ids (:914-916; every rank computes the same fusion) -> each rank scores the candidates it OWNS -> exchange of
 * (score, candidate position) keys + merge (:926-929).  Results are identical on every rank and equal to
 * hrc_hybrid_retrieve over the unsharded store bit for bit.
 *   n_docs_global : documents of the whole corpus (ids outside [0, n_docs_global) score -inf, as on one GPU)
 */
size_t hrc_sharded_hybrid_workspace_bytes(int world, int64_t n_docs, int64_t total_tokens, int n_queries, int lq,
                                          int colbert_k, int n_candidates, int final_k, int path);
int hrc_sharded_hybrid_retrieve(hrc_comm_t* comm, int transport, const void* d_tokens, const int64_t* d_offsets,
                                int64_t n_docs, int64_t total_tokens, int64_t n_docs_global, const void* d_queries,
                                int n_queries, int lq, const int32_t* d_bm25_ids, int n_bm25, int colbert_k, int rrf_k,
                                int n_candidates, int final_k, int32_t id_base, void* d_workspace, size_t workspace_bytes,
                                int32_t* d_ids_out, float* d_scores_out, int path, void* stream);

/*
 * Streamed transfer between a file and device memory (the native on-disk store: tokens.bf16.bin holds the packed bf16
 * token rows of the whole corpus; a rank reads only the byte range of its document shard).  Replaces the persistence
 * half of JinaColBERTRetriever.index / load, local_rag_complete.py:742-753 (torch.save / torch.load of one dense tensor
 * through host memory).  Two pinned staging buffers of chunk_bytes (0 = 256 MiB): while one chunk travels (ONE
 * cudaMemcpyAsync per chunk) the next is read / written with pread / pwrite, so host memory in use is 2 x chunk_bytes
 * whatever the shard size.  Synchronous: returns when the data is in place; *seconds_out (optional) = elapsed time.
 */
int hrc_store_read_file(const char* path, int64_t file_offset, int64_t n_bytes, void* d_dst, size_t chunk_bytes,
                        void* stream, double* seconds_out);
int hrc_store_write_file(const char* path, int64_t file_offset, int64_t n_bytes, const void* d_src, size_t chunk_bytes,
                         void* stream, double* seconds_out);

/*
 * Read-bandwidth probe (bench utility, not on the query path): streams `bytes` of device memory once with 16-byte
 * loads and folds them into *d_out (a uint32 the caller zeroes).  bench.py times it over the resident corpus to
 * state the MaxSim kernel's HBM fraction against a pure-READ peak as well as the read+write copy peak.
 */
int hrc_read_probe(const void* d_buf, size_t bytes, uint32_t* d_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HRC_H_ */
