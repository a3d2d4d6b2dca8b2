#!/usr/bin/env python
"""Secondary measurements for the BASELINE.json configs that are not the headline bench line:
C1 (rerank latency), ragged single query, C3 (batched, tensor-bound), C4 (hybrid pipeline).
One JSON line per config; CUDA-event timing, >= 3 warm-ups, inputs far larger than L2 where it matters."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import hybrid_rag_colbertv2_b200 as hrc  # noqa: E402
from hybrid_rag_colbertv2_b200 import _lib  # noqa: E402
from hybrid_rag_colbertv2_b200.synth import synth_queries, synth_store  # noqa: E402


def timed(fn, steps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


class Clocks:
    def __enter__(self):
        import subprocess
        self.p = subprocess.Popen(["nvidia-smi", "--id=0", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits",
                                   "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        return self

    def __exit__(self, *a):
        import statistics
        self.p.terminate()
        rows = [l.split(",") for l in self.p.stdout.read().strip().splitlines() if "," in l]
        sm = [float(r[0]) for r in rows] or [0.0]
        pw = [float(r[1]) for r in rows] or [0.0]
        self.sm_mhz, self.power_w, self.n = statistics.median(sm), statistics.median(pw), len(rows)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops_sustained"], d["bf16_tflops"]
    return 6650.0, 1400.0, 1590.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="c1,ragged,c3,c4")
    ap.add_argument("--docs", type=int, default=1_000_000)
    ap.add_argument("--c3-queries", type=int, default=256)
    args = ap.parse_args()
    which = args.configs.split(",")
    dev = torch.device("cuda:0")
    hbm, tf_sus, tf_burst = peaks()
    cfg = hrc.RAGConfig(device="cuda:0")

    if "c1" in which:
        store = synth_store(100_000, 32, 512, seed=11, device=dev)
        r = hrc.JinaColBERTRetriever(cfg)
        r.store = store
        q = synth_queries(1, 32, device=dev)
        g = torch.Generator().manual_seed(0)
        cand = torch.randint(0, store.n_docs, (1, 50), generator=g, dtype=torch.int32).to(dev)
        toks = int(store.lengths()[cand[0].long()].sum())
        for name, path in (("tc", _lib.PATH_TC), ("simt", _lib.PATH_SIMT)):
            r.config.maxsim_path = path
            ms_score = timed(lambda: _lib.maxsim_scores_ids(store.tokens, store.offsets, cand, q, path=path), 200)
            ms_full = timed(lambda: r.rerank_ids(q, cand, k=10), 200)
            print(json.dumps({"config": "C1 rerank 50 candidates (<=512 tok), 1 query x 32", "path": name,
                              "candidate_tokens": toks, "score_kernel_us": ms_score * 1e3,
                              "rerank_ids_us": ms_full * 1e3, "docs_per_s": 50 / (ms_full * 1e-3)}), flush=True)
        r.config.maxsim_path = _lib.PATH_AUTO
        # the reference's own CPU path for this config: dense fp32 [50, 512, 128] (padded, as its encoder returns it)
        # scored with torch on the host — true MaxSim (oracle port, padding masked) and the function as coded
        from oracle import maxsim_oracle as o
        torch.set_num_threads(os.cpu_count() or 1)
        lens = store.lengths()[cand[0].long()].cpu()
        off = store.offsets.cpu()
        dense = torch.zeros((50, 512, 128))
        for j, d in enumerate(cand[0].tolist()):
            dense[j, : int(lens[j])] = store.tokens[int(off[d]): int(off[d + 1])].float().cpu()
        qf = q[0].float().cpu()

        def cpu_time(fn, n=30):
            fn()
            t0 = time.perf_counter()
            for _ in range(n):
                fn()
            return (time.perf_counter() - t0) / n * 1e6
        us_true = cpu_time(lambda: torch.argsort(o.maxsim_dense(qf, dense, lens.tolist()), descending=True)[:10])
        us_lit = cpu_time(lambda: torch.argsort(o.literal_reference(qf, dense), descending=True)[:10])
        print(json.dumps({"config": "C1 on the host cores (torch CPU, dense fp32 [50,512,128])", "cores": torch.get_num_threads(),
                          "true_maxsim_rerank_us": us_true, "reference_literal_rerank_us": us_lit}), flush=True)
        del store, r

    if "ragged" in which or "c3" in which:
        store = synth_store(args.docs, 32, 512, seed=12, device=dev)
        r = hrc.JinaColBERTRetriever(cfg)
        r.store = store
        tokens = store.total_tokens
        if "ragged" in which:
            q = synth_queries(1, 32, device=dev)
            out = torch.empty((1, store.n_docs), dtype=torch.float32, device=dev)
            ms = timed(lambda: _lib.maxsim_scores(store.tokens, store.offsets, q, out=out), 10)
            ms_s = timed(lambda: r.search_keys(q, 100), 10)
            gbs = tokens * 256 / (ms * 1e-3) / 1e9
            print(json.dumps({"config": f"ragged single query: {args.docs} docs x U(32..512) tokens", "tokens": tokens,
                              "kernel_ms": ms, "search_ms": ms_s, "achieved_GBps": gbs, "frac_hbm_measured": gbs / hbm,
                              "docs_per_s": store.n_docs / (ms_s * 1e-3)}), flush=True)
        if "c3" in which:
            for nq in sorted({8, 64, args.c3_queries}):
                q = synth_queries(nq, 32, device=dev)
                out = torch.empty((nq, store.n_docs), dtype=torch.float32, device=dev)
                steps = 3 if nq >= 64 else 5
                with Clocks() as ck:
                    ms = timed(lambda: _lib.maxsim_scores(store.tokens, store.offsets, q, out=out), steps, warmup=3)
                ms_s = timed(lambda: r.search_keys(q, 100), steps, warmup=1)
                flops = 2.0 * 32 * 128 * nq * tokens
                tfs = flops / (ms * 1e-3) / 1e12
                print(json.dumps({"config": f"C3 batched: {nq} queries x 32 over {args.docs} docs x U(32..512)",
                                  "tokens": tokens, "kernel_ms": ms, "search_ms": ms_s, "useful_TFLOPs": tfs,
                                  "sm_mhz": ck.sm_mhz, "power_w": ck.power_w, "col_split": os.environ.get("HRC_TC_COL_SPLIT", "1"),
                                  "frac_of_clock_peak": tfs / (8192 * 148 * ck.sm_mhz * 1e6 / 1e12) if ck.sm_mhz else None,
                                  "frac_tensor_sustained": tfs / tf_sus, "frac_tensor_burst": tfs / tf_burst,
                                  "pairs_per_s": nq * store.n_docs / (ms_s * 1e-3),
                                  "corpus_GB": tokens * 256 / 1e9,
                                  "effective_GBps": tokens * 256 / (ms * 1e-3) / 1e9}), flush=True)
        del store, r
        torch.cuda.empty_cache()

    if "c4" in which:
        store = synth_store(args.docs, 128, 128, seed=20260102, device=dev)
        idx = hrc.DualIndexer(cfg)
        idx.colbert_retriever.store = store
        h = hrc.HybridRetriever(cfg, idx, None, verbose=False)
        n_queries = 1000
        queries = synth_queries(n_queries, 32, device=dev)
        g = torch.Generator().manual_seed(4)
        bm25 = torch.randint(0, store.n_docs, (n_queries, 100), generator=g, dtype=torch.int32).to(dev)
        # 30 % overlap with the ColBERT list is produced naturally only by real data; synthesise it
        col_ids, _ = idx.colbert_retriever.search_embeddings(queries[:64], 100)
        bm25[:64, :30] = col_ids[:, torch.randperm(100, generator=g)[:30]]
        for batch in (1, 8, 64):
            n = 32 if batch == 1 else (64 if batch == 8 else 256)
            def run():
                for b in range(0, n, batch):
                    h.retrieve_batch(queries[b:b + batch], bm25[b:b + batch], top_k_final=10)
            ms = timed(run, 2, warmup=1)
            print(json.dumps({"config": f"C4 hybrid pipeline (ColBERT top-100 -> RRF(60) with BM25 top-100 -> top-50 -> rerank "
                                        f"top-10) over {args.docs} docs x 128 tok", "query_batch": batch, "queries_timed": n,
                              "ms_per_query": ms / n, "queries_per_s": n / (ms * 1e-3)}), flush=True)


if __name__ == "__main__":
    main()
