#!/usr/bin/env python
"""End-to-end demo on one B200: index a small corpus, search, rerank and run the three-stage hybrid retrieval through
the reference-shaped classes (no model weights needed: the deterministic SyntheticEncoder stands in for
jinaai/jina-colbert-v2, so lexical overlap between query and document gives a high MaxSim).

    python examples/demo.py
"""
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("HRC_ENCODER", "synthetic")

import hybrid_rag_colbertv2_b200 as hrc  # noqa: E402

TOPICS = ["late interaction retrieval scores every query token against every document token",
          "reciprocal rank fusion merges lexical and semantic rankings",
          "tensor memory holds the accumulators of the fifth generation tensor cores",
          "the tensor memory accelerator streams tiles into shared memory",
          "bm25 is a lexical ranking function based on term frequency",
          "high bandwidth memory feeds the streaming multiprocessors"]


def main():
    corpus = [f"{TOPICS[i % len(TOPICS)]} note {i} about item {i * 7 % 13}" for i in range(600)]
    with tempfile.TemporaryDirectory() as tmp:
        cfg = hrc.RAGConfig(colbert_index_path=os.path.join(tmp, "colbert"), colbert_top_k=50, bm25_top_k=50,
                            rerank_candidates=20, final_top_k=5)
        indexer = hrc.DualIndexer(cfg)
        indexer.build_colbert_index(corpus)                       # JinaColBERTRetriever.index -> packed bf16 store + index.hrc.pt
        retriever = indexer.colbert_retriever
        print(f"store: {retriever.store.n_docs} documents, {retriever.store.total_tokens} tokens, "
              f"{retriever.store.nbytes() / 1e6:.2f} MB on {retriever.store.device}")

        query = "how does reciprocal rank fusion merge rankings"
        print("\nsearch(query, k=5):")
        for r in retriever.search(query=query, k=5):
            print(f"  doc {r['document_id']:4d}  score {r['score']:.4f}  {r['text'][:70]}")

        docs = [corpus[i] for i in (1, 7, 13, 2, 3, 4)]
        print("\nrerank(query, 6 documents, k=3):")
        for r in retriever.rerank(query=query, documents=docs, k=3):
            print(f"  rank {r['rank']}  input #{r['result_index']}  score {r['score']:.4f}  {r['text'][:60]}")

        def bm25(q, k):                                            # stand-in for bm25s (third-party, out of scope)
            words = set(q.lower().split())
            hits = sorted(range(len(corpus)), key=lambda i: -len(words & set(corpus[i].split())))[:k]
            return [{"chunk_id": i, "score": float(len(words & set(corpus[i].split()))), "source": "bm25"} for i in hits]

        hybrid = hrc.HybridRetriever(cfg, indexer, None, bm25_search=bm25)
        print("\nretrieve(query): BM25 top-50 + ColBERT top-50 -> RRF -> top-20 -> rerank top-5")
        for r in hybrid.retrieve(query):
            print(f"  rank {r['rank']}  chunk {r['chunk_id']:4d}  score {r['score']:.4f}  {r['text'][:60]}")
        print(f"\nkernels launched: {hrc._lib.launch_count()}")


if __name__ == "__main__":
    main()
