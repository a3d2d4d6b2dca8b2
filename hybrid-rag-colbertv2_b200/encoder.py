"""Encoder hook.

The reference wraps `SentenceTransformer("jinaai/jina-colbert-v2", trust_remote_code=True)`
(local_rag_complete.py:720-724) and calls `.encode(text_or_list, convert_to_tensor=True)` (:735-739,
:758-761, :782-783).  Neither the package nor the weights exist offline, and the encoder is not part
of the hot path, so the retriever takes any object with that `.encode` signature.  The default is
this deterministic stand-in: token embeddings are a pure function of (text, token position), so a
text re-encoded at rerank time (:783) reproduces the vectors stored at index time.
"""
from __future__ import annotations

import hashlib
from typing import List, Sequence, Union

import numpy as np
import torch

DIM = 128
QUERY_TOKENS = 32  # ColBERT query length after [MASK] augmentation


def _text_seed(text: str) -> int:
    return int.from_bytes(hashlib.blake2b(text.encode("utf-8"), digest_size=8).digest(), "little")


def _unit_rows(seed: int, n_rows: int) -> torch.Tensor:
    rng = np.random.Generator(np.random.Philox(key=seed & (2**64 - 1)))
    x = rng.standard_normal((n_rows, DIM), dtype=np.float32)
    x /= np.maximum(np.linalg.norm(x, axis=1, keepdims=True), 1e-12)
    return torch.from_numpy(x)


class SyntheticEncoder:
    """Deterministic text -> L2-normalised [n_tokens, 128] fp32 token embeddings.

    Words shared between a query and a document produce correlated token vectors (a document token
    is `normalise(word_vector + 0.35 * position_noise)`), so lexical overlap yields high MaxSim, which
    makes end-to-end examples behave sensibly without any model weights.
    """

    def __init__(self, doc_tokens: int = 0, query_tokens: int = QUERY_TOKENS, max_doc_tokens: int = 512):
        self.doc_tokens = doc_tokens          # 0: one token per whitespace word (ragged output)
        self.query_tokens = query_tokens
        self.max_doc_tokens = max_doc_tokens

    def _encode_one(self, text: str, n_tokens: int) -> torch.Tensor:
        words = text.lower().split() or [""]
        if n_tokens <= 0:
            n_tokens = min(len(words), self.max_doc_tokens)
        rows = []
        base = _text_seed(text)
        noise = _unit_rows(base, n_tokens)
        for i in range(n_tokens):
            w = words[i % len(words)]
            rows.append(_unit_rows(_text_seed("w:" + w), 1)[0])
        x = torch.stack(rows) + 0.35 * noise
        return torch.nn.functional.normalize(x, dim=1)

    def encode(self, sentences: Union[str, Sequence[str]], convert_to_tensor: bool = True,
               show_progress_bar: bool = False, is_query: bool = None, **_) -> Union[torch.Tensor, List[torch.Tensor]]:
        if isinstance(sentences, str):
            query = True if is_query is None else is_query
            return self._encode_one(sentences, self.query_tokens if query else self.doc_tokens)
        docs = [self._encode_one(s, self.doc_tokens) for s in sentences]
        if self.doc_tokens > 0:
            return torch.stack(docs) if docs else torch.zeros((0, self.doc_tokens, DIM))
        return docs  # ragged: list of [len_i, 128]
