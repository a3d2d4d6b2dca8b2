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


class ColBERTEncoder:
    """ColBERT-style token encoder behind the `.encode` signature the reference calls (local_rag_complete.py:735-739,
    :758-761, :782-783): a transformer backbone, a linear projection to 128 dimensions, L2 normalisation per token.

    What the reference gets from `SentenceTransformer("jinaai/jina-colbert-v2", trust_remote_code=True)` is a dense,
    padded `[N, Ld, 128]` tensor without a mask, so padding rows enter its scores.  This encoder returns
    what the packed store wants instead: for documents a LIST of `[len_i, 128]` tensors holding only real tokens
    (padding dropped by the attention mask, punctuation dropped as ColBERT does), for a query one `[query_maxlen, 128]`
    tensor padded with [MASK] tokens (ColBERT query augmentation).  The query is encoded once per call — the reference
    encodes it in `search` and again in `rerank`.

    `backbone` is a loaded `transformers` model (or a model id for `AutoModel.from_pretrained(..., trust_remote_code=True)`),
    `tokenizer` its tokenizer (or None to load it by id), `projection` an `nn.Linear(hidden, 128, bias=False)` (or None to
    use the backbone's own 128-wide output / a checkpoint's `linear` weights).  Weights cannot be downloaded offline, so
    the tests drive this class with a small randomly initialised backbone; it is NOT part of the measured hot path.
    """

    def __init__(self, backbone, tokenizer=None, projection=None, dim: int = DIM, query_maxlen: int = QUERY_TOKENS,
                 doc_maxlen: int = 512, device: Union[str, torch.device] = "cuda", batch_size: int = 32,
                 query_prefix: str = "", doc_prefix: str = "", skip_punctuation: bool = True):
        if isinstance(backbone, str):
            from transformers import AutoModel, AutoTokenizer
            name = backbone
            backbone = AutoModel.from_pretrained(name, trust_remote_code=True)
            tokenizer = tokenizer or AutoTokenizer.from_pretrained(name, trust_remote_code=True)
        if tokenizer is None:
            raise ValueError("ColBERTEncoder needs a tokenizer")
        self.device = torch.device(device)
        self.backbone = backbone.to(self.device).eval()
        self.tokenizer = tokenizer
        self.projection = projection.to(self.device).eval() if projection is not None else None
        self.dim, self.query_maxlen, self.doc_maxlen, self.batch_size = dim, query_maxlen, doc_maxlen, batch_size
        self.query_prefix, self.doc_prefix = query_prefix, doc_prefix
        self.skip_ids = set()
        if skip_punctuation:
            import string
            for ch in string.punctuation:
                ids = tokenizer(ch, add_special_tokens=False)["input_ids"]
                if len(ids) == 1:
                    self.skip_ids.add(int(ids[0]))

    @torch.no_grad()
    def _embed(self, input_ids: torch.Tensor, attention_mask: torch.Tensor) -> torch.Tensor:
        h = self.backbone(input_ids=input_ids.to(self.device), attention_mask=attention_mask.to(self.device)).last_hidden_state
        if self.projection is not None:
            h = self.projection(h)
        if h.shape[-1] != self.dim:
            h = h[..., : self.dim]                      # matryoshka-style truncation (jina-colbert-v2 supports 128 / 96 / 64)
        return torch.nn.functional.normalize(h.float(), dim=-1)

    def encode_query(self, query: str) -> torch.Tensor:
        """[query_maxlen, 128]: the query's tokens followed by [MASK] tokens that the model fills in (query augmentation)."""
        tok = self.tokenizer(self.query_prefix + query, truncation=True, max_length=self.query_maxlen, return_tensors="pt")
        ids, mask = tok["input_ids"][0], tok["attention_mask"][0]
        pad = self.query_maxlen - ids.numel()
        if pad > 0:
            mask_id = self.tokenizer.mask_token_id if self.tokenizer.mask_token_id is not None else self.tokenizer.pad_token_id
            ids = torch.cat([ids, torch.full((pad,), int(mask_id), dtype=ids.dtype)])
            mask = torch.cat([mask, torch.ones(pad, dtype=mask.dtype)])      # [MASK] tokens are attended to
        return self._embed(ids.unsqueeze(0), mask.unsqueeze(0))[0]

    def encode_documents(self, documents: Sequence[str]) -> List[torch.Tensor]:
        """One `[len_i, 128]` tensor per document: real tokens only (no padding rows, no punctuation)."""
        out: List[torch.Tensor] = []
        for b in range(0, len(documents), self.batch_size):
            texts = [self.doc_prefix + d for d in documents[b:b + self.batch_size]]
            tok = self.tokenizer(texts, truncation=True, max_length=self.doc_maxlen, padding=True, return_tensors="pt")
            emb = self._embed(tok["input_ids"], tok["attention_mask"])
            keep = tok["attention_mask"].bool()
            if self.skip_ids:
                skip = torch.tensor(sorted(self.skip_ids), dtype=tok["input_ids"].dtype)
                keep &= ~torch.isin(tok["input_ids"], skip)
            keep = keep.to(emb.device)
            for i in range(len(texts)):
                rows = emb[i][keep[i]]
                out.append(rows if rows.shape[0] > 0 else emb[i][:1])      # never an empty document
        return out

    def encode(self, sentences: Union[str, Sequence[str]], convert_to_tensor: bool = True,
               show_progress_bar: bool = False, is_query: bool = None, **_) -> Union[torch.Tensor, List[torch.Tensor]]:
        if isinstance(sentences, str):
            if is_query is None or is_query:
                return self.encode_query(sentences)
            return self.encode_documents([sentences])[0]
        return self.encode_documents(list(sentences))
