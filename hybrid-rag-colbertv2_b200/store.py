"""Packed, padding-free corpus token-embedding store (bf16 [total_tokens, 128] + CSR int64 offsets).

Replaces the reference's single dense fp32 tensor `corpus_embeddings` [N, Ld, D]
(local_rag_complete.py:735-739, persisted at :742-746, loaded at :748-753).  The reference keeps no
attention mask, so a dense tensor without lengths is taken as "every row is a real token".
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

DIM = 128


def _as_bf16_rows(x: torch.Tensor) -> torch.Tensor:
    if x.shape[-1] != DIM:
        raise ValueError(f"token embeddings must be {DIM}-dimensional, got {x.shape[-1]}")
    return x.to(torch.bfloat16)


def lengths_to_offsets(lengths: Union[Sequence[int], np.ndarray, torch.Tensor]) -> torch.Tensor:
    lens = torch.as_tensor(lengths, dtype=torch.int64).cpu()
    off = torch.zeros(lens.numel() + 1, dtype=torch.int64)
    torch.cumsum(lens, 0, out=off[1:])
    return off


def shard_doc_ranges(offsets: Union[torch.Tensor, np.ndarray], world_size: int) -> List[Tuple[int, int]]:
    """Contiguous document ranges per rank, balanced by TOKEN count.

    Rank r owns the documents whose first token lies in [T*r/W, T*(r+1)/W); every document belongs
    to exactly one rank, and the split depends only on (offsets, world_size).
    """
    off = np.asarray(offsets.cpu() if isinstance(offsets, torch.Tensor) else offsets, dtype=np.int64)
    n_docs = off.shape[0] - 1
    total = int(off[-1])
    bounds = [0]
    for r in range(1, world_size):
        target = (total * r) // world_size
        bounds.append(int(np.searchsorted(off[:n_docs], target, side="left")))
    bounds.append(n_docs)
    for i in range(1, len(bounds)):
        bounds[i] = max(bounds[i], bounds[i - 1])
    return [(bounds[r], bounds[r + 1]) for r in range(world_size)]


@dataclass
class PackedStore:
    """tokens: bf16 [T, 128] and offsets: int64 [n_docs + 1], both on `device`; ids are global via doc_id_base."""

    tokens: torch.Tensor
    offsets: torch.Tensor
    doc_id_base: int = 0

    def __post_init__(self):
        if self.tokens.dtype != torch.bfloat16 or self.tokens.dim() != 2 or self.tokens.shape[1] != DIM:
            raise ValueError("tokens must be bf16 [total_tokens, 128]")
        if self.offsets.dtype != torch.int64 or self.offsets.dim() != 1 or self.offsets.numel() < 1:
            raise ValueError("offsets must be int64 [n_docs + 1]")
        self.tokens = self.tokens.contiguous()
        self.offsets = self.offsets.contiguous()

    # ---- shape -------------------------------------------------------------------------------
    @property
    def n_docs(self) -> int:
        return self.offsets.numel() - 1

    @property
    def total_tokens(self) -> int:
        return int(self.tokens.shape[0])

    @property
    def device(self) -> torch.device:
        return self.tokens.device

    def __len__(self) -> int:
        return self.n_docs

    def lengths(self) -> torch.Tensor:
        return self.offsets[1:] - self.offsets[:-1]

    def nbytes(self) -> int:
        return self.tokens.numel() * 2 + self.offsets.numel() * 8

    def validate(self, check_values: bool = True) -> dict:
        """On-device integrity check (hrc_store_validate): CSR offsets well-formed and, with check_values, no NaN /
        infinite token value (one such row poisons the max of its document).  Raises ValueError; a CUDA store only.
        Run once after loading a file — the reference's `load` (:748-753) trusts the file."""
        from . import _lib
        return _lib.store_validate(self.tokens, self.offsets, check_values=check_values)

    # ---- construction ------------------------------------------------------------------------
    @staticmethod
    def _validate(offsets_cpu: torch.Tensor, total: int, allow_empty: bool) -> None:
        if int(offsets_cpu[0]) != 0 or int(offsets_cpu[-1]) != total:
            raise ValueError("offsets must start at 0 and end at total_tokens")
        d = offsets_cpu[1:] - offsets_cpu[:-1]
        if d.numel() and int(d.min()) < (0 if allow_empty else 1):
            raise ValueError("empty (zero-token) documents are rejected at index build "
                             "(MaxSim over no tokens is undefined); pass allow_empty=True to score them -inf")

    @classmethod
    def from_packed(cls, tokens: torch.Tensor, offsets: torch.Tensor, device: Union[str, torch.device] = "cuda",
                    doc_id_base: int = 0, allow_empty: bool = False) -> "PackedStore":
        off_cpu = offsets.detach().to("cpu", torch.int64)
        cls._validate(off_cpu, int(tokens.shape[0]), allow_empty)
        return cls(_as_bf16_rows(tokens).to(device), off_cpu.to(device), doc_id_base)

    @classmethod
    def from_ragged(cls, docs: Sequence[torch.Tensor], device: Union[str, torch.device] = "cuda",
                    allow_empty: bool = False) -> "PackedStore":
        lens = [int(d.shape[0]) for d in docs]
        off = lengths_to_offsets(lens)
        if len(docs):
            tokens = torch.cat([_as_bf16_rows(d.reshape(-1, d.shape[-1])) for d in docs], 0)
        else:
            tokens = torch.zeros((0, DIM), dtype=torch.bfloat16)
        return cls.from_packed(tokens, off, device, allow_empty=allow_empty)

    @classmethod
    def from_dense(cls, embeddings: torch.Tensor, lengths: Optional[Union[Sequence[int], torch.Tensor]] = None,
                   device: Union[str, torch.device] = "cuda", allow_empty: bool = False) -> "PackedStore":
        """Dense [N, Ld, D] (+ optional per-document lengths) -> packed.  Padding rows are dropped."""
        if embeddings.dim() == 2:  # the reference treats a 2-D tensor as ONE document (:816-817)
            embeddings = embeddings.unsqueeze(0)
        if embeddings.dim() != 3:
            raise ValueError("dense embeddings must be [N, Ld, D]")
        n, ld, _ = embeddings.shape
        if lengths is None:
            off = torch.arange(0, (n + 1) * ld, ld, dtype=torch.int64) if ld > 0 else torch.zeros(n + 1, dtype=torch.int64)
            tokens = _as_bf16_rows(embeddings.reshape(n * ld, -1))
        else:
            lens = torch.as_tensor(lengths, dtype=torch.int64).cpu()
            if lens.numel() != n or (lens.numel() and (int(lens.max()) > ld or int(lens.min()) < 0)):
                raise ValueError("lengths must have one entry in [0, Ld] per document")
            off = lengths_to_offsets(lens)
            mask = torch.arange(ld).unsqueeze(0) < lens.unsqueeze(1)
            tokens = _as_bf16_rows(embeddings.to("cpu")[mask])
        return cls.from_packed(tokens, off, device, allow_empty=allow_empty)

    def to_dense(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """(fp32 [N, Ld_max, 128] zero-padded, int64 lengths [N]) on the CPU — the reference's layout (:735-739)."""
        lens = self.lengths().cpu()
        n, ld = self.n_docs, int(lens.max()) if self.n_docs else 0
        dense = torch.zeros((n, ld, DIM), dtype=torch.float32)
        mask = torch.arange(ld).unsqueeze(0) < lens.unsqueeze(1)
        dense[mask] = self.tokens.float().cpu()
        return dense, lens

    # ---- sharding ----------------------------------------------------------------------------
    def shard(self, rank: int, world_size: int, device: Optional[Union[str, torch.device]] = None) -> "PackedStore":
        d0, d1 = shard_doc_ranges(self.offsets, world_size)[rank]
        off = self.offsets[d0:d1 + 1]
        t0, t1 = int(off[0]), int(off[-1])
        dev = self.device if device is None else device
        return PackedStore(self.tokens[t0:t1].to(dev), (off - off[0]).to(dev), self.doc_id_base + d0)

    # ---- persistence (native format; the reference-compatible index.pt lives in retriever.py) --
    def save(self, path: str, chunk_bytes: int = 0) -> None:
        """Native store: tokens.bf16.bin (packed rows) + offsets.i64.bin (CSR) + meta.json.  A CUDA store is streamed
        to disk through two pinned staging buffers (hrc_store_write_file): no whole-store host copy."""
        os.makedirs(path, exist_ok=True)
        tok_file = os.path.join(path, "tokens.bf16.bin")
        if self.tokens.is_cuda:
            from . import _lib
            open(tok_file, "wb").close()
            self.last_io_seconds = _lib.store_write_file(tok_file, 0, self.tokens, chunk_bytes) if self.total_tokens else 0.0
        else:
            self.tokens.view(torch.int16).numpy().tofile(tok_file)
        self.offsets.cpu().numpy().tofile(os.path.join(path, "offsets.i64.bin"))
        with open(os.path.join(path, "meta.json"), "w") as f:
            json.dump({"format": "hrc-packed-v1", "dim": DIM, "n_docs": self.n_docs,
                       "total_tokens": self.total_tokens, "doc_id_base": self.doc_id_base}, f)

    @classmethod
    def load(cls, path: str, device: Union[str, torch.device] = "cuda", rank: int = 0, world_size: int = 1,
             allow_empty: bool = False, chunk_bytes: int = 0) -> "PackedStore":
        """Load (a document shard of) a native store straight to `device`, reading only that shard's bytes.  On a CUDA
        device the bytes stream file -> pinned staging buffer -> the rank's device slice (hrc_store_read_file, one
        cudaMemcpyAsync per >= 256 MB chunk): host memory in use is two staging buffers, never the shard."""
        with open(os.path.join(path, "meta.json")) as f:
            meta = json.load(f)
        if meta.get("format") != "hrc-packed-v1" or meta.get("dim") != DIM:
            raise ValueError(f"{path}: not an hrc-packed-v1 store")
        off = torch.from_numpy(np.fromfile(os.path.join(path, "offsets.i64.bin"), dtype=np.int64))
        cls._validate(off, meta["total_tokens"], allow_empty)
        d0, d1 = shard_doc_ranges(off, world_size)[rank]
        t0, t1 = int(off[d0]), int(off[d1])
        tok_file = os.path.join(path, "tokens.bf16.bin")
        dev = torch.device(device)
        if dev.type == "cuda":
            from . import _lib
            tokens = torch.empty((t1 - t0, DIM), dtype=torch.bfloat16, device=dev)
            secs = _lib.store_read_file(tok_file, t0 * DIM * 2, tokens, chunk_bytes) if t1 > t0 else 0.0
            store = cls(tokens, (off[d0:d1 + 1] - off[d0]).to(dev), meta.get("doc_id_base", 0) + d0)
            store.last_io_seconds = secs
            store.validate()          # a truncated or corrupted token file must not be searched
            return store
        raw = np.fromfile(tok_file, dtype=np.int16, count=(t1 - t0) * DIM, offset=t0 * DIM * 2)
        tokens = torch.from_numpy(raw).view(torch.bfloat16).reshape(t1 - t0, DIM)
        return cls(tokens.to(device), (off[d0:d1 + 1] - off[d0]).to(device), meta.get("doc_id_base", 0) + d0)
