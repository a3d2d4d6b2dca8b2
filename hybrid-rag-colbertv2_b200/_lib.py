"""ctypes binding of libhrc.so (C ABI declared in include/hrc.h).

There is deliberately NO fallback: if the shared library is missing, or the device is not a CUDA
sm_100 GPU, every operation raises.  PyTorch is used only for device memory and streams.
"""
from __future__ import annotations

import ctypes
import os
import threading
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HRC_LIB_PATH") or os.path.join(_HERE, "libhrc.so")   # override: A/B of library builds

PATH_AUTO, PATH_SIMT, PATH_TC, PATH_TC_DM = 0, 1, 2, 3
DIM = 128
MAX_TOPK = 2048
TC_MAX_LQ = 32          # query tokens per slot on the tensor-core path; up to 8 slots (lq <= 256)
ABI_VERSION = 200

#: every symbol include/hrc.h declares: name -> (restype, argtypes)
_c = ctypes
_P, _I, _I32, _I64, _SZ, _U64 = _c.c_void_p, _c.c_int, _c.c_int32, _c.c_int64, _c.c_size_t, _c.c_uint64
SYMBOLS = {
    "hrc_version": (_I, []),
    "hrc_last_error": (_c.c_char_p, []),
    "hrc_launch_count": (_U64, []),
    "hrc_set_watchdog_ms": (None, [_U64]),
    "hrc_trace_enable": (_I, [_I]),
    "hrc_trace_collect": (_I, [_P, _I]),
    "hrc_store_register": (_I, [_P, _I64]),
    "hrc_store_release": (None, [_P]),
    "hrc_store_validate": (_I, [_P, _P, _I64, _I64, _I, _P, _SZ, _P, _P]),
    "hrc_maxsim_workspace_bytes": (_SZ, [_I64, _I, _I]),
    "hrc_maxsim_scores": (_I, [_P, _P, _I64, _I64, _P, _I, _I, _P, _I, _P, _SZ, _P]),
    "hrc_maxsim_scores_ids": (_I, [_P, _P, _I64, _I64, _P, _I, _P, _I, _I, _P, _I, _P, _SZ, _P]),
    "hrc_meanpool_cosine_scores": (_I, [_P, _P, _I64, _I64, _P, _I, _I, _P, _P]),
    "hrc_search_workspace_bytes": (_SZ, [_I64, _I64, _I, _I, _I, _I]),
    "hrc_search": (_I, [_P, _P, _I64, _I64, _P, _I, _I, _I, _I32, _P, _SZ, _P, _P, _P, _I, _P]),
    "hrc_search_host_workspace_bytes": (_SZ, [_I64, _I64, _I, _I, _I, _I]),
    "hrc_search_host": (_I, [_P, _P, _I64, _I64, _P, _I, _I, _I, _I32, _P, _SZ, _P, _P, _I, _P]),
    "hrc_rerank_workspace_bytes": (_SZ, [_I, _I, _I, _I]),
    "hrc_rerank": (_I, [_P, _P, _I64, _I64, _P, _I, _P, _I, _I, _I, _P, _SZ, _P, _P, _P, _P, _I, _P]),
    "hrc_hybrid_retrieve_workspace_bytes": (_SZ, [_I64, _I64, _I, _I, _I, _I, _I, _I]),
    "hrc_hybrid_retrieve": (_I, [_P, _P, _I64, _I64, _P, _I, _I, _P, _I, _I, _I, _I, _I, _I32, _P, _SZ, _P, _P, _I, _P]),
    "hrc_topk_workspace_bytes": (_SZ, [_I64, _I, _I]),
    "hrc_topk": (_I, [_P, _P, _I64, _I, _I, _I32, _P, _P, _SZ, _P]),
    "hrc_topk_merge": (_I, [_P, _I, _I, _I, _P, _P]),
    "hrc_keys_unpack": (_I, [_P, _I64, _P, _P, _P]),
    "hrc_rrf_fuse": (_I, [_P, _I, _P, _I, _I, _I, _I, _P, _P, _P, _P]),
    "hrc_synth_tokens": (_I, [_P, _I64, _I64, _U64, _P]),
    "hrc_read_probe": (_I, [_P, _SZ, _P, _P]),
    "hrc_store_read_file": (_I, [_c.c_char_p, _I64, _I64, _P, _SZ, _P, _P]),
    "hrc_store_write_file": (_I, [_c.c_char_p, _I64, _I64, _P, _SZ, _P, _P]),
    "hrc_comm_unique_id": (_I, [_P]),
    "hrc_comm_init": (_I, [_P, _I, _I, _P]),
    "hrc_comm_enable_p2p": (_I, [_P, _I, _P]),
    "hrc_comm_world": (_I, [_P]),
    "hrc_comm_rank": (_I, [_P]),
    "hrc_comm_destroy": (_I, [_P]),
    "hrc_allgather_merge_workspace_bytes": (_SZ, [_I, _I, _I]),
    "hrc_allgather_merge_topk": (_I, [_P, _P, _I, _I, _I, _P, _SZ, _P, _P, _P, _P]),
    "hrc_sharded_search_workspace_bytes": (_SZ, [_I, _I64, _I64, _I, _I, _I, _I]),
    "hrc_sharded_search": (_I, [_P, _I, _P, _P, _I64, _I64, _P, _I, _I, _I, _I32, _P, _SZ, _P, _P, _P, _I, _P]),
    "hrc_sharded_search_host_workspace_bytes": (_SZ, [_I, _I64, _I64, _I, _I, _I, _I]),
    "hrc_sharded_search_host": (_I, [_P, _I, _P, _P, _I64, _I64, _P, _I, _I, _I, _I32, _P, _SZ, _P, _P, _I, _P]),
    "hrc_sharded_hybrid_workspace_bytes": (_SZ, [_I, _I64, _I64, _I, _I, _I, _I, _I, _I]),
    "hrc_sharded_hybrid_retrieve": (_I, [_P, _I, _P, _P, _I64, _I64, _I64, _P, _I, _I, _P, _I, _I, _I, _I, _I, _I32, _P, _SZ,
                                         _P, _P, _I, _P]),
}
TRANSPORT_NCCL, TRANSPORT_P2P = 0, 1
COMM_ID_BYTES = 128

_lib: Optional[ctypes.CDLL] = None


class HrcError(RuntimeError):
    """A libhrc call returned a non-zero status."""


def load() -> ctypes.CDLL:
    """Load libhrc.so (once) and type every exported symbol.  Raises if the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HrcError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU / PyTorch fallback for the MaxSim path)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the ABI and the header disagree
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def _check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().hrc_last_error().decode("utf-8", "replace")
        raise HrcError(f"{what} failed (status {rc}): {msg}")


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)      # 0.2 us instead of 2 us per call


def _stream(device: torch.device) -> int:
    """Handle (cudaStream_t) of torch's current stream on `device`."""
    if _raw_stream is not None and device.index is not None:
        return _raw_stream(device.index)
    return torch.cuda.current_stream(device).cuda_stream


class _NoSwitch:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NO_SWITCH = _NoSwitch()


def _on(device: torch.device):
    """Context manager making `device` current for the C call (kernel launches go to the current device); a no-op —
    and no cudaSetDevice pair — when it already is, which is every call of a one-process-per-GPU program."""
    if device.index is None or device.index == torch.cuda.current_device():
        return _NO_SWITCH
    return torch.cuda.device(device)


def _require_cuda(*tensors: torch.Tensor) -> torch.device:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise HrcError("libhrc operates on CUDA tensors only (no CPU fallback)")
        if not t.is_contiguous():
            raise HrcError("libhrc needs contiguous tensors")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise HrcError(f"tensors on different devices: {dev} vs {t.device}")
    assert dev is not None
    return dev


def version() -> int:
    return load().hrc_version()


def launch_count() -> int:
    return int(load().hrc_launch_count())


def set_watchdog_ms(ms: int) -> None:
    """mbarrier waits inside the tensor-core kernels trap after `ms` without progress (0 = never)."""
    load().hrc_set_watchdog_ms(int(ms))


def trace_enable(capacity: int) -> None:
    """Bracket the next `capacity` scoring-kernel launches with CUDA events (0 switches tracing off)."""
    _check(load().hrc_trace_enable(int(capacity)), "hrc_trace_enable")


def trace_collect(max_n: int = 1 << 16):
    """Device times (ms, launch order) of the scoring kernels traced since the last collect; resets the trace."""
    buf = (ctypes.c_float * max_n)()
    n = load().hrc_trace_collect(buf, max_n)
    if n < 0:
        _check(1, "hrc_trace_collect")
    return [float(buf[i]) for i in range(n)]


def store_register(tokens: torch.Tensor) -> None:
    """Pre-encode the TMA descriptors of a token store (optional; the first scoring call does it otherwise)."""
    _require_cuda(tokens)
    if tokens.shape[0] == 0:
        return
    with _on(tokens.device):
        _check(load().hrc_store_register(_ptr(tokens), int(tokens.shape[0])), "hrc_store_register")


def store_release(tokens: torch.Tensor) -> None:
    """Drop the cached TMA descriptors of a token store (call before freeing it)."""
    if _lib is not None and tokens is not None and tokens.is_cuda:
        _lib.hrc_store_release(_ptr(tokens))


def store_validate(tokens: torch.Tensor, offsets: torch.Tensor, *, check_values: bool = True) -> dict:
    """Integrity check of a packed store (hrc_store_validate), run once at load / build time: raises ValueError when the
    CSR offsets are malformed or (check_values) a token value is NaN / infinite.  Returns
    {'longest_doc_tokens': ...} for a sound store.  Synchronises the current stream."""
    dev = _require_cuda(tokens, offsets)
    _check_store(tokens, offsets)
    ws = torch.empty(4, dtype=torch.int64, device=dev)
    rep = (ctypes.c_int64 * 4)()
    with _on(dev):
        rc = load().hrc_store_validate(_ptr(tokens), _ptr(offsets), offsets.numel() - 1, int(tokens.shape[0]),
                                       1 if check_values else 0, _ptr(ws), 32, rep, _stream(dev))
    if rc in (3, 4):
        raise ValueError(load().hrc_last_error().decode())
    _check(rc, "hrc_store_validate")
    return {"longest_doc_tokens": int(rep[3])}


class Workspace:
    """Reusable 256-byte-aligned device scratch that only ever grows: no per-call allocation once warm.  One buffer PER
    STREAM (keyed by the stream handle the call is issued on): two threads driving the same retriever on different
    streams never share scratch, and a buffer is only ever replaced from the stream whose kernels use it, which is
    what makes handing the old one back to torch's stream-ordered caching allocator safe."""

    def __init__(self):
        self.bufs: dict = {}

    @property
    def buf(self) -> Optional[torch.Tensor]:
        """The largest buffer held (diagnostics / tests)."""
        return max(self.bufs.values(), key=lambda b: b.numel(), default=None)

    def get(self, dev, nbytes: int, stream: int = 0) -> torch.Tensor:
        buf = self.bufs.get(stream)
        if buf is None or buf.device != dev or buf.numel() < nbytes:
            buf = self.bufs[stream] = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=dev)   # 512-B aligned
        return buf


def _ws(workspace: Optional[Workspace], dev, nbytes: int, stream: int = 0):
    """(pointer, bytes, keep-alive) of a scratch buffer of at least nbytes for calls issued on `stream`;
    (None, 0, None) when none is needed."""
    if nbytes == 0:
        return None, 0, None
    buf = (workspace or Workspace()).get(dev, nbytes, stream)
    return buf.data_ptr(), buf.numel(), buf


def maxsim_workspace_bytes(n_items: int, n_queries: int, lq: int) -> int:
    return int(load().hrc_maxsim_workspace_bytes(n_items, n_queries, lq))


def _check_store(tokens, offsets):
    assert tokens.dtype == torch.bfloat16 and tokens.dim() == 2 and tokens.shape[1] == DIM
    assert offsets.dtype == torch.int64 and offsets.dim() == 1 and offsets.numel() >= 1


def _check_queries(queries):
    assert queries.dtype == torch.bfloat16 and queries.dim() == 3 and queries.shape[2] == DIM


def maxsim_scores(tokens: torch.Tensor, offsets: torch.Tensor, queries: torch.Tensor, *,
                  path: int = PATH_AUTO, out: Optional[torch.Tensor] = None,
                  workspace: Optional[Workspace] = None) -> torch.Tensor:
    """scores[q, d] = sum_i max_{t in doc d} <queries[q, i], tokens[t]>  ->  fp32 [n_queries, n_docs]."""
    dev = _require_cuda(tokens, offsets, queries)
    _check_store(tokens, offsets)
    _check_queries(queries)
    n_docs = offsets.numel() - 1
    nq, lq = int(queries.shape[0]), int(queries.shape[1])
    if out is None:
        out = torch.empty((nq, n_docs), dtype=torch.float32, device=dev)
    else:
        assert out.shape == (nq, n_docs) and out.dtype == torch.float32 and out.is_contiguous()
    st = _stream(dev)
    ws_ptr, ws_bytes, _keep = _ws(workspace, dev, 0 if path == PATH_SIMT else maxsim_workspace_bytes(n_docs, nq, lq), st)
    with _on(dev):
        rc = load().hrc_maxsim_scores(_ptr(tokens), _ptr(offsets), n_docs, int(tokens.shape[0]), _ptr(queries),
                                      nq, lq, _ptr(out), path, ws_ptr, ws_bytes, st)
    _check(rc, "hrc_maxsim_scores")
    return out


def meanpool_cosine_scores(tokens: torch.Tensor, offsets: torch.Tensor, queries: torch.Tensor, *,
                           out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """The reference's `_maxsim_score` exactly as coded (local_rag_complete.py:821-829): cosine of the
    mean-pooled query and document token vectors -> fp32 [n_queries, n_docs]."""
    dev = _require_cuda(tokens, offsets, queries)
    _check_store(tokens, offsets)
    _check_queries(queries)
    n_docs = offsets.numel() - 1
    nq, lq = int(queries.shape[0]), int(queries.shape[1])
    if out is None:
        out = torch.empty((nq, n_docs), dtype=torch.float32, device=dev)
    else:
        assert out.shape == (nq, n_docs) and out.dtype == torch.float32 and out.is_contiguous()
    with _on(dev):
        rc = load().hrc_meanpool_cosine_scores(_ptr(tokens), _ptr(offsets), n_docs, int(tokens.shape[0]),
                                               _ptr(queries), nq, lq, _ptr(out), _stream(dev))
    _check(rc, "hrc_meanpool_cosine_scores")
    return out


def maxsim_scores_ids(tokens: torch.Tensor, offsets: torch.Tensor, cand_ids: torch.Tensor,
                      queries: torch.Tensor, *, path: int = PATH_AUTO,
                      workspace: Optional[Workspace] = None) -> torch.Tensor:
    """scores[q, j] = maxsim(queries[q], doc cand_ids[q, j])  ->  fp32 [n_queries, n_cand]."""
    dev = _require_cuda(tokens, offsets, cand_ids, queries)
    assert cand_ids.dtype == torch.int32 and cand_ids.dim() == 2 and cand_ids.shape[0] == queries.shape[0]
    _check_store(tokens, offsets)
    _check_queries(queries)
    n_docs = offsets.numel() - 1
    nq, lq = int(queries.shape[0]), int(queries.shape[1])
    n_cand = int(cand_ids.shape[1])
    out = torch.empty((nq, n_cand), dtype=torch.float32, device=dev)
    st = _stream(dev)
    ws_ptr, ws_bytes, _keep = _ws(workspace, dev, 0 if path == PATH_SIMT else maxsim_workspace_bytes(n_cand, nq, lq), st)
    with _on(dev):
        rc = load().hrc_maxsim_scores_ids(_ptr(tokens), _ptr(offsets), n_docs, int(tokens.shape[0]), _ptr(cand_ids),
                                          n_cand, _ptr(queries), nq, lq, _ptr(out), path, ws_ptr, ws_bytes, st)
    _check(rc, "hrc_maxsim_scores_ids")
    return out


def search(tokens: torch.Tensor, offsets: torch.Tensor, queries: torch.Tensor, k: int, *, id_base: int = 0,
           path: int = PATH_AUTO, workspace: Optional[Workspace] = None, unpack: bool = True):
    """Fused MaxSim + top-k (+ unpack) in one C call.  Returns (keys int64 [nq,k], ids int32 | None, scores fp32 | None)."""
    dev = _require_cuda(tokens, offsets, queries)
    _check_store(tokens, offsets)
    _check_queries(queries)
    n_docs = offsets.numel() - 1
    nq, lq = int(queries.shape[0]), int(queries.shape[1])
    st = _stream(dev)
    ws_ptr, ws_bytes, _keep = _ws(workspace, dev, int(load().hrc_search_workspace_bytes(n_docs, int(tokens.shape[0]), nq,
                                                                                        lq, k, path)), st)
    keys = torch.empty((nq, k), dtype=torch.int64, device=dev)
    ids = torch.empty((nq, k), dtype=torch.int32, device=dev) if unpack else None
    scores = torch.empty((nq, k), dtype=torch.float32, device=dev) if unpack else None
    with _on(dev):
        rc = load().hrc_search(_ptr(tokens), _ptr(offsets), n_docs, int(tokens.shape[0]), _ptr(queries), nq, lq, k,
                               id_base, ws_ptr, ws_bytes, _ptr(keys), _ptr(ids), _ptr(scores), path, st)
    _check(rc, "hrc_search")
    return keys, ids, scores


class HostSearch:
    """End-to-end search with host buffers (hrc_search_host): pinned fp32 queries in, pinned ids / scores out,
    one C call and one stream synchronisation per search.  Device scratch and pinned staging are kept between calls."""

    def __init__(self):
        self.key = None
        self._lock = threading.Lock()      # the staging buffers are this object's: one search at a time per object

    def __call__(self, *args, **kwargs):
        with self._lock:
            return self._call(*args, **kwargs)

    def _ensure(self, dev, nq: int, lq: int, n_docs: int, total_tokens: int, k: int, path: int):
        key = (str(dev), nq, lq, n_docs, total_tokens, k, path)
        if self.key != key:
            need = int(load().hrc_search_host_workspace_bytes(n_docs, total_tokens, nq, lq, k, path))
            self.ws = torch.empty(max(need, 256), dtype=torch.uint8, device=dev)
            self.ws_bytes = need
            self.q_pinned = torch.empty((nq, lq, DIM), dtype=torch.float32).pin_memory()
            self.out = torch.empty((2, nq, k), dtype=torch.int32).pin_memory()     # ids | scores: adjacent -> ONE D2H copy
            self.ids, self.scores = self.out[0], self.out[1].view(torch.float32)
            self.key = key

    def _call(self, tokens: torch.Tensor, offsets: torch.Tensor, queries_host: torch.Tensor, k: int, *,
              id_base: int = 0, path: int = PATH_AUTO, copy: bool = True):
        """queries_host: fp32 CPU tensor [nq, lq, 128] (pinned: used in place; pageable: staged through a pinned
        buffer).  Returns (ids int32 [nq, k], scores fp32 [nq, k]) CPU tensors.  With copy=True (default) they are
        fresh tensors the caller owns; copy=False returns this object's PINNED staging buffers, which the next call
        overwrites."""
        dev = _require_cuda(tokens, offsets)
        _check_store(tokens, offsets)
        assert not queries_host.is_cuda and queries_host.dtype == torch.float32 and queries_host.dim() == 3
        assert queries_host.shape[2] == DIM and queries_host.is_contiguous()
        n_docs = offsets.numel() - 1
        nq, lq = int(queries_host.shape[0]), int(queries_host.shape[1])
        self._ensure(dev, nq, lq, n_docs, int(tokens.shape[0]), k, path)
        src = queries_host
        if not src.is_pinned():
            self.q_pinned.copy_(src)
            src = self.q_pinned
        with _on(dev):
            stream = torch.cuda.current_stream(dev)
            rc = load().hrc_search_host(_ptr(tokens), _ptr(offsets), n_docs, int(tokens.shape[0]), src.data_ptr(), nq,
                                        lq, k, id_base, _ptr(self.ws), self.ws_bytes, self.ids.data_ptr(),
                                        self.scores.data_ptr(), path, stream.cuda_stream)
            _check(rc, "hrc_search_host")
            stream.synchronize()
        if copy:
            return self.ids.clone(), self.scores.clone()
        return self.ids, self.scores


def hybrid_retrieve(tokens: torch.Tensor, offsets: torch.Tensor, queries: torch.Tensor, bm25_ids: torch.Tensor, *,
                    colbert_k: int, rrf_k: int, n_candidates: int, final_k: int, id_base: int = 0,
                    path: int = PATH_AUTO, workspace: Optional[Workspace] = None):
    """ColBERT top-k -> RRF with the BM25 lists -> rerank of the stored candidates, one C call (hrc_hybrid_retrieve).
    Returns (global doc ids int32 [nq, final_k], MaxSim scores fp32 [nq, final_k])."""
    dev = _require_cuda(tokens, offsets, queries, bm25_ids)
    _check_store(tokens, offsets)
    _check_queries(queries)
    assert bm25_ids.dtype == torch.int32 and bm25_ids.dim() == 2 and bm25_ids.shape[0] == queries.shape[0]
    n_docs = offsets.numel() - 1
    nq, lq = int(queries.shape[0]), int(queries.shape[1])
    need = int(load().hrc_hybrid_retrieve_workspace_bytes(n_docs, int(tokens.shape[0]), nq, lq, colbert_k, n_candidates,
                                                          final_k, path))
    st = _stream(dev)
    ws_ptr, ws_bytes, _keep = _ws(workspace, dev, need, st)
    ids = torch.empty((nq, final_k), dtype=torch.int32, device=dev)
    scores = torch.empty((nq, final_k), dtype=torch.float32, device=dev)
    with _on(dev):
        rc = load().hrc_hybrid_retrieve(_ptr(tokens), _ptr(offsets), n_docs, int(tokens.shape[0]), _ptr(queries), nq, lq,
                                        _ptr(bm25_ids), int(bm25_ids.shape[1]), colbert_k, rrf_k, n_candidates, final_k,
                                        id_base, ws_ptr, ws_bytes, _ptr(ids), _ptr(scores), path, st)
    _check(rc, "hrc_hybrid_retrieve")
    return ids, scores


_RERANK_WS: dict = {}      # (n_cand, n_queries, lq, k) -> workspace bytes: the rerank is latency-bound, every ctypes call counts


def rerank(tokens: torch.Tensor, offsets: torch.Tensor, cand_ids: torch.Tensor, queries: torch.Tensor, k: int, *,
           path: int = PATH_AUTO, workspace: Optional[Workspace] = None, want_cand_scores: bool = False):
    """Fused candidate MaxSim + sorted top-k in one C call.
    Returns (pos int32 [nq,k], doc ids int32 [nq,k], scores fp32 [nq,k], candidate scores fp32 [nq,n_cand] | None)."""
    dev = _require_cuda(tokens, offsets, cand_ids, queries)
    assert cand_ids.dtype == torch.int32 and cand_ids.dim() == 2 and cand_ids.shape[0] == queries.shape[0]
    _check_store(tokens, offsets)
    _check_queries(queries)
    n_docs = offsets.numel() - 1
    nq, lq = int(queries.shape[0]), int(queries.shape[1])
    n_cand = int(cand_ids.shape[1])
    shape = (n_cand, nq, lq, k)
    need = _RERANK_WS.get(shape)
    if need is None:
        need = _RERANK_WS[shape] = int(load().hrc_rerank_workspace_bytes(n_cand, nq, lq, k))
    st = _stream(dev)
    ws_ptr, ws_bytes, _keep = _ws(workspace, dev, need, st)
    cand_scores = torch.empty((nq, n_cand), dtype=torch.float32, device=dev) if want_cand_scores else None
    out = torch.empty((3, nq, k), dtype=torch.int32, device=dev)      # one allocation: pos | ids | scores (fp32 view)
    pos, ids, scores = out[0], out[1], out[2].view(torch.float32)
    with _on(dev):
        rc = load().hrc_rerank(_ptr(tokens), _ptr(offsets), n_docs, int(tokens.shape[0]), _ptr(cand_ids), n_cand,
                               _ptr(queries), nq, lq, k, ws_ptr, ws_bytes, pos.data_ptr(), ids.data_ptr(),
                               scores.data_ptr(), _ptr(cand_scores), path, st)
    _check(rc, "hrc_rerank")
    return pos, ids, scores, cand_scores


def topk_workspace_bytes(n: int, n_rows: int, k: int) -> int:
    return int(load().hrc_topk_workspace_bytes(n, n_rows, k))


def topk(scores: torch.Tensor, k: int, *, ids: Optional[torch.Tensor] = None, id_base: int = 0,
         workspace: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Per-row top-k keys (int64 view of the uint64 keys), sorted best first.  [n_rows, k]."""
    dev = _require_cuda(scores, ids)
    assert scores.dtype == torch.float32 and scores.dim() == 2
    n_rows, n = int(scores.shape[0]), int(scores.shape[1])
    if ids is not None:
        assert ids.dtype == torch.int32 and ids.shape == scores.shape
    need = topk_workspace_bytes(n, n_rows, k)
    if need and (workspace is None or workspace.numel() * workspace.element_size() < need):
        workspace = torch.empty(need, dtype=torch.uint8, device=dev)
    keys = torch.empty((n_rows, k), dtype=torch.int64, device=dev)
    with _on(dev):
        rc = load().hrc_topk(_ptr(scores), _ptr(ids), n, n_rows, k, id_base, _ptr(keys), _ptr(workspace),
                             0 if workspace is None else workspace.numel() * workspace.element_size(),
                             _stream(dev))
    _check(rc, "hrc_topk")
    return keys


def topk_merge(keys_in: torch.Tensor, k: int) -> torch.Tensor:
    """Per-row top-k of candidate keys [n_rows, n_in] -> sorted keys [n_rows, k]."""
    dev = _require_cuda(keys_in)
    assert keys_in.dtype == torch.int64 and keys_in.dim() == 2
    n_rows, n_in = int(keys_in.shape[0]), int(keys_in.shape[1])
    out = torch.empty((n_rows, k), dtype=torch.int64, device=dev)
    with _on(dev):
        rc = load().hrc_topk_merge(_ptr(keys_in), n_in, n_rows, k, _ptr(out), _stream(dev))
    _check(rc, "hrc_topk_merge")
    return out


def keys_unpack(keys: torch.Tensor):
    """keys (int64 view) -> (ids int32, scores fp32), same shape; empty slots give (-1, -inf)."""
    dev = _require_cuda(keys)
    assert keys.dtype == torch.int64
    ids = torch.empty(keys.shape, dtype=torch.int32, device=dev)
    scores = torch.empty(keys.shape, dtype=torch.float32, device=dev)
    with _on(dev):
        rc = load().hrc_keys_unpack(_ptr(keys), keys.numel(), _ptr(ids), _ptr(scores), _stream(dev))
    _check(rc, "hrc_keys_unpack")
    return ids, scores


def rrf_fuse(ids_a: torch.Tensor, ids_b: torch.Tensor, rrf_k: int, top_n: int):
    """Row-wise RRF of two ranked id lists (int32 [rows, n]); returns (ids int32, scores fp64, counts int32)."""
    dev = _require_cuda(ids_a, ids_b)
    assert ids_a.dtype == torch.int32 and ids_b.dtype == torch.int32 and ids_a.dim() == 2 and ids_b.dim() == 2
    assert ids_a.shape[0] == ids_b.shape[0]
    rows = int(ids_a.shape[0])
    ids = torch.empty((rows, top_n), dtype=torch.int32, device=dev)
    scores = torch.empty((rows, top_n), dtype=torch.float64, device=dev)
    counts = torch.empty((rows,), dtype=torch.int32, device=dev)
    with _on(dev):
        rc = load().hrc_rrf_fuse(_ptr(ids_a), int(ids_a.shape[1]), _ptr(ids_b), int(ids_b.shape[1]), rows, rrf_k,
                                 top_n, _ptr(ids), _ptr(scores), _ptr(counts), _stream(dev))
    _check(rc, "hrc_rrf_fuse")
    return ids, scores, counts


def synth_tokens(out: torch.Tensor, token_begin: int, seed: int) -> torch.Tensor:
    """Fill `out` (bf16 [n, 128]) with the deterministic synthetic rows token_begin .. token_begin+n."""
    dev = _require_cuda(out)
    assert out.dtype == torch.bfloat16 and out.dim() == 2 and out.shape[1] == DIM
    with _on(dev):
        rc = load().hrc_synth_tokens(_ptr(out), token_begin, int(out.shape[0]), seed & (2**64 - 1), _stream(dev))
    _check(rc, "hrc_synth_tokens")
    return out


def read_probe(buf: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Stream `buf` once with 16-byte loads (bench utility: the pure-read bandwidth probe)."""
    dev = _require_cuda(buf)
    if out is None:
        out = torch.zeros(1, dtype=torch.int32, device=dev)
    with _on(dev):
        rc = load().hrc_read_probe(_ptr(buf), buf.numel() * buf.element_size(), _ptr(out), _stream(dev))
    _check(rc, "hrc_read_probe")
    return out


# ------------------------------------------------------------------------------------------------------------------
# multi-GPU exchange inside libhrc (one process per GPU)
# ------------------------------------------------------------------------------------------------------------------
class Comm:
    """A libhrc communicator over the ranks of a torch.distributed group: NCCL all-gather or direct peer stores over
    NVLink (transport P2P), both inside libhrc.so.  torch.distributed is only used ONCE, to hand rank 0's 128-byte id
    to the other ranks."""

    def __init__(self, device: torch.device, group=None, p2p_max_keys: int = 0):
        import torch.distributed as dist
        self.device = device
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        lib = load()
        idbuf = (ctypes.c_uint8 * COMM_ID_BYTES)()
        if self.rank == 0:
            _check(lib.hrc_comm_unique_id(idbuf), "hrc_comm_unique_id")
        payload = [bytes(idbuf)]
        dist.broadcast_object_list(payload, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        raw = (ctypes.c_uint8 * COMM_ID_BYTES).from_buffer_copy(payload[0])
        handle = ctypes.c_void_p()
        with torch.cuda.device(device):
            _check(lib.hrc_comm_init(raw, self.world, self.rank, ctypes.byref(handle)), "hrc_comm_init")
        self.handle = handle
        self.p2p_max_keys = 0
        if p2p_max_keys:
            self.enable_p2p(p2p_max_keys)

    def enable_p2p(self, max_keys: int) -> None:
        with torch.cuda.device(self.device):
            _check(load().hrc_comm_enable_p2p(self.handle, int(max_keys), _stream(self.device)), "hrc_comm_enable_p2p")
        self.p2p_max_keys = max(self.p2p_max_keys, int(max_keys))

    def close(self) -> None:
        if getattr(self, "handle", None):
            load().hrc_comm_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


def allgather_merge_topk(comm: Comm, local_keys: torch.Tensor, k: int, *, transport: int = TRANSPORT_NCCL,
                         workspace: Optional[Workspace] = None, unpack: bool = False):
    """local_keys int64 [n_rows, k] (0 = empty) -> merged keys int64 [n_rows, k] (+ ids, scores with unpack=True)."""
    dev = _require_cuda(local_keys)
    assert local_keys.dtype == torch.int64 and local_keys.dim() == 2 and local_keys.shape[1] == k
    n_rows = int(local_keys.shape[0])
    st = _stream(dev)
    ws_ptr, ws_bytes, _keep = _ws(workspace, dev, int(load().hrc_allgather_merge_workspace_bytes(comm.world, n_rows, k)), st)
    keys = torch.empty((n_rows, k), dtype=torch.int64, device=dev)
    ids = torch.empty((n_rows, k), dtype=torch.int32, device=dev) if unpack else None
    scores = torch.empty((n_rows, k), dtype=torch.float32, device=dev) if unpack else None
    with _on(dev):
        rc = load().hrc_allgather_merge_topk(comm.handle, _ptr(local_keys), n_rows, k, transport, ws_ptr, ws_bytes,
                                             _ptr(keys), _ptr(ids), _ptr(scores), st)
    _check(rc, "hrc_allgather_merge_topk")
    return keys, ids, scores


def sharded_search(comm: Comm, tokens: torch.Tensor, offsets: torch.Tensor, queries: torch.Tensor, k: int, *,
                   id_base: int = 0, path: int = PATH_AUTO, transport: int = TRANSPORT_NCCL,
                   workspace: Optional[Workspace] = None, unpack: bool = True):
    """Local search over this rank's shard + exchange + merge in ONE C call.  Returns (keys, ids | None, scores | None)."""
    dev = _require_cuda(tokens, offsets, queries)
    _check_store(tokens, offsets)
    _check_queries(queries)
    n_docs = offsets.numel() - 1
    nq, lq = int(queries.shape[0]), int(queries.shape[1])
    need = int(load().hrc_sharded_search_workspace_bytes(comm.world, n_docs, int(tokens.shape[0]), nq, lq, k, path))
    st = _stream(dev)
    ws_ptr, ws_bytes, _keep = _ws(workspace, dev, need, st)
    keys = torch.empty((nq, k), dtype=torch.int64, device=dev)
    ids = torch.empty((nq, k), dtype=torch.int32, device=dev) if unpack else None
    scores = torch.empty((nq, k), dtype=torch.float32, device=dev) if unpack else None
    with _on(dev):
        rc = load().hrc_sharded_search(comm.handle, transport, _ptr(tokens), _ptr(offsets), n_docs, int(tokens.shape[0]),
                                       _ptr(queries), nq, lq, k, id_base, ws_ptr, ws_bytes, _ptr(keys), _ptr(ids),
                                       _ptr(scores), path, st)
    _check(rc, "hrc_sharded_search")
    return keys, ids, scores


class ShardedHostSearch:
    """hrc_sharded_search_host: pinned fp32 queries in, pinned ids / scores out, one C call and one synchronisation."""

    def __init__(self, comm: Comm):
        self.comm = comm
        self.key = None
        self._lock = threading.Lock()

    def __call__(self, *args, **kwargs):
        with self._lock:
            return self._call(*args, **kwargs)

    def _call(self, tokens, offsets, queries_host, k, *, id_base=0, path=PATH_AUTO, transport=TRANSPORT_NCCL, copy=True):
        dev = _require_cuda(tokens, offsets)
        _check_store(tokens, offsets)
        assert not queries_host.is_cuda and queries_host.dtype == torch.float32 and queries_host.dim() == 3
        n_docs, total = offsets.numel() - 1, int(tokens.shape[0])
        nq, lq = int(queries_host.shape[0]), int(queries_host.shape[1])
        key = (str(dev), nq, lq, n_docs, total, k, path)
        if self.key != key:
            need = int(load().hrc_sharded_search_host_workspace_bytes(self.comm.world, n_docs, total, nq, lq, k, path))
            self.ws = torch.empty(max(need, 256), dtype=torch.uint8, device=dev)
            self.ws_bytes = need
            self.q_pinned = torch.empty((nq, lq, DIM), dtype=torch.float32).pin_memory()
            self.out = torch.empty((2, nq, k), dtype=torch.int32).pin_memory()     # ids | scores: adjacent -> ONE D2H copy
            self.ids, self.scores = self.out[0], self.out[1].view(torch.float32)
            self.key = key
        src = queries_host.contiguous()
        if not src.is_pinned():
            self.q_pinned.copy_(src)
            src = self.q_pinned
        with _on(dev):
            stream = torch.cuda.current_stream(dev)
            rc = load().hrc_sharded_search_host(self.comm.handle, transport, _ptr(tokens), _ptr(offsets), n_docs, total,
                                                src.data_ptr(), nq, lq, k, id_base, _ptr(self.ws), self.ws_bytes,
                                                self.ids.data_ptr(), self.scores.data_ptr(), path, stream.cuda_stream)
            _check(rc, "hrc_sharded_search_host")
            stream.synchronize()
        return (self.ids.clone(), self.scores.clone()) if copy else (self.ids, self.scores)


def sharded_hybrid_retrieve(comm: Comm, tokens: torch.Tensor, offsets: torch.Tensor, n_docs_global: int,
                            queries: torch.Tensor, bm25_ids: torch.Tensor, *, colbert_k: int, rrf_k: int,
                            n_candidates: int, final_k: int, id_base: int = 0, path: int = PATH_AUTO,
                            transport: int = TRANSPORT_NCCL, workspace: Optional[Workspace] = None):
    """The document-sharded hybrid pipeline, one C call per rank; every rank gets the same (ids, scores)."""
    dev = _require_cuda(tokens, offsets, queries, bm25_ids)
    _check_store(tokens, offsets)
    _check_queries(queries)
    assert bm25_ids.dtype == torch.int32 and bm25_ids.dim() == 2 and bm25_ids.shape[0] == queries.shape[0]
    n_docs, total = offsets.numel() - 1, int(tokens.shape[0])
    nq, lq = int(queries.shape[0]), int(queries.shape[1])
    need = int(load().hrc_sharded_hybrid_workspace_bytes(comm.world, n_docs, total, nq, lq, colbert_k, n_candidates, final_k, path))
    st = _stream(dev)
    ws_ptr, ws_bytes, _keep = _ws(workspace, dev, need, st)
    ids = torch.empty((nq, final_k), dtype=torch.int32, device=dev)
    scores = torch.empty((nq, final_k), dtype=torch.float32, device=dev)
    with _on(dev):
        rc = load().hrc_sharded_hybrid_retrieve(comm.handle, transport, _ptr(tokens), _ptr(offsets), n_docs, total,
                                                int(n_docs_global), _ptr(queries), nq, lq, _ptr(bm25_ids),
                                                int(bm25_ids.shape[1]), colbert_k, rrf_k, n_candidates, final_k, id_base,
                                                ws_ptr, ws_bytes, _ptr(ids), _ptr(scores), path, st)
    _check(rc, "hrc_sharded_hybrid_retrieve")
    return ids, scores


def store_read_file(path: str, file_offset: int, dst: torch.Tensor, chunk_bytes: int = 0) -> float:
    """Stream dst.nbytes bytes of `path` from `file_offset` straight into the CUDA tensor `dst` (pinned double buffer,
    one cudaMemcpyAsync per chunk).  Returns the elapsed seconds."""
    dev = _require_cuda(dst)
    secs = ctypes.c_double(0.0)
    with _on(dev):
        rc = load().hrc_store_read_file(os.fsencode(path), int(file_offset), dst.numel() * dst.element_size(), _ptr(dst),
                                        int(chunk_bytes), _stream(dev), ctypes.byref(secs))
    _check(rc, "hrc_store_read_file")
    return float(secs.value)


def store_write_file(path: str, file_offset: int, src: torch.Tensor, chunk_bytes: int = 0) -> float:
    """Stream the CUDA tensor `src` into `path` at `file_offset` (pinned double buffer).  Returns the elapsed seconds."""
    dev = _require_cuda(src)
    secs = ctypes.c_double(0.0)
    with _on(dev):
        rc = load().hrc_store_write_file(os.fsencode(path), int(file_offset), src.numel() * src.element_size(), _ptr(src),
                                         int(chunk_bytes), _stream(dev), ctypes.byref(secs))
    _check(rc, "hrc_store_write_file")
    return float(secs.value)
