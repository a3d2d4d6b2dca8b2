#!/usr/bin/env python
"""bench.py — MaxSim docs scored/sec (top-k search), the metric BASELINE.json names.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one full-corpus MaxSim top-100 search of one 32-token query (config C2: 1M passages x 128
tokens per GPU, 32.8 GB bf16, far larger than the 126 MB L2, so every step streams from HBM).  With N>1
each rank holds its own 1M-document shard of an N-million-document corpus (weak scaling, SURVEY.md
§8(e)); a step adds the all-gather of k keys per rank and the on-device merge.
Prints ONE JSON line (rank 0).  `--impl reference` times the CPU oracle port of the same path.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "maxsim_docs_scored_per_sec"
UNIT = "docs/s"
K = 100
LQ = 32
DOC_LEN = 128
SEED = 20260102


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--docs-per-gpu", type=int, default=1_000_000)
    ap.add_argument("--cpu-sample-docs", type=int, default=20_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_config(args, n_gpus):
    return {
        "workload": f"C2 full-corpus MaxSim top-{K}: {args.docs_per_gpu} passages x {DOC_LEN} tokens per GPU, "
                    f"1 query x {LQ} tokens, dim 128, bf16 packed store",
        "docs_per_gpu": args.docs_per_gpu, "doc_tokens": DOC_LEN, "query_tokens": LQ, "k": K,
        "global_docs": args.docs_per_gpu * n_gpus, "parallelism": f"doc-shard x{n_gpus}",
        "l2_note": "inputs (32.8 GB/GPU) are larger than L2 (126 MB); no explicit flush needed",
    }


# --------------------------------------------------------------------------------------------------
# CPU oracle timing (cpu_baseline leg and --impl reference)
# --------------------------------------------------------------------------------------------------
def cpu_sample(n_docs):
    """A bounded sample of the C2 workload on the host: n_docs x 128 L2-normalised tokens, fp32."""
    import torch
    g = torch.Generator().manual_seed(SEED)
    tok = torch.nn.functional.normalize(torch.randn((n_docs * DOC_LEN, 128), generator=g), dim=-1)
    q = torch.nn.functional.normalize(torch.randn((1, LQ, 128), generator=g), dim=-1)
    off = torch.arange(0, n_docs * DOC_LEN + 1, DOC_LEN)
    return q, tok, off


def cpu_step(q, tok, off):
    """The reference's CPU path for this workload: torch MaxSim (einsum -> max -> sum) + torch.topk (:764-767)."""
    import torch
    from oracle import maxsim_oracle as o
    dense = tok.view(-1, DOC_LEN, 128)                 # C2 documents are uniform: the reference's dense layout
    parts = [o.maxsim_dense(q[0], dense[i:i + 4096]).reshape(-1) for i in range(0, dense.shape[0], 4096)]
    scores = torch.cat(parts)
    return torch.topk(scores, k=min(K, scores.numel()))


def time_cpu(n_docs, min_seconds=10.0, max_seconds=30.0):
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    q, tok, off = cpu_sample(n_docs)
    cpu_step(q, tok, off)  # warm-up
    t0 = time.perf_counter()
    passes = 0
    while True:
        cpu_step(q, tok, off)
        passes += 1
        el = time.perf_counter() - t0
        if el >= min_seconds or (passes >= 2 and el >= max_seconds):
            break
    # the reference's function EXACTLY as coded (:821-829, mean-pool cosine — not MaxSim) on the same sample, for the record
    from oracle import maxsim_oracle as o
    dense = tok.view(-1, DOC_LEN, 128)
    o.literal_reference(q[0], dense)
    t1 = time.perf_counter()
    lit_passes = 0
    while time.perf_counter() - t1 < 2.0:
        torch.topk(o.literal_reference(q[0], dense), k=min(K, n_docs))
        lit_passes += 1
    lit_el = time.perf_counter() - t1
    return {"value": n_docs * passes / el, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{passes} passes over {n_docs} docs x {DOC_LEN} tokens (fp32, torch CPU oracle: "
                      f"einsum -> max -> sum -> torch.topk), {el:.1f} s",
            "tokens_per_s": n_docs * passes * DOC_LEN / el,
            "reference_literal_docs_per_s": n_docs * lit_passes / lit_el,
            "reference_literal_note": "local_rag_complete.py:821-829 as coded (mean-pool cosine, not MaxSim) + torch.topk, same sample"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    n = args.cpu_sample_docs
    q, tok, off = cpu_sample(n)
    for _ in range(max(args.warmup, 1)):
        cpu_step(q, tok, off)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_step(q, tok, off)
    el = time.perf_counter() - t0
    value = n * args.steps / el
    cfg = workload_config(args, args.gpus)
    cfg["reference_note"] = ("the reference is a CPU/MPS Python script; its MaxSim path is timed as the torch CPU oracle "
                             "port on a bounded sample (each step = one pass over the sample)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{args.steps} steps x {n} docs x {DOC_LEN} tokens"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# clocks sampler
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock, power and throttle reasons sampled every ~5 ms with NVML while the timed region runs (a thread in
    this process); falls back to `nvidia-smi -lms 20` when the NVML binding is unavailable."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.samples = []          # (t, sm_mhz, max_mhz, power_w, [reasons])
        self.index = index
        self.proc = None
        self._stop = threading.Event()
        self._thread = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            bits = {"hw_slowdown": pynvml.nvmlClocksThrottleReasonHwSlowdown,
                    "hw_thermal_slowdown": pynvml.nvmlClocksThrottleReasonHwThermalSlowdown,
                    "sw_thermal_slowdown": pynvml.nvmlClocksThrottleReasonSwThermalSlowdown,
                    "sw_power_cap": pynvml.nvmlClocksThrottleReasonSwPowerCap}

            def loop():
                while not self._stop.is_set():
                    try:
                        sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                        pw = pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0
                        r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                        self.samples.append((time.perf_counter(), float(sm), float(mx), pw, [n for n, b in bits.items() if r & b]))
                    except Exception:  # noqa: BLE001
                        pass
                    time.sleep(0.004)

            self._thread = threading.Thread(target=loop, daemon=True)
            self._thread.start()
            self.source = "nvml, 4 ms period"
            return
        except Exception:  # noqa: BLE001
            self._thread = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
            self.source = "nvidia-smi -lms 20"
        except OSError:
            self.proc = None

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[self.index])
            except (ValueError, IndexError):
                pass
        return self.index

    def _read(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            try:
                self.samples.append((time.perf_counter(), float(f[0]), float(f[1]), float(f[2]),
                                     [n for n, v in zip(self.NAMES, f[3:7]) if v.lower().startswith("active")]))
            except (ValueError, IndexError):
                continue

    def stop(self, t0, t1):
        self._stop.set()
        if self._thread is None and self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"]}
        if self.proc is not None:
            time.sleep(0.05)
            self.proc.terminate()
        rows = [s for s in self.samples if t0 <= s[0] <= t1] or self.samples[-3:]
        reasons = sorted({r for s in rows for r in s[4]})
        return {"sm_mhz": statistics.median([s[1] for s in rows]) if rows else None,
                "sm_max_mhz": max([s[2] for s in rows]) if rows else None,
                "power_w": statistics.median([s[3] for s in rows]) if rows else None,
                "reasons": reasons, "samples": len(rows), "source": getattr(self, "source", "")}


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic():
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f)
    return None


def run_ours(args):
    import torch
    import torch.distributed as dist
    import hybrid_rag_colbertv2_b200 as hrc
    from hybrid_rag_colbertv2_b200 import _lib
    from hybrid_rag_colbertv2_b200.synth import synth_queries, synth_store

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    n_global = args.docs_per_gpu * world
    store = synth_store(n_global, DOC_LEN, DOC_LEN, seed=SEED, device=dev, rank=rank, world_size=world)
    retr = hrc.JinaColBERTRetriever(hrc.RAGConfig(device=str(dev)))
    retr.store = store
    searcher = hrc.ShardedSearcher(retr) if world > 1 else None
    n_q = 8
    queries = synth_queries(n_q, LQ, device=dev)                    # rotate queries; the corpus is what streams
    q_host = [synth_queries(n_q, LQ)[i:i + 1].float().pin_memory() for i in range(n_q)]   # what an encoder hands over

    def step(i):
        q = queries[i % n_q:i % n_q + 1]
        return searcher.search_keys(q, K) if searcher else retr.search_keys(q, K)

    def step_e2e(i):
        qh = q_host[i % n_q]
        if searcher is None:                      # one C call: H2D, bf16, MaxSim, top-k, unpack, D2H; one sync
            return retr.search_host(qh, K)
        return searcher.search_host(qh, K)        # H2D, local search, all-gather, merge, unpack, D2H; one sync

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    # ---- device-resident timing -------------------------------------------------------------------
    for i in range(args.warmup):
        step(i)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    barrier()
    t_wall1 = time.perf_counter()
    gpu_launches = _lib.launch_count() - launches0
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None

    # ---- dominant kernel alone (same stream, CUDA events) for the roofline --------------------------
    scores_buf = torch.empty((1, store.n_docs), dtype=torch.float32, device=dev)
    for i in range(2):
        _lib.maxsim_scores(store.tokens, store.offsets, queries[0:1], out=scores_buf)
    barrier()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    for i in range(args.steps):
        _lib.maxsim_scores(store.tokens, store.offsets, queries[i % n_q:i % n_q + 1], out=scores_buf)
    k1.record()
    barrier()
    kernel_ms = max_over_ranks(k0.elapsed_time(k1)) / args.steps
    # the rest of a step (radix top-k of the score row) alone, for the kernel's share of the step
    ws = torch.empty(max(_lib.topk_workspace_bytes(store.n_docs, 1, K), 1), dtype=torch.uint8, device=dev)
    for i in range(2):
        _lib.topk(scores_buf, K, workspace=ws)
    barrier()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for i in range(args.steps):
        _lib.topk(scores_buf, K, workspace=ws)
    s1.record()
    barrier()
    topk_ms = max_over_ranks(s0.elapsed_time(s1)) / args.steps

    # ---- end to end through the public API with host buffers ---------------------------------------
    for i in range(args.warmup):
        step_e2e(i)
    barrier()
    x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    x0.record()
    for i in range(args.steps):
        ids_h, scores_h = step_e2e(i)
    x1.record()
    barrier()
    e2e_ms_total = max_over_ranks(x0.elapsed_time(x1))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    docs_per_step = n_global
    ms_per_step = ms_total / args.steps
    value = docs_per_step / (ms_per_step * 1e-3)
    peak, peak_src = load_peaks()
    algo_bytes = 256.0 * store.total_tokens                        # 256 B per document token
    achieved = algo_bytes / (kernel_ms * 1e-3) / 1e9
    traffic = load_traffic()
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic", "config": workload_config(args, world),
        "tokens_per_s": value * DOC_LEN,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic["bytes_per_launch"] if traffic else None,
                     "kernel": "maxsim_tc_kernel<1>", "kernel_ms": kernel_ms, "algorithmic_bytes_per_launch": algo_bytes,
                     "peak_source": peak_src, "frac_of_nominal": {"7.7TB/s_hgx": achieved / 7700.0, "8.0TB/s_dgx": achieved / 8000.0},
                     "topk_ms": topk_ms, "kernel_share_of_step": kernel_ms / (kernel_ms + topk_ms)},
        "e2e": {"value": docs_per_step / (e2e_ms_total / args.steps * 1e-3), "unit": UNIT,
                "h2d_bytes_per_step": int(q_host[0].numel() * 4), "d2h_bytes_per_step": K * 8,
                "ms_per_step": e2e_ms_total / args.steps,
                "api": ("JinaColBERTRetriever.search_host: pinned fp32 query -> hrc_search_host (H2D, bf16, MaxSim, top-k, "
                        "unpack, D2H) -> ids/scores on the host" if world == 1 else
                        "ShardedSearcher.search_host: pinned fp32 query -> H2D -> local search -> NCCL all-gather -> merge -> "
                        "unpack -> D2H (pinned) -> ids/scores on the host")},
        "gpu_launches": int(gpu_launches),
        "clocks": clocks,
    }
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = time_cpu(args.cpu_sample_docs)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
