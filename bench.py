#!/usr/bin/env python
"""bench.py — MaxSim docs scored/sec (top-k search), the metric BASELINE.json names.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline: a step = one full-corpus MaxSim top-100 search of one 32-token query (config C2: 1M passages x 128
tokens per GPU, 32.8 GB bf16, far larger than the 126 MB L2, so every step streams from HBM).  With N>1 each
rank holds its own 1M-document shard of an N-million-document corpus (weak scaling); a step
adds the exchange of k keys per rank and the merge (over peer memory inside the search's own final kernel when every
rank can map its peers, else ncclAllGather + a merge kernel; --transport).  The dominant kernel is timed INSIDE the timed
steps (hrc_trace_*: CUDA events around its launches on its own stream), so kernel_ms <= ms_per_step by
construction.

After the headline the same process measures, outside the headline's timed region, a `secondary` object with the
other BASELINE.json configs: `sustained` (C2 back to back for >= 2 s, clocks recorded), `read_peak` (pure-read
bandwidth probe over the same corpus), `c4` (hybrid pipeline, 1k queries, sequential and in batches of 64), `c3`
(256 queries x 1M ragged passages: tensor roofline), `c1` (rerank of 50 candidates: latency), `ragged` (single
query over the C3 corpus) and, with N>1, `c5` (10M passages x 128 tokens split over the ranks), `c4_sharded` (the
hybrid pipeline over that corpus), `pipelined`, a `parity_check` of the merged top-k over the real interconnect and
the `breakdown` of a sharded step.

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
SEED_C3 = 20260103


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--docs-per-gpu", type=int, default=1_000_000)
    ap.add_argument("--cpu-sample-docs", type=int, default=20_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="headline only (profiling runs)")
    ap.add_argument("--secondary", default="sustained,read_peak,c4,c3,c1,ragged,c5,pipelined",
                    help="comma-separated subset of the secondary measurements")
    ap.add_argument("--c5-global-docs", type=int, default=10_000_000)
    ap.add_argument("--transport", default="auto", choices=["auto", "nccl", "p2p", "torch"],
                    help="exchange step of the sharded search (N > 1): peer stores + flags over NVLink inside libhrc when "
                         "every rank can map its peers, else ncclAllGather inside libhrc (auto, default); or force one; "
                         "or torch.distributed")
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

    def __init__(self, index, period_s=0.004):
        self.samples = []          # (t, sm_mhz, max_mhz, power_w, [reasons])
        self.index = index
        self.period_s = period_s   # 4 ms for the 0.1 s headline region; the multi-second secondary runs poll every 25 ms
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
                    time.sleep(self.period_s)

            self._thread = threading.Thread(target=loop, daemon=True)
            self._thread.start()
            self.source = f"nvml, {self.period_s * 1e3:.0f} ms period"
            return self
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
        return self

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
# helpers
# --------------------------------------------------------------------------------------------------
def load_peaks():
    """(hbm GB/s, bf16 TFLOP/s sustained, bf16 TFLOP/s burst, source)"""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return (float(d["hbm_gbs"]), float(d.get("bf16_tflops_sustained", 1390.6)), float(d.get("bf16_tflops", 1665.1)),
                "measured (MEASURED_PEAKS.json: hbm_gbs = burst read+write copy; bf16_tflops_sustained = 4 s of cuBLAS)")
    return 6650.0, 1400.0, 1590.0, "fallback (B200_PROFILING.md)"


def load_profile_json(name):
    path = os.path.join(ROOT, "profiles", name)
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f)
    return None


def pct(xs, p):
    xs = sorted(xs)
    if not xs:
        return None
    return xs[min(len(xs) - 1, max(0, int(round(p / 100.0 * (len(xs) - 1)))))]


def summary(xs):
    return {"median": statistics.median(xs), "p10": pct(xs, 10), "p90": pct(xs, 90), "min": min(xs), "max": max(xs),
            "n": len(xs)} if xs else None


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
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
    hbm_peak, tf_sus, tf_burst, peak_src = load_peaks()
    want = set() if args.no_secondary else set(args.secondary.split(","))

    n_global = args.docs_per_gpu * world
    store = synth_store(n_global, DOC_LEN, DOC_LEN, seed=SEED, device=dev, rank=rank, world_size=world)
    retr = hrc.JinaColBERTRetriever(hrc.RAGConfig(device=str(dev)))
    retr.store = store
    searcher = hrc.ShardedSearcher(retr, transport=args.transport) if world > 1 else None
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

    def timed_loop(fn, steps, warmup, sample_clocks=False, clock_period_s=0.004):
        """W untimed warm-ups, then `steps` calls bracketed by barrier + synchronize; CUDA events on the launching
        stream, one per step boundary; the scoring kernels inside are traced.  -> dict (ms are max over ranks)."""
        for i in range(warmup):
            fn(i)
        # everything that only one rank does (NVML init of the clock sampler: ~10 ms) happens BEFORE the barrier: a rank
        # that enters the timed region late makes every other rank wait for it in its first all-gather
        sampler = ClockSampler(local_rank, clock_period_s).start() if (sample_clocks and rank == 0) else None
        _lib.trace_enable(8 * steps + 8)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        barrier()
        launches0 = _lib.launch_count()
        t0 = time.perf_counter()
        ev[0].record()
        for i in range(steps):
            fn(i)
            ev[i + 1].record()
        barrier()
        t1 = time.perf_counter()
        launches = _lib.launch_count() - launches0
        kern = _lib.trace_collect()
        _lib.trace_enable(0)
        per_step = [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]
        total = max_over_ranks(ev[0].elapsed_time(ev[steps]))
        per_launch = max(1, len(kern) // max(steps, 1))
        kern_step = [sum(kern[i * per_launch:(i + 1) * per_launch]) for i in range(len(kern) // per_launch)]
        out = {"ms_total": total, "ms_per_step": total / steps, "step_ms": summary(per_step),
               "kernel_ms": summary(kern_step), "kernel_launches_per_step": per_launch,
               "kernel_ms_mean": max_over_ranks(sum(kern) / max(steps, 1)), "launches": int(launches)}
        if sampler is not None:
            out["clocks"] = sampler.stop(t0, t1)
        return out

    # ---- headline: device-resident timing, kernel traced inside the same steps ----------------------------------
    head = timed_loop(step, args.steps, args.warmup, sample_clocks=True)
    ms_per_step = head["ms_per_step"]
    kernel_ms = head["kernel_ms_mean"]              # mean over the SAME timed steps (max over ranks)

    # ---- end to end through the public API with host buffers ----------------------------------------------------
    # (one second of idle first: the headline loop has warmed the GPU towards its power cap, and the two loops should
    #  start from the same state — the pause is outside every timed region)
    barrier()
    time.sleep(1.0)
    e2e = timed_loop(step_e2e, args.steps, args.warmup)

    secondary = {}
    traffic = load_profile_json("ncu_traffic.json")

    # ---- multi-GPU: parity of the merged top-k over real NCCL, and where a sharded step's time goes ------------
    parity_check, breakdown = None, None
    if world > 1:
        parity_check = sharded_parity_check(torch, dist, _lib, retr, searcher, queries, rank, world, dev, args)
        breakdown = sharded_breakdown(torch, dist, hrc, _lib, retr, searcher, queries, dev, max_over_ranks, barrier)

    # ---- multi-GPU, pipelined: exchange + merge on a side stream, so a rank's next scan does not wait for the collective ---
    if world > 1 and searcher.transport != "torch" and "pipelined" in want:
        pending = []

        def step_async(i):
            pending.append(searcher.search_keys_async(queries[i % n_q:i % n_q + 1], K))
            if len(pending) > 8:
                pending.pop(0)
        pl = timed_loop(step_async, max(args.steps, 60), args.warmup)
        last = pending[-1].result()
        ref_keys = searcher.search_keys(queries[(max(args.steps, 60) - 1) % n_q:(max(args.steps, 60) - 1) % n_q + 1], K)
        secondary["pipelined"] = {
            "what": "the headline step issued through ShardedSearcher.search_keys_async: local MaxSim + top-k on the main "
                    "stream, exchange + merge on a side stream (a stream of independent queries; same work per step)",
            "ms_per_step": pl["ms_per_step"], "docs_per_s": n_global / (pl["ms_per_step"] * 1e-3), "steps": max(args.steps, 60),
            "kernel_ms": pl["kernel_ms"], "last_result_equals_synchronous_search": bool(torch.equal(last, ref_keys))}

    # ---- secondary: sustained C2, read peak, C4 on the C2 corpus -------------------------------------------------
    if "sustained" in want:
        n_sus = max(50, int(2200.0 / max(ms_per_step, 0.1)))          # >= 2 s back to back
        sus = timed_loop(step, n_sus, 3, sample_clocks=True, clock_period_s=0.025)
        gbs = 256.0 * store.total_tokens / (sus["kernel_ms_mean"] * 1e-3) / 1e9
        secondary["sustained"] = {
            "what": f"C2 step back to back for {sus['ms_total'] / 1e3:.2f} s ({n_sus} steps), same code as the headline",
            "ms_per_step": sus["ms_per_step"], "docs_per_s": n_global / (sus["ms_per_step"] * 1e-3),
            "kernel_ms": sus["kernel_ms"], "kernel_ms_mean": sus["kernel_ms_mean"], "achieved_GBps": gbs,
            "frac_hbm_copy_peak": gbs / hbm_peak, "frac_of_nominal_7.7TBps": gbs / 7700.0, "clocks": sus.get("clocks")}
    read_peak = None
    if "read_peak" in want:
        out = torch.zeros(1, dtype=torch.int32, device=dev)
        for _ in range(2):
            _lib.read_probe(store.tokens, out)
        barrier()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record()
        for _ in range(10):
            _lib.read_probe(store.tokens, out)
        r1.record()
        barrier()
        read_ms = max_over_ranks(r0.elapsed_time(r1)) / 10
        read_peak = store.total_tokens * 256.0 / (read_ms * 1e-3) / 1e9
        secondary["read_peak"] = {"what": "hrc_read_probe: 16-byte vector loads over the same 32.8 GB corpus, 10 passes",
                                  "ms": read_ms, "GBps": read_peak,
                                  "tma_ring_read": load_profile_json("r02_tma_ring_read.json")}
    if "c4" in want and world == 1:
        secondary["c4"] = bench_c4(torch, hrc, _lib, retr, dev, synth_queries)
        # the sequential pipeline runs a C2 search per query and is measured over 5 s, i.e. at the power cap: compare it with
        # the SUSTAINED C2 step (same thermal state), not with the 0.1 s burst of the headline
        base = secondary.get("sustained", {}).get("ms_per_step")
        secondary["c4"]["sequential"]["vs_sustained_c2_step"] = (secondary["c4"]["sequential"]["ms_per_query"] / base) if base else None
        secondary["c4"]["sequential"]["vs_headline_c2_step"] = secondary["c4"]["sequential"]["ms_per_query"] / ms_per_step

    # ---- secondary on the ragged C3 corpus (the C2 corpus is released first) -------------------------------------
    n_docs_c2_local, tokens_c2_local = store.n_docs, store.total_tokens
    if world == 1 and want & {"c3", "c1", "ragged"}:
        _lib.store_release(store.tokens)
        retr.store = None
        del store
        torch.cuda.empty_cache()
        rag = synth_store(args.docs_per_gpu, 32, 512, seed=SEED_C3, device=dev)
        retr.store = rag
        if "c3" in want:
            secondary["c3"] = bench_c3(torch, _lib, retr, rag, dev, synth_queries, tf_sus, tf_burst, timed_loop)
        if "ragged" in want:
            rg = timed_loop(lambda i: retr.search_keys(queries[i % n_q:i % n_q + 1], K), 10, 3)
            gbs = 256.0 * rag.total_tokens / (rg["kernel_ms_mean"] * 1e-3) / 1e9
            secondary["ragged"] = {"what": f"single query over {rag.n_docs} passages x U(32..512) tokens ({rag.total_tokens} tokens)",
                                   "ms_per_step": rg["ms_per_step"], "kernel_ms": rg["kernel_ms"], "achieved_GBps": gbs,
                                   "frac_hbm_copy_peak": gbs / hbm_peak, "docs_per_s": rag.n_docs / (rg["ms_per_step"] * 1e-3)}
        if "c1" in want:
            secondary["c1"] = bench_c1(torch, _lib, retr, rag, dev, queries)
        _lib.store_release(rag.tokens)
        retr.store = None
        del rag
        torch.cuda.empty_cache()

    # ---- C5: 10M passages x 128 tokens over the ranks (needs N > 1: 41 GB per GPU at N = 8) ---------------------
    if world > 1 and "c5" in want:
        per_rank_gb = args.c5_global_docs / world * DOC_LEN * 256 / 1e9
        free_b, _ = torch.cuda.mem_get_info()
        if per_rank_gb * 1e9 < free_b + n_docs_c2_local * DOC_LEN * 256 - 8e9:
            _lib.store_release(retr.store.tokens)
            retr.store = None
            store = None
            torch.cuda.empty_cache()
            c5 = synth_store(args.c5_global_docs, DOC_LEN, DOC_LEN, seed=SEED + 3, device=dev, rank=rank, world_size=world)
            retr.store = c5
            r5 = timed_loop(lambda i: searcher.search_keys(queries[i % n_q:i % n_q + 1], K), 20, 3, sample_clocks=True)
            gbs = 256.0 * c5.total_tokens / (r5["kernel_ms_mean"] * 1e-3) / 1e9
            pc5 = sharded_parity_check(torch, dist, _lib, retr, searcher, queries, rank, world, dev, args,
                                       n_global=args.c5_global_docs, seed=SEED + 3)
            secondary["c5"] = {"what": f"C5: {args.c5_global_docs} passages x {DOC_LEN} tokens "
                                       f"({args.c5_global_docs * DOC_LEN * 256 / 1e9:.1f} GB bf16) over {world} GPUs, "
                                       f"{c5.n_docs} per GPU; local top-{K} + NCCL all-gather + merge",
                               "ms_per_step": r5["ms_per_step"], "docs_per_s": args.c5_global_docs / (r5["ms_per_step"] * 1e-3),
                               "kernel_ms": r5["kernel_ms"], "per_gpu_GBps": gbs, "frac_hbm_copy_peak": gbs / hbm_peak,
                               "frac_of_nominal_7.7TBps": gbs / 7700.0, "clocks": r5.get("clocks"), "parity_check": pc5}
            if "c4" in want:
                secondary["c4_sharded"] = bench_c4_sharded(torch, hrc, _lib, retr, searcher, dev, synth_queries, timed_loop,
                                                           args.c5_global_docs, world)
        else:
            secondary["c5"] = {"skipped": f"{per_rank_gb:.0f} GB per GPU does not fit at N={world}"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    docs_per_step = n_global
    value = docs_per_step / (ms_per_step * 1e-3)
    algo_bytes = 256.0 * tokens_c2_local                            # 256 B per document token
    achieved = algo_bytes / (kernel_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic", "config": workload_config(args, world),
        "tokens_per_s": value * DOC_LEN,
        "step_ms": head["step_ms"],
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                     "traffic": traffic["bytes_per_launch"] if traffic else None,
                     "kernel": "maxsim_dm_kernel<TK=1> (doc-major single-query MaxSim, top-k fused into the epilogue)", "kernel_ms": kernel_ms,
                     "kernel_ms_per_step": head["kernel_ms"], "kernel_launches_per_step": head["kernel_launches_per_step"],
                     "kernel_timing": "CUDA events around the kernel's launches INSIDE the timed steps (hrc_trace_*), "
                                      "mean over the same steps as ms_per_step",
                     "algorithmic_bytes_per_launch": algo_bytes, "peak_source": peak_src,
                     "read_peak_gbs": read_peak, "frac_vs_read_peak": (achieved / read_peak) if read_peak else None,
                     "frac_of_nominal": {"7.7TB/s_hgx": achieved / 7700.0, "8.0TB/s_dgx": achieved / 8000.0},
                     "kernel_share_of_step": kernel_ms / ms_per_step},
        "e2e": {"value": docs_per_step / (e2e["ms_per_step"] * 1e-3), "unit": UNIT,
                "h2d_bytes_per_step": int(q_host[0].numel() * 4), "d2h_bytes_per_step": K * 8,
                "ms_per_step": e2e["ms_per_step"], "step_ms": e2e["step_ms"], "kernel_ms": e2e["kernel_ms"],
                "api": ("JinaColBERTRetriever.search_host: pinned fp32 query -> hrc_search_host (H2D, bf16, MaxSim, top-k, "
                        "unpack, D2H) -> ids/scores on the host" if world == 1 else
                        f"ShardedSearcher.search_host (transport {searcher.transport}): pinned fp32 query -> hrc_sharded_search_host "
                        "(H2D, bf16, local MaxSim + top-k, exchange of k keys per rank, merge + unpack, D2H), one C call per rank")},
        "gpu_launches": head["launches"],
        "clocks": head.get("clocks"),
    }
    if parity_check is not None:
        line["parity_check"] = parity_check
    if breakdown is not None:
        line["breakdown"] = breakdown
    if secondary:
        line["secondary"] = secondary
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = time_cpu(args.cpu_sample_docs)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------------
# secondary configs
# --------------------------------------------------------------------------------------------------
def cuda_time(torch, fn, steps, warmup=3):
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


def bench_c4(torch, hrc, _lib, retr, dev, synth_queries):
    """C4: BM25 top-100 (synthesised: bm25s is unavailable offline) + ColBERT top-100 -> RRF(60) top-50 -> rerank
    top-10, 1,000 queries over the C2 corpus — one query at a time (HBM-bound) and in batches of 64 (tensor-bound)."""
    store = retr.store
    cfg = hrc.RAGConfig(device=str(dev))
    idx = hrc.DualIndexer(cfg)
    idx.colbert_retriever = retr
    h = hrc.HybridRetriever(cfg, idx, None, verbose=False)
    n_queries = 1000
    queries = synth_queries(n_queries, LQ, seed=SEED + 11, device=dev)
    g = torch.Generator().manual_seed(4)
    bm25 = torch.randint(0, store.n_docs, (n_queries, 100), generator=g, dtype=torch.int32).to(dev)
    out = {"what": f"C4 hybrid pipeline x {n_queries} queries over {store.n_docs} passages x {DOC_LEN} tokens: ColBERT "
                   "top-100 + given BM25 top-100 -> RRF(60) top-50 -> rerank top-10 (hrc_hybrid_retrieve, one C call per batch)"}
    for batch in (1, 64):
        def run():
            for b in range(0, n_queries, batch):
                h.retrieve_batch(queries[b:b + batch], bm25[b:b + batch], top_k_final=10)
        for b in range(0, 3 * batch, batch):                    # warm-up: 3 batches
            h.retrieve_batch(queries[b:b + batch], bm25[b:b + batch], top_k_final=10)
        torch.cuda.synchronize()
        l0 = _lib.launch_count()
        ms = cuda_time(torch, run, 1, warmup=0)
        key = "sequential" if batch == 1 else f"batch{batch}"
        out[key] = {"ms_per_query": ms / n_queries, "queries_per_s": n_queries / (ms * 1e-3), "queries_timed": n_queries,
                    "launches_per_call": (_lib.launch_count() - l0) / (n_queries / batch)}
    return out


def bench_c4_sharded(torch, hrc, _lib, retr, searcher, dev, synth_queries, timed_loop, n_global, world):
    """C4 over the C5 corpus (N > 1): the hybrid pipeline with the corpus sharded by documents — global ColBERT top-100
    (local search + exchange + merge) -> RRF with the given BM25 lists (global ids) -> every rank scores the candidates
    it owns -> exchange + merge of the rerank keys; one C call per rank and query (hrc_sharded_hybrid_retrieve).
    parity_check: the library's result equals the same stages done with torch.distributed collectives, on every rank."""
    n_queries = 100
    queries = synth_queries(n_queries, LQ, seed=SEED + 12, device=dev)
    g = torch.Generator().manual_seed(5)                          # same lists on every rank
    bm25 = torch.randint(0, n_global, (n_queries, 100), generator=g, dtype=torch.int32).to(dev)
    r = timed_loop(lambda i: searcher.retrieve_batch(queries[i:i + 1], bm25[i:i + 1], top_k_final=10), n_queries - 3, 3)
    cross = hrc.ShardedSearcher(retr, transport="torch")
    same = True
    for i in (0, 7, 50):
        a_ids, a_sc = searcher.retrieve_batch(queries[i:i + 2], bm25[i:i + 2], top_k_final=10)
        b_ids, b_sc = cross.retrieve_batch(queries[i:i + 2], bm25[i:i + 2], top_k_final=10)
        same = same and bool(torch.equal(a_ids, b_ids)) and bool(torch.equal(a_sc, b_sc))
    flag = torch.tensor([1 if same else 0], device=dev)
    torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN)
    return {"what": f"C4 hybrid pipeline over the C5 corpus ({n_global} passages x {DOC_LEN} tokens on {world} GPUs), "
                    f"{n_queries - 3} queries one at a time: hrc_sharded_hybrid_retrieve (transport {searcher.transport})",
            "ms_per_query": r["ms_per_step"], "queries_per_s": 1e3 / r["ms_per_step"], "kernel_ms": r["kernel_ms"],
            "kernel_ms_mean": r["kernel_ms_mean"], "step_minus_kernels_ms": r["ms_per_step"] - r["kernel_ms_mean"],
            "note": "kernel_ms = the two MaxSim kernels of a query (shard scan + owned candidates), traced inside the steps; "
                    "the run lasts seconds, so it sits at the power cap like `sustained` (compare with kernel_ms_mean, not "
                    "with the 20-step c5 figure)",
            "launches_per_query": r["launches"] / max(1, n_queries - 3),
            "parity_check": "ok" if bool(flag[0]) else "library result differs from the torch.distributed cross-check"}


def bench_c3(torch, _lib, retr, rag, dev, synth_queries, tf_sus, tf_burst, timed_loop):
    """C3: 256 queries x 32 tokens over 1M passages of 32..512 tokens — the tensor-bound config."""
    nq = 256
    q = synth_queries(nq, LQ, seed=SEED_C3 + 1, device=dev)
    r = timed_loop(lambda i: retr.search_keys(q, K), 5, 3, sample_clocks=True, clock_period_s=0.025)
    flops = 2.0 * LQ * 128 * nq * rag.total_tokens                  # useful flops only (no M padding)
    k_ms = r["kernel_ms_mean"]
    tfs = flops / (k_ms * 1e-3) / 1e12
    return {"what": f"C3: {nq} queries x {LQ} tokens over {rag.n_docs} passages x U(32..512) tokens ({rag.total_tokens} tokens, "
                    f"{rag.total_tokens * 256 / 1e9:.1f} GB); step = MaxSim + top-{K} per query",
            "ms_per_step": r["ms_per_step"], "kernel_ms": r["kernel_ms"], "kernel_ms_mean": k_ms,
            "step_minus_kernel_ms": r["ms_per_step"] - k_ms, "useful_tflops": tfs,
            "pairs_per_s": nq * rag.n_docs / (r["ms_per_step"] * 1e-3),
            "roofline": {"bound": "tensor", "achieved": tfs, "peak": tf_sus, "unit": "TFLOP/s", "frac": tfs / tf_sus,
                         "frac_of_burst": tfs / tf_burst, "kernel": "maxsim_tc_kernel<MT=2,ZP=0,CG=2>",
                         "useful_flops_per_launch": flops,
                         "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (the kernel runs 0.4 s per launch, at the power cap)"},
            "clocks": r.get("clocks"), "launches_per_step": r["launches"] / 5}


def bench_c1(torch, _lib, retr, rag, dev, queries):
    """C1: ColBERT rerank of 50 candidates (docs 32..512 tokens), 1 query x 32 tokens: latency of one call."""
    g = torch.Generator().manual_seed(0)
    cand = torch.randint(0, rag.n_docs, (1, 50), generator=g, dtype=torch.int32).to(dev)
    q = queries[0:1]
    toks = int(rag.lengths()[cand[0].long()].sum())
    ws = _lib.Workspace()
    out = {"what": f"C1: rerank of 50 candidates ({toks} tokens), 1 query x {LQ} tokens, top-10; 200 calls back to back",
           "candidate_tokens": toks}
    for name, path in (("tc", _lib.PATH_TC), ("simt", _lib.PATH_SIMT)):
        us_score = cuda_time(torch, lambda: _lib.maxsim_scores_ids(rag.tokens, rag.offsets, cand, q, path=path, workspace=ws), 200) * 1e3
        us_call = cuda_time(torch, lambda: _lib.rerank(rag.tokens, rag.offsets, cand, q, 10, path=path, workspace=ws), 200) * 1e3
        out[name] = {"score_kernel_call_us": us_score, "rerank_call_us": us_call}
    us_api = cuda_time(torch, lambda: retr.rerank_ids(q, cand, k=10), 200) * 1e3
    out["rerank_ids_us"] = us_api
    out["docs_per_s"] = 50 / (us_api * 1e-6)
    # the same launch replayed from a CUDA graph (GraphedRerank: inputs written in place, no Python / ctypes per call)
    import hybrid_rag_colbertv2_b200 as hrc
    plan = hrc.GraphedRerank(retr, 1, 50, k=10)
    plan.queries.copy_(q)
    plan.candidates.copy_(cand)
    want = [t.clone() for t in retr.rerank_ids(q, cand, k=10)]
    got = plan.run()
    torch.cuda.synchronize()
    out["graph_replay_us"] = cuda_time(torch, plan.run, 200) * 1e3
    out["graph_replay_equals_direct_call"] = all(bool(torch.equal(a, b)) for a, b in zip(got, want))
    return out


# --------------------------------------------------------------------------------------------------
# multi-GPU checks (outside every timed region)
# --------------------------------------------------------------------------------------------------
def sharded_parity_check(torch, dist, _lib, retr, searcher, queries, rank, world, dev, args, n_global=None, seed=SEED):
    """Over real NCCL: (1) the merged key list equals the CPU merge (oracle.merge_keys) of every rank's local top-k,
    bit for bit, on every rank; (2) the documents it names, regenerated from the counter-based corpus generator and
    re-scored by the CPU oracle, carry the returned scores and are in the oracle's order.  -> "ok" or the reason."""
    import numpy as np
    n_global = n_global or args.docs_per_gpu * world
    try:
        problems = []
        for qi in (0, 5):
            q = queries[qi:qi + 1]
            merged = searcher.search_keys(q, K)                               # [1, K]
            local = retr.search_keys(q, K)
            if local.shape[1] < K:
                local = torch.cat([local, torch.zeros((1, K - local.shape[1]), dtype=local.dtype, device=dev)], 1)
            allk = [torch.empty_like(local) for _ in range(world)]
            dist.all_gather(allk, local.contiguous())
            allm = [torch.empty_like(merged) for _ in range(world)]
            dist.all_gather(allm, merged.contiguous())
            if rank != 0:
                continue
            from oracle import maxsim_oracle as o                             # the checker, never the thing measured
            cat = torch.cat(allk, 1).cpu().numpy().view(np.uint64)
            want = o.merge_keys(cat, K)
            got = merged.cpu().numpy().view(np.uint64)
            if not (got == want).all():
                problems.append(f"query {qi}: merged keys differ from the CPU merge of the local lists")
            if not all(torch.equal(m, merged) for m in allm):
                problems.append(f"query {qi}: ranks hold different merged lists")
            ids, scores = o.unpack_keys(got[0])
            if len(set(ids.tolist())) != K or ids.min() < 0 or ids.max() >= n_global:
                problems.append(f"query {qi}: ids not unique / out of range")
                continue
            # regenerate the named documents (uniform 128-token passages: tokens [id * 128, id * 128 + 128))
            rows = torch.empty((K * DOC_LEN, 128), dtype=torch.bfloat16, device=dev)
            for j, d in enumerate(ids.tolist()):
                _lib.synth_tokens(rows[j * DOC_LEN:(j + 1) * DOC_LEN], int(d) * DOC_LEN, seed)
            exp = o.maxsim_scores(q.float().cpu(), rows.float().cpu(), torch.arange(0, K * DOC_LEN + 1, DOC_LEN))[0]
            err = o.check_ranking(list(range(K)), scores.tolist(), exp, K, 1e-3)
            if err is not None:
                problems.append(f"query {qi}: oracle re-score: {err}")
        flag = torch.tensor([len(problems)], device=dev)
        dist.broadcast(flag, 0)
        if rank != 0:
            return None
        return "ok" if not problems else "; ".join(problems)
    except Exception as exc:  # noqa: BLE001
        return f"parity check raised {type(exc).__name__}: {exc}"


def sharded_breakdown(torch, dist, hrc, _lib, retr, searcher, queries, dev, max_over_ranks, barrier):
    """Device time of the parts of a sharded step, each timed alone (20 reps back to back, CUDA events, max over ranks):
    the local search, and the exchange + merge with each transport."""
    from hybrid_rag_colbertv2_b200.sharded import all_gather_keys
    q = queries[0:1]
    local = retr.search_keys(q, K).contiguous()
    ws = _lib.Workspace()
    p2p = searcher if searcher.transport == "p2p" else None
    if p2p is None:
        try:                                                    # fails on every rank or on none (hrc_comm_enable_p2p)
            p2p = hrc.ShardedSearcher(retr, transport="p2p")
        except _lib.HrcError:
            p2p = None
    extra = hrc.ShardedSearcher(retr, transport="nccl") if (searcher.comm is None and p2p is None) else None
    nccl_comm = searcher.comm if searcher.comm is not None else (p2p.comm if p2p is not None else extra.comm)
    parts = {"local_search_us": lambda: retr.search_keys(q, K),
             "exchange_merge_nccl_us": lambda: _lib.allgather_merge_topk(nccl_comm, local, K, transport=_lib.TRANSPORT_NCCL, workspace=ws),
             "exchange_merge_torch_us": lambda: _lib.topk_merge(all_gather_keys(local, K), K),
             "step_nccl_us": lambda: _lib.sharded_search(nccl_comm, retr.store.tokens, retr.store.offsets, q, K,
                                                         id_base=retr.store.doc_id_base, workspace=ws, unpack=False)}
    if p2p is not None:
        parts["exchange_merge_p2p_us"] = lambda: _lib.allgather_merge_topk(p2p.comm, local, K, transport=_lib.TRANSPORT_P2P, workspace=ws)
        parts["step_p2p_us"] = lambda: _lib.sharded_search(p2p.comm, retr.store.tokens, retr.store.offsets, q, K,
                                                           id_base=retr.store.doc_id_base, transport=_lib.TRANSPORT_P2P,
                                                           workspace=ws, unpack=False)
    out = {}
    for name, fn in parts.items():
        for _ in range(3):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        barrier()
        out[name] = max_over_ranks(e0.elapsed_time(e1)) / 20 * 1e3
    out["overhead_us"] = out[f"exchange_merge_{searcher.transport}_us"]
    out["transport"] = searcher.transport
    # what the exchange adds to a whole search, seen where the search itself is short (4,000 passages per rank): the
    # difference of two ~60 us numbers instead of two ~4,500 us ones
    from hybrid_rag_colbertv2_b200.synth import synth_store
    world, rank = dist.get_world_size(), dist.get_rank()
    small = hrc.JinaColBERTRetriever(retr.config)
    small.store = synth_store(4000 * world, DOC_LEN, DOC_LEN, seed=SEED + 21, device=dev, rank=rank, world_size=world)
    small_searcher = hrc.ShardedSearcher(small, transport=searcher.transport)
    small_parts = {"local": lambda: small.search_keys(q, K), "sharded": lambda: small_searcher.search_keys(q, K)}
    small_us = {}
    for name, fn in small_parts.items():
        for _ in range(10):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200):
            fn()
        e1.record()
        barrier()
        small_us[name] = max_over_ranks(e0.elapsed_time(e1)) / 200 * 1e3
    out["small_shard"] = {"what": f"4,000 passages per rank, transport {searcher.transport}, 200 searches back to back",
                          "local_search_us": small_us["local"], "sharded_search_us": small_us["sharded"],
                          "exchange_overhead_us": small_us["sharded"] - small_us["local"]}
    small_searcher.close()
    out["note"] = ("each part timed alone, 20 reps back to back; exchange = k x 8 bytes per rank; nccl = ncclAllGather + merge "
                   "kernel, p2p = push kernel (peer stores + release flag) + merge kernel (acquires the flags), both inside libhrc; "
                   "step_p2p_us is the whole sharded search with the exchange FUSED into the search's final selection kernel "
                   "(peer stores, flags, wait, merge: 2 launches per search, scripts/check_sharded_nccl.py small_shard_step_us)")
    return out


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
