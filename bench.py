#!/usr/bin/env python
"""Headline benchmark: queries/sec at 1M x 384-d, top_k=5 (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch 256] [--impl reference]

A *step* is one pass of the hot path over one batch of `--batch` synthetic unit-norm queries:
prepare (normalise) -> score against the resident corpus -> select -> exact re-rank -> results.
Every timed region is exactly K steps between two CUDA events on the launching stream (barrier + synchronize on both
sides, max over ranks); the region is REPEATED until half a second of device time has been measured, `ms_per_step` is the
median region / K, and the minimum and the p99 are reported beside it -- so the nvidia-smi clock sampler sees the GPU
under the measured load, not a 4 ms blip.

  N = 1   value    configs[1], device-resident inputs and outputs
          e2e      the same through the C ABI with HOST buffers (pinned query batch in, ids / distances / counts out),
                   host<->device copies inside the timed region (pipelined b2r_query_async / b2r_wait; the blocking
                   b2r_query beside it)
          batch1, ingest, config0, config3 (1M x 512 + type filter / compiled clause, k=10), config5 (10M x 768,
          interleaved upserts and batch-64 queries, k=20): the other BASELINE configs, each with its roofline
  N > 1   one process per GPU (torchrun), the ROW-SHARDED path north_star asks for: every rank owns its own 1M x 384
          shard (N M rows in all), the query batch is replicated, each rank answers from its shard with the exact engine,
          ONE NCCL all_gather exchanges the [nq,k] lists and every rank merges.  `value` counts what the ranks processed:
          N x batch shard-queries per step (each rank scores the batch against its own 1M x 384 shard -- the metric's unit
          of work); the merged answers, which cover N M rows, come out at value / N per second (`merged_queries_per_s`).
          Weak scaling: per-GPU work is fixed, so value / (N x value(1 GPU)) is the price of the exchange.
          The run checks itself: the merged rows of one batch must equal a single-GPU answer over the union of the shards.
          Beside it: `replicas` (corpus replicated, queries sharded, no collective -- round 1's headline) and
          `sharded_c4` (BASELINE config 4: 100M x 384 row-sharded over the N GPUs, batch 1024, top_k 100 -- strong
          scaling of a fixed corpus), also self-checked against the exact fp64 scan on a sample.
  --impl reference   the CPU arm: the reference's query path restated in C (oracle/exact_topk.c: hnswlib's L2Sqr /
          InnerProduct arithmetic, fp32 accumulate, exhaustive) on every host core, full 256-query batches against the
          1M x 384 corpus.  The reference's real stack (chromadb 0.4.22 -> chroma-hnswlib 0.7.3) is not in this image
          (probed at run time: `import chromadb`, also under baseline/_ref), so this is kind="port".
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_ROWS, DIM, TOP_K = 1_000_000, 384, 5
HNSW_ROWS = 100_000          # bounded prefix of the corpus the HNSW baseline is built on (~10-20 s of CPU)
METRIC, UNIT = "queries/sec @1Mx384-d top_k=5", "queries/s"
MIN_TIMED_S = 0.5            # every timed region is repeated until this much device time has been measured


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return {"hbm_gbs": j["hbm_gbs"], "bf16_tflops": j["bf16_tflops"],
                "bf16_tflops_sustained": j.get("bf16_tflops_sustained", j["bf16_tflops"]), "src": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "src": "fallback"}


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` at this workload, from the committed
    `ncu --set full` capture (profiles/roofline_traffic.json); None when no capture is recorded."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        return json.load(open(p)).get(kernel, {}).get("dram_bytes_per_launch")
    except (OSError, ValueError):
        return None


def headline_config(world, nq):
    """The `config` object of the JSON line -- the GPU arm and the CPU arm print the same one for the same N."""
    cfg = {"workload": f"configs[1]: {N_ROWS}x{DIM} bf16 corpus (+fp32 master for the exact re-rank), batch {nq}, "
                       f"top_k={TOP_K}, cosine",
           "rows": N_ROWS, "dim": DIM, "batch": nq, "top_k": TOP_K, "space": "cosine",
           "l2_policy": "corpus (768 MB) is larger than L2 (126 MB); query batches rotate"}
    if world == 1:
        cfg["parallelism"] = "single GPU"
    else:
        cfg["parallelism"] = (f"row-sharded: {world} shards of {N_ROWS}x{DIM} ({world * N_ROWS} rows in all), one per GPU; queries "
                              f"replicated; ONE exchange of the per-rank [batch, top_k] lists + merge per batch (NCCL all_gather + merge kernel, or the "
                              f"library's fused peer-to-peer form: comm.form_timed says which was measured); value "
                              f"counts {world} x batch shard-queries per step (weak scaling: fixed work per GPU), the merged answers "
                              f"over all rows come out at value / {world}")
        cfg["rows_total"] = world * N_ROWS
    return cfg


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap,utilization.gpu")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "50"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, busy = [], [], set(), 0
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 10:
                continue
            try:
                util = float(f[9])
                clk, cmax = float(f[1]), float(f[2])
            except ValueError:
                continue
            if util < 50:                     # only samples taken while the GPU was under the measured load
                continue
            busy += 1
            sm.append(clk); mx.append(cmax)
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": busy, "samples_total": len(self.lines),
                "sampled": "nvidia-smi every 50 ms across every timed region of this run; samples with GPU utilisation >= 50 % kept"}


_REAL_STDOUT = None


def claim_stdout():
    """stdout carries exactly ONE JSON line.  Libraries that print there (NCCL's version banner, OpenMP notices)
    are sent to stderr: fd 1 is pointed at fd 2 for the life of the process, the JSON line goes to the saved fd."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (json.dumps(line) + "\n").encode())


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def summarize(regions_ms, K):
    """regions_ms: device time of each repeat of the K-step region.  Median / min / p99 per step, and the FIRST region on its
    own: exactly K steps once, a few milliseconds, the GPU not yet at its power cap -- what a single short region (round 1's
    methodology) reports."""
    r = sorted(regions_ms)
    n = len(r)
    return {"ms_per_step": r[n // 2] / K, "ms_per_step_min": r[0] / K, "ms_per_step_p99": r[min(n - 1, int(0.99 * n))] / K,
            "ms_per_step_first_region": regions_ms[0] / K, "repeats": n, "timed_region_s": sum(r) * 1e-3}


# ------------------------------------------------------------------------------------------
# CPU arm
# ------------------------------------------------------------------------------------------
def cpu_corpus(n, d, seed):
    import numpy as np
    rng = np.random.default_rng(seed)
    X = np.empty((n, d), dtype=np.float32)
    step = 1 << 17
    for s in range(0, n, step):
        X[s:s + step] = rng.standard_normal((min(step, n - s), d), dtype=np.float32)
    X /= np.linalg.norm(X, axis=1, keepdims=True)
    return X


def cpu_port_qps(X, Q, k, steps, warmup=0):
    """Oracle C port, fp32 accumulate, all OpenMP threads.  Returns (qps, seconds/step, threads)."""
    from oracle import c_oracle
    for _ in range(warmup):
        c_oracle.topk(X, Q, k, "cosine", acc64=False)
    t0 = time.perf_counter()
    for _ in range(steps):
        c_oracle.topk(X, Q, k, "cosine", acc64=False)
    dt = (time.perf_counter() - t0) / steps
    return Q.shape[0] / dt, dt, c_oracle.threads()


def cpu_hnsw_leg(X, Q, k, rows):
    """The index the reference really queries (Chroma -> hnswlib HNSW, defaults M=16 / ef_construction=100 /
    search ef=max(10,k)), restated in oracle/hnsw_port.c, on a bounded prefix of the corpus: speed AND recall@k
    against the exact answer over the same rows.  HNSW is approximate; the engine under test is exact."""
    import numpy as np
    from oracle import c_oracle
    Xs = np.ascontiguousarray(X[:rows])
    t0 = time.perf_counter()
    idx = c_oracle.Hnsw(Xs, "cosine", M=16, ef_construction=100)
    build_s = time.perf_counter() - t0
    exact_rows, _, _ = c_oracle.topk(Xs, Q, k, "cosine", acc64=True)
    out = {"kind": "port (restatement of chroma-hnswlib 0.7.3, Chroma defaults)", "rows": rows, "M": 16,
           "ef_construction": 100, "build_s": build_s, "threads": c_oracle.threads(), "queries": int(Q.shape[0])}
    for ef in (10, 100):
        idx.query(Q[:8], k, ef)
        t0 = time.perf_counter()
        r, _ = idx.query(Q, k, ef)
        dt = time.perf_counter() - t0
        rec = float(np.mean([len(set(r[i].tolist()) & set(exact_rows[i].tolist())) / k for i in range(Q.shape[0])]))
        out[f"ef{ef}"] = {"qps": Q.shape[0] / dt, f"recall_at_{k}": rec}
    out["note"] = ("isotropic synthetic unit vectors are HNSW's worst case (no low-dimensional structure): recall at "
                   "Chroma's default ef=10 is a few percent; real sentence embeddings cluster and score higher")
    idx.close()
    return out


def probe_chroma():
    """The reference's real vector store, if an install ever appears (this image has none): chromadb from the
    environment or from baseline/_ref.  Returns the module or None."""
    ref = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isdir(ref) and ref not in sys.path:
        sys.path.append(ref)
    try:
        import chromadb                                           # noqa: F401
        return chromadb
    except Exception:                                             # noqa: BLE001  (ImportError, or a broken install)
        return None


def chroma_leg(chromadb, X, Q, k, rows):
    """Time the reference's own call (collection.query as app/utils/embedder.py:595-601 makes it) on a bounded prefix,
    and pin the oracle to it: the distances Chroma returns for the ids it returns must be the oracle's."""
    import numpy as np
    from oracle import c_oracle
    client = chromadb.Client()
    col = client.create_collection("bench", metadata={"hnsw:space": "cosine"})
    ids = [f"r{i}" for i in range(rows)]
    t0 = time.perf_counter()
    for s in range(0, rows, 4096):
        col.add(ids=ids[s:s + 4096], embeddings=X[s:s + 4096].tolist())
    build_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    res = col.query(query_embeddings=Q.tolist(), n_results=k, include=["distances"])
    dt = time.perf_counter() - t0
    er, ed, _ = c_oracle.topk(np.ascontiguousarray(X[:rows]), Q, k, "cosine", acc64=True)
    rec, worst = [], 0.0
    for i in range(Q.shape[0]):
        got = [int(s[1:]) for s in res["ids"][i]]
        rec.append(len(set(got) & set(er[i].tolist())) / k)
        for r_, d_ in zip(got, res["distances"][i]):
            exact = 1.0 - float(np.dot(Q[i].astype(np.float64), X[r_].astype(np.float64)))
            worst = max(worst, abs(exact - d_) / max(abs(exact), 1e-12))
    return {"kind": "reference (chromadb %s)" % getattr(chromadb, "__version__", "?"), "rows": rows, "build_s": build_s,
            "qps": Q.shape[0] / dt, f"recall_at_{k}": float(np.mean(rec)), "max_rel_distance_error_vs_oracle": worst}


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    claim_stdout()
    import numpy as np
    from oracle import c_oracle
    c_oracle.build()
    threads = c_oracle.set_threads(0)       # torchrun exports OMP_NUM_THREADS=1: take every core this process may use
    X = cpu_corpus(N_ROWS, DIM, 0xC0FFEE)
    Q = cpu_corpus(args.batch, DIM, 0xBEEF)
    qps, dt, threads = cpu_port_qps(X, Q, TOP_K, max(1, args.steps), max(0, args.warmup))
    extra = {}
    if not args.no_hnsw:
        extra["hnsw"] = cpu_hnsw_leg(X, Q, TOP_K, HNSW_ROWS)
    chroma = probe_chroma()
    if chroma is not None:
        try:
            extra["chroma"] = chroma_leg(chroma, X, Q[:64], TOP_K, HNSW_ROWS)
        except Exception as e:                                    # noqa: BLE001
            extra["chroma"] = f"present but failed: {e!r}"
    else:
        extra["chroma"] = "chromadb / chroma-hnswlib not importable in this image (probed; also baseline/_ref)"
    sample = (f"every step = the full {args.batch}-query batch against the full {N_ROWS}x{DIM} fp32 corpus, exhaustive "
              f"fp32 scan (oracle/exact_topk.c, OpenMP, {threads} threads)")
    if world > 1:
        sample += (f"; at N={world} the GPU arm's unit of work is one batch against ONE {N_ROWS}-row shard per rank, which is "
                   "exactly what a step of this arm does (the CPU's cost per shard-query does not depend on N)")
    line = {
        "impl": "reference", "metric": METRIC, "value": qps, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": headline_config(world, args.batch),
        "cpu_baseline": dict({"value": qps, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                              "what": "the exhaustive form of the reference's distance arithmetic; the reference itself asks an "
                                      "approximate HNSW index (see hnsw: speed and recall of its restatement)"}, **extra),
        "e2e": {"value": qps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
def run_gpu(args):
    claim_stdout()                                  # NCCL prints its version banner on stdout
    import numpy as np
    import torch
    import torch.distributed as dist
    from multimodal_rag_b200 import _lib
    from multimodal_rag_b200.sharded import DeviceShard

    rank, local_rank, world = dist_env()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    nq, k, K, W = args.batch, TOP_K, args.steps, max(args.warmup, 3)
    pk = peaks()
    stream = torch.cuda.current_stream().cuda_stream
    F = torch.nn.functional

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def timed(fn, steps, min_s=MIN_TIMED_S, max_repeats=2000):
        """Repeat the `steps`-step region (CUDA events, barrier + synchronize on both sides, max over ranks) until
        min_s of device time is on record.  Returns the list of region times in ms."""
        out, total = [], 0.0
        while True:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            e0.record()
            for i in range(steps):
                fn(i)
            e1.record()
            barrier()
            ms = max_over_ranks(e0.elapsed_time(e1))
            out.append(ms)
            total += ms
            if total >= min_s * 1e3 or len(out) >= max_repeats:
                return out

    def timed_wall(fn, steps, min_s=MIN_TIMED_S, max_repeats=2000):
        """The same for calls that synchronise themselves (host buffers): wall clock around `steps` calls."""
        out, total = [], 0.0
        while True:
            barrier()
            t0 = time.perf_counter()
            fn(steps)
            torch.cuda.synchronize()
            ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
            out.append(ms)
            total += ms
            if total >= min_s * 1e3 or len(out) >= max_repeats:
                return out

    def kernel_time(h, fn, steps):
        """CUDA events bracketed around the scoring-kernel launches inside the library: (ms per step, launches per step)"""
        tot, cnt = ctypes.c_double(), ctypes.c_int64()
        _lib.check(lib.b2r_set_kernel_timing(h, 1))
        _lib.check(lib.b2r_kernel_time_ms(h, ctypes.byref(tot), ctypes.byref(cnt), 1))
        for i in range(steps):
            fn(i)
        torch.cuda.synchronize()
        _lib.check(lib.b2r_kernel_time_ms(h, ctypes.byref(tot), ctypes.byref(cnt), 1))
        _lib.check(lib.b2r_set_kernel_timing(h, 0))
        return tot.value / steps, cnt.value / steps

    def hbm_roofline(kernel, bytes_per_launch, kern_ms, extra=None):
        gbs = bytes_per_launch / (kern_ms * 1e-3) / 1e9 if kern_ms else 0.0
        r = {"bound": "hbm", "kernel": kernel, "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s",
             "frac": gbs / pk["hbm_gbs"], "traffic": None, "peak_src": pk["src"], "kernel_us": kern_ms * 1e3,
             "algorithmic_bytes_per_launch": bytes_per_launch}
        if extra:
            r.update(extra)
        return r

    def fill_shard(sh, rows, dim, seed, ingest_log=None):
        g = torch.Generator(device=dev).manual_seed(seed)
        step = 1 << 18
        for s in range(0, rows, step):
            m = min(step, rows - s)
            x = F.normalize(torch.randn(m, dim, generator=g, device=dev), dim=1)
            if ingest_log is None:
                sh.ingest(x)
                continue
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            sh.ingest(x)                       # K1: fused normalise + bf16 pack + fp32 master, device-resident input
            e1.record()
            e1.synchronize()
            ingest_log.append((m, e0.elapsed_time(e1)))
        torch.cuda.synchronize()

    def unit_queries(n, dim, seed, count):
        g = torch.Generator(device=dev).manual_seed(seed)
        return [F.normalize(torch.randn(n, dim, generator=g, device=dev), dim=1) for _ in range(count)]

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                      # spans every timed region below

    # =====================================================================================
    # configs[1] on one GPU: replicated corpus (same seed on every rank), this rank's own query batches
    # =====================================================================================
    ingest_ms = []
    shard = DeviceShard(DIM, "cosine", capacity=N_ROWS, row_base=0, device=local_rank)
    fill_shard(shard, N_ROWS, DIM, 0xC0FFEE, ingest_ms)
    full = [(m, t) for m, t in ingest_ms[1:] if m == 1 << 18] or ingest_ms      # first call carries one-off setup
    ing_rows, ing_ms = sum(m for m, _ in full), sum(t for _, t in full)
    ing_bytes_per_row = DIM * 4 + DIM * 2 + DIM * 4 + 1       # fp32 in, bf16 + fp32 master + type code out
    ingest = {"kernel": "ingest_kernel (K1: L2-normalise + bf16 pack + fp32 master, 262144-row batches, device input)",
              "rows_per_s": ing_rows / (ing_ms * 1e-3), "algorithmic_bytes_per_row": ing_bytes_per_row,
              "roofline": {"bound": "hbm", "achieved": ing_rows * ing_bytes_per_row / (ing_ms * 1e-3) / 1e9, "peak": pk["hbm_gbs"],
                           "unit": "GB/s", "frac": ing_rows * ing_bytes_per_row / (ing_ms * 1e-3) / 1e9 / pk["hbm_gbs"],
                           "peak_src": pk["src"]}}
    n_batches = 4                              # rotate query batches so no step repeats its predecessor
    Qd = unit_queries(nq, DIM, 0xBEEF + rank, n_batches)
    Qh = [q.cpu().pin_memory() for q in Qd]
    out = shard.alloc_out(nq, k)

    def step_device(i):
        q = Qd[i % n_batches]
        _lib.check(lib.b2r_query(shard.h, q.data_ptr(), nq, k, None, out["rows"].data_ptr(), out["dist"].data_ptr(),
                                 out["cnt"].data_ptr(), stream), "b2r_query")

    for i in range(W):
        step_device(i)
    launches0 = lib.b2r_launch_count(shard.h)
    single = summarize(timed(step_device, K), K)
    gpu_launches = int(round((lib.b2r_launch_count(shard.h) - launches0) / single["repeats"]))
    kern_ms_per_step, launches_per_step = kernel_time(shard.h, step_device, max(K, 50))
    corpus_bytes = N_ROWS * DIM * 2
    replicas = {"what": "corpus replicated on every GPU, each rank answers its own query batches, no data-path collective",
                "qps": world * nq / (single["ms_per_step"] * 1e-3), **single}

    line = {"metric": METRIC, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": headline_config(world, nq)}

    if world == 1:
        # ---- e2e: host buffers through the C ABI ----
        h_rows = torch.empty((nq, k), dtype=torch.int64).pin_memory()
        h_dist = torch.empty((nq, k), dtype=torch.float32).pin_memory()
        h_cnt = torch.empty((nq,), dtype=torch.int32).pin_memory()

        def step_host(i):
            q = Qh[i % n_batches]
            _lib.check(lib.b2r_query(shard.h, q.data_ptr(), nq, k, None, h_rows.data_ptr(), h_dist.data_ptr(),
                                     h_cnt.data_ptr(), stream), "b2r_query")

        lat = []

        def run_blocking(steps):
            for i in range(steps):
                t1 = time.perf_counter()
                step_host(i)                      # returns with the results in the host arrays (the call synchronises)
                lat.append(time.perf_counter() - t1)

        run_blocking(W)
        lat.clear()
        blocking = summarize(timed_wall(run_blocking, K), K)
        assert int(h_cnt.min()) == k
        lat.sort()
        e2e_latency_us = {"p50": lat[len(lat) // 2] * 1e6, "p99": lat[int(0.99 * (len(lat) - 1))] * 1e6, "max": lat[-1] * 1e6,
                          "calls": len(lat)}

        # the same through the pipelined form of the call (b2r_query_async / b2r_wait): two batches in flight, the copies
        # of one overlap the kernels of the other; every step still moves its own inputs and results inside the timed region
        h_out = [(torch.empty((nq, k), dtype=torch.int64).pin_memory(), torch.empty((nq, k), dtype=torch.float32).pin_memory(),
                  torch.empty((nq,), dtype=torch.int32).pin_memory()) for _ in range(2)]

        def run_pipelined(steps):
            prev = None
            for i in range(steps):
                q = Qh[i % n_batches]
                r_, d_, c_ = h_out[i % 2]
                t = ctypes.c_uint64()
                _lib.check(lib.b2r_query_async(shard.h, q.data_ptr(), nq, k, None, r_.data_ptr(), d_.data_ptr(), c_.data_ptr(),
                                               stream, ctypes.byref(t)), "b2r_query_async")
                if prev is not None:
                    _lib.check(lib.b2r_wait(shard.h, prev), "b2r_wait")
                prev = t.value
            _lib.check(lib.b2r_wait(shard.h, prev), "b2r_wait")

        run_pipelined(W)
        pipelined = summarize(timed_wall(run_pipelined, K), K)
        assert int(h_out[0][2].min()) == k and int(h_out[1][2].min()) == k
        step_host(K - 1)                              # the blocking call on the last batch must give the same rows
        assert torch.equal(h_out[(K - 1) % 2][0], h_rows)
        e2e = {"value": nq / (pipelined["ms_per_step"] * 1e-3), "unit": UNIT, "h2d_bytes_per_step": nq * DIM * 4,
               "d2h_bytes_per_step": nq * k * 12 + nq * 4, **pipelined,
               "api": "b2r_query_async + b2r_wait, two batches in flight (copies of one overlap the kernels of the other)",
               "blocking_call": {"value": nq / (blocking["ms_per_step"] * 1e-3), "unit": UNIT, **blocking,
                                 "latency_us_per_call": e2e_latency_us,
                                 "api": "b2r_query with host arrays: H2D copy, kernels, ONE packed D2H copy, stream sync"}}

        # ---- batch-1 scan (the HBM-bound headline of north_star), same corpus ----
        q1 = [q[:1].contiguous() for q in Qd]
        o1 = shard.alloc_out(1, k)

        def step_b1(i):
            q = q1[i % n_batches]
            _lib.check(lib.b2r_query(shard.h, q.data_ptr(), 1, k, None, o1["rows"].data_ptr(), o1["dist"].data_ptr(),
                                     o1["cnt"].data_ptr(), stream), "b2r_query")

        K1 = max(K, 50)
        for i in range(10):
            step_b1(i)
        b1 = summarize(timed(step_b1, K1), K1)
        b1_kern_ms, b1_lps = kernel_time(shard.h, step_b1, K1)
        batch1 = {"qps": 1e3 / b1["ms_per_step"], "us_per_query": b1["ms_per_step"] * 1e3, **b1,
                  "roofline": hbm_roofline("gemm_topk_kernel (K3 streams the corpus for batch 1 too)", corpus_bytes,
                                           b1_kern_ms / max(b1_lps, 1e-9))}

        # ---- configs[0]: the reference's own CPU-runnable case (10k x 384, one cosine query, top_k 5) on the GPU ----
        c0 = DeviceShard(DIM, "cosine", capacity=10_000, row_base=0, device=local_rank)
        fill_shard(c0, 10_000, DIM, 0xC0)

        def step_c0(i):
            q = q1[i % n_batches]
            _lib.check(lib.b2r_query(c0.h, q.data_ptr(), 1, k, None, o1["rows"].data_ptr(), o1["dist"].data_ptr(),
                                     o1["cnt"].data_ptr(), stream), "b2r_query")

        for i in range(10):
            step_c0(i)
        s0 = summarize(timed(step_c0, 200, min_s=0.1), 200)
        config0 = {"workload": "configs[0]: 10000x384, one cosine query, top_k=5 (latency-bound: 7.7 MB of corpus)",
                   "us_per_query": s0["ms_per_step"] * 1e3, "qps": 1e3 / s0["ms_per_step"]}
        c0.close()
    shard.close()

    config3 = config5 = None
    if world == 1 and not args.no_configs:
        config3 = leg_config3(lib, _lib, dev, timed, summarize, kernel_time, hbm_roofline, pk, K)
        config5 = leg_config5(lib, _lib, dev, timed, summarize, hbm_roofline, pk, K)

    # =====================================================================================
    # N > 1: the row-sharded path is the headline
    # =====================================================================================
    sharded_c4 = None
    if world > 1:
        sh2 = DeviceShard(DIM, "cosine", capacity=N_ROWS, row_base=rank * N_ROWS, device=local_rank)
        fill_shard(sh2, N_ROWS, DIM, 0xC0FFEE + 1 + rank)
        Q2 = unit_queries(nq, DIM, 0xBEEF, n_batches)              # replicated queries
        Q2h = [q.cpu().pin_memory() for q in Q2]
        o2 = sh2.alloc_out(nq, k)
        o2b = [sh2.alloc_out(nq, k) for _ in range(2)]
        # the un-pipelined NCCL form first (scan -> all_gather -> merge on one stream): the baseline the library's own exchange replaces
        def step_sharded_nccl(i):
            sh2.query_device(Q2[i % n_batches], k, o2)
        for i in range(W):
            step_sharded_nccl(i)
        shd_nccl = summarize(timed(step_sharded_nccl, K, min_s=0.2), K)
        if args.p2p_exchange:
            sh2.enable_p2p_exchange(nq_max=max(nq, 1024), k_max=128)
        # the fused exchange (b2r_query_push): mailboxes mapped with CUDA IPC; NCCL stays what query_device / _pipelined use
        fused_ok = True
        try:
            sh2.enable_p2p_exchange(nq_max=max(nq, 1024), k_max=128, default=False)
        except Exception as e:                        # no peer access between these GPUs: the NCCL forms remain
            sys.stderr.write(f"[bench] fused exchange unavailable: {e}\n")
            fused_ok = False
        # (enable_p2p_exchange agrees on a failure across the ranks before it raises: fused_ok is the same everywhere)

        union = {}

        def check_form(which):
            """Collective: one batch through exchange form `which`; True when its merged rows / distances equal a single-GPU answer
            over the union of the shards (rank 0 builds the union once)."""
            if which == "fused":
                sh2.query_device_fused(Q2[0], k, o2)
                sh2.drain()
            else:
                sh2.query_device(Q2[0], k, o2)
            torch.cuda.synchronize()
            good = True
            if rank == 0:
                if not union:
                    un = DeviceShard(DIM, "cosine", capacity=world * N_ROWS, row_base=0, device=local_rank, group=None, world=1)
                    for r in range(world):
                        fill_shard(un, N_ROWS, DIM, 0xC0FFEE + 1 + r)
                    ou = un.alloc_out(nq, k)
                    un.query_local(Q2[0], k, ou)
                    torch.cuda.synchronize()
                    union["rows"], union["dist"] = ou["rows"].clone(), ou["dist"].clone()
                    un.close()
                good = (bool(torch.equal(union["rows"], o2["m_rows"])) and int(o2["m_cnt"].min()) == k and
                        bool(torch.allclose(union["dist"], o2["m_dist"], rtol=1e-6, atol=0)))
            flag = torch.tensor([1 if good else 0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            return int(flag.item()) == 1

        def step_sharded_fused(i):
            # the finalize of batch i stores its lists into every rank's mailbox; the merge of batch i rides in the last kernel
            # of batch i+1; the last step of a region merges what is still pending
            sh2.query_device_fused(Q2[i % n_batches], k, o2b[i % 2])
            if i == K - 1:
                sh2.drain()

        def step_sharded_pipelined(i):
            # the exchange of batch i (all_gather + merge, side stream) overlaps the scan of batch i+1; the last step of a
            # region waits for every exchange in flight, so a region ends with all of its merged answers in place
            sh2.query_device_pipelined(Q2[i % n_batches], k, o2b[i % 2])
            if i == K - 1:
                sh2.drain()

        def step_sharded_serial(i):
            # scan -> exchange -> merge on one stream: the latency of one batch
            sh2.query_device(Q2[i % n_batches], k, o2)

        # Both forms are timed briefly and the faster one is measured in full.  Which one wins depends on N: on a side stream
        # NCCL's kernel competes with the next scan for SMs -- a gain at 2 GPUs (+9 us instead of +21 us per step), a loss at 8
        # (profiles/r2_exchange.md).
        for i in range(W):
            step_sharded_pipelined(i)
        sh2.drain()
        for i in range(W):
            step_sharded_serial(i)
        trial = {"pipelined": summarize(timed(step_sharded_pipelined, K, min_s=0.1), K)["ms_per_step"],
                 "serial": summarize(timed(step_sharded_serial, K, min_s=0.1), K)["ms_per_step"]}
        fused_rejected = False
        if fused_ok:
            for i in range(W):
                step_sharded_fused(i)
            sh2.drain()
            if check_form("fused"):                   # a form that answers wrongly is never a candidate
                trial["fused"] = summarize(timed(step_sharded_fused, K, min_s=0.1), K)["ms_per_step"]
            else:
                fused_ok, fused_rejected = False, True
                sys.stderr.write("[bench] the fused exchange's merged rows differ from the single-GPU answer: form dropped\n")
        forms = sorted(trial)
        pick = torch.tensor([forms.index(min(forms, key=trial.get))], device=dev)
        dist.broadcast(pick, 0)
        form = forms[int(pick.item())]
        use_pipelined = form == "pipelined"
        step_sharded = {"pipelined": step_sharded_pipelined, "serial": step_sharded_serial, "fused": step_sharded_fused if fused_ok else None}[form]
        launches0 = lib.b2r_launch_count(sh2.h)
        shd = summarize(timed(step_sharded, K), K)
        shd_serial = {"ms_per_step": trial["serial"]}
        per_rank_ms = [None] * world
        dist.all_gather_object(per_rank_ms, shd["ms_per_step"])
        kern_ms_per_step, launches_per_step = kernel_time(sh2.h, step_sharded, max(K, 50))

        # e2e: pinned host queries in, merged results out to pinned host arrays, every step
        hq = sh2.alloc_host(nq, k)

        def run_sharded_host(steps):
            for i in range(steps):
                sh2.query_host(Q2h[i % n_batches], k, o2, hq)

        run_sharded_host(W)
        shd_e2e = summarize(timed_wall(run_sharded_host, K), K)

        # self-check: the merged rows of batch 0 == a single-GPU answer over the union of the shards (rank 0 builds it),
        # through the form that was timed
        if not check_form(form):
            raise SystemExit("bench: row-sharded result differs from the single-GPU answer over the union")
        detail = f"merged rows/distances of one {nq}-query batch == single-GPU answer over the {world * N_ROWS}-row union"
        bytes_per_rank = o2["layout"][0]
        if form == "fused":
            comm_backend = "nccl for setup and timing reductions only; data path: peer-to-peer stores over NVLink from the query's own kernels"
        elif args.p2p_exchange:
            comm_backend = "nccl for setup and timing reductions; data path: peer-to-peer stores over NVLink (b2r_xchg_push / b2r_xchg_merge)"
        else:
            comm_backend = "nccl"
        line.update({"value": world * nq / (shd["ms_per_step"] * 1e-3), **shd,
                     "merged_queries_per_s": nq / (shd["ms_per_step"] * 1e-3),
                     "ms_per_step_per_rank": per_rank_ms, "verified": True, "verified_how": detail if rank == 0 else None,
                     "comm": {"backend": comm_backend,
                              "nranks": world, "collectives_per_step": 1,
                              "collective": "exchange of the per-rank [batch, top_k] x (int64 row, fp64 distance) + [batch] int32 count lists",
                              "bytes_sent_per_rank_per_step": bytes_per_rank, "bytes_gathered_per_rank_per_step": world * bytes_per_rank,
                              "nvlink_bytes_stored_per_rank_per_step": (world - 1) * nq * (k * 16 + 4) if form == "fused" else None,
                              "exchange": ("fused into the query's kernels (b2r_query_push): the finalize / fix-up kernels store every final list into the "
                                           "peers' mailboxes over NVLink (CUDA IPC), the merge of batch i rides in the last kernel of batch i+1 and the flag words are "
                                           "written by the next call's first kernel; no collective, no exchange kernel and no extra launch between two scans "
                                           "(DeviceShard.query_device_fused)") if form == "fused" else
                                          sh2.exchange_mode + ("; issued on a side stream behind an event so that it overlaps the scan of the next batch "
                                                               "(DeviceShard.query_device_pipelined)" if use_pipelined else "; on the scan's stream (DeviceShard.query_device)"),
                              "form_timed": form, "trial_ms_per_step": trial, "fused_rejected": fused_rejected,
                              "trial_note": "every form is timed for 0.1 s and the fastest one is measured in full",
                              "nccl_all_gather_ms_per_step": shd_nccl["ms_per_step"],
                              "nccl_note": "the same step with torch.distributed all_gather_into_tensor + b2r_merge_shards_packed on one stream"},
                     "e2e": {"value": world * nq / (shd_e2e["ms_per_step"] * 1e-3), "unit": UNIT, **shd_e2e,
                             "h2d_bytes_per_step": nq * DIM * 4, "d2h_bytes_per_step": nq * k * 12 + nq * 4,
                             "api": "DeviceShard.query_host: pinned host batch -> H2D -> b2r_query_ex -> all_gather -> merge -> D2H of the merged "
                                    "rows / distances / counts -> sync, on every rank, every step"},
                     "replicas": replicas})
        sh2.close()

        if not args.no_c4:
            sharded_c4 = leg_config4(args, lib, _lib, dev, dist, rank, world, timed, summarize, sum_over_ranks, pk, K)
    else:
        line.update({"value": nq / (single["ms_per_step"] * 1e-3), **single, "e2e": e2e,
                     "value_first_region": nq / (single["ms_per_step_first_region"] * 1e-3),
                     "value_note": "value = median of the K-step regions repeated for 0.5 s (sustained, power-capped clocks); value_first_region = the "
                                   "first K-step region alone (~4 ms, burst clocks) -- the quantity round 1's single short region measured"})

    clocks = sampler.stop() if rank == 0 else None
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- CPU baseline (rank 0, N=1 only): bounded sample of the same workload ----
    cpu = None
    if world == 1 and not args.no_cpu:
        from oracle import c_oracle
        c_oracle.build()
        thr = c_oracle.set_threads(0)
        Xh = cpu_corpus(N_ROWS, DIM, 0xC0FFEE)
        Qs = Qh[0].numpy()
        qps_c, dt_c, thr = cpu_port_qps(Xh, Qs, k, 2, 0)
        cpu = {"value": qps_c, "unit": UNIT, "cores": thr, "kind": "port",
               "sample": f"full {N_ROWS}-row fp32 corpus, the full {nq}-query batch, 2 passes, "
                         "exhaustive fp32 scan (oracle/exact_topk.c, OpenMP)"}
        if not args.no_hnsw:
            cpu["hnsw"] = cpu_hnsw_leg(Xh, Qs, k, HNSW_ROWS)
            cpu["hnsw_config0"] = cpu_hnsw_leg(Xh, Qs, k, 10_000)     # configs[0]: the reference's own size
        chroma = probe_chroma()
        cpu["chroma"] = ("chromadb / chroma-hnswlib not importable in this image (probed; also baseline/_ref)" if chroma is None
                         else "present: see `bench.py --impl reference`")
        try:
            from oracle import exact_oracle as eo
            t0 = time.perf_counter()
            eo.topk_bruteforce_f32(Qs[:32], Xh, k, "cosine")
            cpu["numpy_sgemm_qps"] = 32 / (time.perf_counter() - t0)
        except Exception as e:                                  # noqa: BLE001
            cpu["numpy_sgemm_qps"] = f"failed: {e}"

    # Dominant kernel of a step: gemm_topk_kernel (K3, tcgen05), ONE launch per step, which streams the
    # packed corpus exactly once: algorithmic bytes per launch = rows * padded_dim * 2 (DESIGN.md "Roofline").
    # At batch 256 x 384 dims the kernel sits at the roofline ridge (t_HBM ~ t_MMA), so the tensor-side
    # fraction is reported beside the HBM one.
    flops = 2.0 * nq * N_ROWS * DIM
    kernel = "gemm_topk_kernel" if launches_per_step < 1.5 else "scan_topk_kernel"
    achieved_gbs = corpus_bytes * launches_per_step / (kern_ms_per_step * 1e-3) / 1e9 if kern_ms_per_step else 0.0
    tflops = flops / (kern_ms_per_step * 1e-3) / 1e12 if kern_ms_per_step else 0.0
    # The kernel is timed inside a long step (regions repeated for 0.5 s, the GPU under its power cap), so the tensor ceiling
    # that applies is cuBLAS's SUSTAINED figure.  Which roof binds is decided by the floors: bytes / HBM peak vs flops / tensor
    # peak -- at batch 256 x 384 dims the tensor floor (139.5 us) is above the HBM floor (119.3 us).
    t_hbm = corpus_bytes / (pk["hbm_gbs"] * 1e9)
    t_tensor = flops / (pk["bf16_tflops_sustained"] * 1e12)
    hbm_side = {"achieved": achieved_gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": achieved_gbs / pk["hbm_gbs"], "floor_us": t_hbm * 1e6}
    tensor_side = {"achieved": tflops, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": tflops / pk["bf16_tflops_sustained"],
                   "floor_us": t_tensor * 1e6, "peak_kind": "sustained (kernel timed inside a long step)",
                   "peak_burst": pk["bf16_tflops"], "frac_of_burst": tflops / pk["bf16_tflops"],
                   "counter_derived_utilisation": "profiles/r2_tensor_pipe.md: UTCHMMA instruction count x shape / elapsed SM cycles = 74 % at this shape (ncu, kernel alone)"}
    bound = "tensor" if t_tensor >= t_hbm else "hbm"
    top = tensor_side if bound == "tensor" else hbm_side
    line["roofline"] = {"bound": bound, "achieved": top["achieved"], "peak": top["peak"], "unit": top["unit"], "frac": top["frac"],
                        "traffic": ncu_traffic(kernel), "peak_src": pk["src"],
                        "kernel": kernel, "launches_per_step": launches_per_step, "kernel_ms_per_step": kern_ms_per_step,
                        "algorithmic_bytes_per_launch": corpus_bytes, "flops_per_step": flops,
                        "hbm": hbm_side, "tensor": tensor_side,
                        "note": "one launch per query batch: the kernel's time includes its in-kernel threshold seeding; both roofs are "
                                "reported, `bound` is the one whose floor is higher"}
    line["gpu_launches"] = gpu_launches
    line["clocks"] = clocks
    line["ingest"] = ingest
    if world == 1:
        line["batch1"] = batch1
        line["config0"] = config0
    if config3 is not None:
        line["config3"] = config3
    if config5 is not None:
        line["config5"] = config5
    if cpu is not None:
        line["cpu_baseline"] = cpu
    if sharded_c4 is not None:
        line["sharded_c4"] = sharded_c4
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def leg_config3(lib, _lib, dev, timed, summarize, kernel_time, hbm_roofline, pk, K):
    """BASELINE configs[2]: 1M x 512 mixed text+image collection with a metadata type filter, top_k=10, through the
    Chroma-shaped collection (metadata dicts in, `where=` on the query): the {"type": ...} mask and a general compiled
    clause, batch 1 and batch 256."""
    import numpy as np
    import torch
    from multimodal_rag_b200 import B200Collection
    rows, dim, k = 1_000_000, 512, 10
    rng = np.random.default_rng(0x7E57)
    types = rng.choice(np.array(["text", "table", "image"]), size=rows, p=[0.6, 0.1, 0.3])
    pages = rng.integers(0, 40, size=rows)
    c = B200Collection("config3", {"hnsw:space": "cosine"}, capacity=rows, device=dev.index)
    g = torch.Generator(device=dev).manual_seed(0xC3)
    t0 = time.perf_counter()
    for s in range(0, rows, 1 << 17):
        m = min(1 << 17, rows - s)
        x = torch.nn.functional.normalize(torch.randn(m, dim, generator=g, device=dev), dim=1)
        c.add(ids=[f"d{i}" for i in range(s, s + m)], embeddings=x,
              metadatas=[{"type": str(t), "page": int(p)} for t, p in zip(types[s:s + m], pages[s:s + m])])
    add_s = time.perf_counter() - t0
    st = torch.cuda.current_stream().cuda_stream
    out = {"workload": f"configs[2]: {rows}x{dim} bf16 corpus, types text 60% / table 10% / image 30%, top_k={k}, cosine",
           "add_rows_per_s_with_metadata": rows / add_s, "bars": {"batch1_qps": 5343, "batch256_qps": 0.97e6}, "filters": {}}
    filters = {"type_mask": {"type": "image"},
               "compiled_clause": {"$and": [{"type": {"$in": ["image", "table"]}}, {"page": {"$gte": 3}}]}}
    n_pass = {"type_mask": int((types == "image").sum()),
              "compiled_clause": int((((types == "image") | (types == "table")) & (pages >= 3)).sum())}
    for name, where in filters.items():
        f, keep = c.device_filter(where)
        leg = {"where": where, "rows_passing": n_pass[name]}
        for nq in (1, 256):
            gq = torch.Generator(device=dev).manual_seed(0xBEEF3 + nq)
            Q = [torch.nn.functional.normalize(torch.randn(nq, dim, generator=gq, device=dev), dim=1) for _ in range(4)]
            o_rows = torch.empty((nq, k), dtype=torch.int64, device=dev)
            o_dist = torch.empty((nq, k), dtype=torch.float32, device=dev)
            o_cnt = torch.empty((nq,), dtype=torch.int32, device=dev)

            def step(i):
                _lib.check(lib.b2r_query(c.handle, Q[i % 4].data_ptr(), nq, k, ctypes.byref(f), o_rows.data_ptr(),
                                         o_dist.data_ptr(), o_cnt.data_ptr(), st), "b2r_query")

            steps = max(K, 20)
            for i in range(5):
                step(i)
            s = summarize(timed(step, steps, min_s=0.25), steps)
            kern_ms, lps = kernel_time(c.handle, step, steps)
            assert int(o_cnt.min()) == k
            # every returned row must pass the filter (checked on the host tables)
            rr = o_rows.cpu().numpy().ravel()
            ok = (types[rr] == "image") if name == "type_mask" else (((types[rr] == "image") | (types[rr] == "table")) & (pages[rr] >= 3))
            assert bool(ok.all()), "a returned row does not pass the filter"
            # host-in / host-out through the collection's own query_rows (numpy in, numpy out)
            Qn = [q.cpu().numpy() for q in Q]
            c.query_rows(Qn[0], k, where)
            t0 = time.perf_counter()
            reps = 0
            while time.perf_counter() - t0 < 0.2:
                c.query_rows(Qn[reps % 4], k, where)
                reps += 1
            host_ms = (time.perf_counter() - t0) / reps * 1e3
            leg[f"batch{nq}"] = {"qps": nq / (s["ms_per_step"] * 1e-3), **s,
                                 "e2e_qps": nq / (host_ms * 1e-3), "e2e_api": "B200Collection.query_rows(numpy, k, where) -> numpy",
                                 "roofline": hbm_roofline("gemm_topk_kernel", rows * dim * 2, kern_ms / max(lps, 1e-9),
                                                          {"tensor_tflops": 2.0 * nq * rows * dim / (kern_ms / max(lps, 1e-9) * 1e-3) / 1e12})}
        del keep
        out["filters"][name] = leg
    c.close()
    return out


def leg_config5(lib, _lib, dev, timed, summarize, hbm_roofline, pk, K):
    """BASELINE configs[4]: 10M x 768 streaming workload -- a pre-reserved shard, then the loop {upsert 8192 rows, ~10 % of
    them overwriting existing rows (tombstone + append through K1); batch-64 query, top_k=20}.  The visibility check: the
    first queries of every batch are rows of the batch just upserted, so each must come back as its own nearest
    neighbour (distance ~0) -- an add is visible to the next query on the stream."""
    import numpy as np
    import torch
    from multimodal_rag_b200.sharded import DeviceShard
    rows0, dim, k, nq, up = 10_000_000, 768, 20, 64, 8192
    free_b, _ = torch.cuda.mem_get_info()
    need = (rows0 + 600 * up) * (dim * 6 + 8)
    if free_b < need * 1.05:
        return {"skipped": f"needs {need / 1e9:.0f} GB of free HBM, {free_b / 1e9:.0f} GB available"}
    cap = rows0 + 600 * up
    sh = DeviceShard(dim, "cosine", capacity=cap, row_base=0, device=dev.index, world=1)
    g = torch.Generator(device=dev).manual_seed(0xC5)
    ing = []
    for s in range(0, rows0, 1 << 18):
        m = min(1 << 18, rows0 - s)
        x = torch.nn.functional.normalize(torch.randn(m, dim, generator=g, device=dev), dim=1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); sh.ingest(x); e1.record(); e1.synchronize()
        ing.append((m, e0.elapsed_time(e1)))
    full = [(m, t) for m, t in ing[1:] if m == 1 << 18]
    ing_rows, ing_ms = sum(m for m, _ in full), sum(t for _, t in full)
    bytes_per_row = dim * 4 + dim * 2 + dim * 4 + 1
    st = torch.cuda.current_stream().cuda_stream
    rng = np.random.default_rng(5)
    o = sh.alloc_out(nq, k)
    state = {"rows": rows0, "visible": True, "steps": 0}
    # upsert batches and query batches are generated on the device BEFORE each timed region (every batch is new data:
    # a rolled copy of a base batch), so the region holds only the path: tombstone + K1 + query
    n_pre = 8
    base = [torch.nn.functional.normalize(torch.randn(up, dim, generator=g, device=dev), dim=1) for _ in range(n_pre)]
    gq = torch.Generator(device=dev).manual_seed(0xBEEF5)
    Qbase = [torch.nn.functional.normalize(torch.randn(nq, dim, generator=gq, device=dev), dim=1) for _ in range(n_pre)]
    n_over = up // 10
    made = {"n": 0}

    def make_batches(count):
        ups, Qs = [], []
        for _ in range(count):
            made["n"] += 1
            u = torch.roll(base[made["n"] % n_pre], shifts=made["n"], dims=1).contiguous()
            q = Qbase[made["n"] % n_pre].clone()
            q[:8] = u[:8]                                  # visibility probes: 8 rows of the batch about to be upserted
            ups.append(u); Qs.append(q)
        torch.cuda.synchronize()
        return ups, Qs

    def step(u, q):
        # ~10 % of the batch overwrites existing rows: tombstone them (the vector half of upsert), then append all 8192
        old = rng.integers(0, rows0, size=n_over, dtype=np.int64)
        _lib.check(lib.b2r_tombstone(sh.h, old.ctypes.data, n_over, st), "b2r_tombstone")
        first = sh.ingest(u)
        sh.query_local(q, k, o)
        state["rows"] = first + up
        state["last_first"] = first
        state["steps"] += 1

    def check_visible():
        torch.cuda.synchronize()
        top = o["rows"][:8, 0].cpu().numpy()
        d0 = o["dist"][:8, 0].cpu().numpy()
        want = state["last_first"] + np.arange(8)
        return bool((top == want).all() and (np.abs(d0) < 1e-5).all())

    ups, Qs = make_batches(3)
    for i in range(3):
        step(ups[i], Qs[i])
        assert check_visible(), "an upserted row is not its own nearest neighbour in the next query"
    steps = max(K, 10)
    regions = []
    total = 0.0
    while total < 500.0 and state["steps"] + steps < 590:
        ups, Qs = make_batches(steps)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for i in range(steps):
            step(ups[i], Qs[i])
        e1.record()
        torch.cuda.synchronize()
        regions.append(e0.elapsed_time(e1))
        total += regions[-1]
        state["visible"] = state["visible"] and check_visible()
    Qs = Qs[:n_pre]
    s = summarize(regions, steps)
    # the query kernel alone, on the grown shard
    tot, cnt = ctypes.c_double(), ctypes.c_int64()
    _lib.check(lib.b2r_set_kernel_timing(sh.h, 1))
    _lib.check(lib.b2r_kernel_time_ms(sh.h, ctypes.byref(tot), ctypes.byref(cnt), 1))
    for i in range(10):
        sh.query_local(Qs[i % len(Qs)], k, o)
    torch.cuda.synchronize()
    _lib.check(lib.b2r_kernel_time_ms(sh.h, ctypes.byref(tot), ctypes.byref(cnt), 1))
    _lib.check(lib.b2r_set_kernel_timing(sh.h, 0))
    kern_ms = tot.value / max(1, cnt.value)
    rows_now = state["rows"]
    out = {"workload": f"configs[4]: {rows0}x{dim} streaming: per step tombstone {n_over} + upsert {up} rows (K1) + batch-{nq} query, top_k={k}",
           "qps": nq / (s["ms_per_step"] * 1e-3), "upsert_rows_per_s_in_loop": up / (s["ms_per_step"] * 1e-3), **s,
           "rows_at_end": rows_now, "visibility_check": state["visible"], "fallbacks": sh.fallbacks(),
           "bars": {"qps": 22800, "ingest_rows_per_s_roofline": pk["hbm_gbs"] * 1e9 / bytes_per_row},
           "roofline": hbm_roofline("gemm_topk_kernel", rows_now * dim * 2, kern_ms),
           "ingest": {"rows_per_s": ing_rows / (ing_ms * 1e-3), "algorithmic_bytes_per_row": bytes_per_row,
                      "roofline": {"bound": "hbm", "achieved": ing_rows * bytes_per_row / (ing_ms * 1e-3) / 1e9, "peak": pk["hbm_gbs"],
                                   "unit": "GB/s", "frac": ing_rows * bytes_per_row / (ing_ms * 1e-3) / 1e9 / pk["hbm_gbs"]}}}
    assert state["visible"], "an upserted row is not its own nearest neighbour in the next query"
    sh.close()
    out["collection_api"] = leg_config5_collection(dev, rows0, dim, k, nq, up, out["ms_per_step"])
    return out


def leg_config5_collection(dev, rows0, dim, k, nq, up, device_ms):
    """The same streaming loop through the Chroma-shaped collection (string ids, host id tables): B200Collection.upsert
    (~10 % of each batch overwrites existing ids) + B200Collection.query_rows, device-resident vectors in, numpy results
    out.  Reported beside the raw-ABI number: what the host tables cost."""
    import numpy as np
    import torch
    from multimodal_rag_b200 import B200Collection
    steps = 24
    c = B200Collection("config5", {"hnsw:space": "cosine"}, capacity=rows0 + (steps + 4) * up, device=dev.index)
    g = torch.Generator(device=dev).manual_seed(0xC5)
    t0 = time.perf_counter()
    for s in range(0, rows0, 1 << 18):
        m = min(1 << 18, rows0 - s)
        c.add(ids=[f"r{i}" for i in range(s, s + m)],
              embeddings=torch.nn.functional.normalize(torch.randn(m, dim, generator=g, device=dev), dim=1))
    load_s = time.perf_counter() - t0
    rng = np.random.default_rng(55)
    base = torch.nn.functional.normalize(torch.randn(up, dim, generator=g, device=dev), dim=1)
    Qb = torch.nn.functional.normalize(torch.randn(nq, dim, generator=g, device=dev), dim=1)
    n_over = up // 10
    batches = []
    for j in range(steps + 2):
        u = torch.roll(base, shifts=j + 1, dims=1).contiguous()
        q = Qb.clone(); q[:8] = u[:8]
        ids = [f"n{j}_{i}" for i in range(up)]
        for t, o in enumerate(rng.integers(0, rows0, size=n_over).tolist()):
            ids[up - 1 - t] = f"r{o}"                      # overwrite an existing id
        ids = list(dict.fromkeys(ids))
        batches.append((ids, u[: len(ids)], q))
    torch.cuda.synchronize()
    ok = True

    def step(j):
        ids, u, q = batches[j]
        c.upsert(ids=ids, embeddings=u)
        rows, dist, cnt = c.query_rows(q, k)
        return rows, dist

    for j in range(2):
        rows, dist = step(j)
    t0 = time.perf_counter()
    for j in range(2, steps + 2):
        rows, dist = step(j)
        ok = ok and c.ids_of(rows[:8, 0]) == batches[j][0][:8] and bool((np.abs(dist[:8, 0]) < 1e-5).all())
    ms = (time.perf_counter() - t0) / steps * 1e3
    n_live = c.count()
    c.close()
    return {"api": "B200Collection.upsert(ids, device tensor) + B200Collection.query_rows(device tensor, k) -> numpy",
            "qps": nq / (ms * 1e-3), "ms_per_step": ms, "steps": steps, "visibility_check": ok, "live_rows_at_end": n_live,
            "bulk_add_rows_per_s": rows0 / load_s, "slowdown_vs_raw_abi": ms / device_ms}


def leg_config4(args, lib, _lib, dev, dist, rank, world, timed, summarize, sum_over_ranks, pk, K):
    """BASELINE configs[3]: 100M x 384 row-sharded over the N GPUs, batch 1024, top_k 100 (strong scaling of a fixed corpus)."""
    import torch
    from multimodal_rag_b200.sharded import DeviceShard
    rows_total = args.c4_rows
    per = rows_total // world
    free_b, _ = torch.cuda.mem_get_info()
    bytes_per_row = DIM * 2 + DIM * 4 + 8
    scaled = False
    if per * bytes_per_row > 0.8 * free_b:                 # does not fit beside the fp32 master: say so
        per = int(0.8 * free_b / bytes_per_row) // 4096 * 4096
        scaled = True
    t = torch.tensor([per], dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    per = int(t.item())
    sh3 = DeviceShard(DIM, "cosine", capacity=per, row_base=rank * per, device=dev.index)
    g = torch.Generator(device=dev).manual_seed(0xC4 + rank)
    for s0 in range(0, per, 1 << 18):
        m = min(1 << 18, per - s0)
        sh3.ingest(torch.nn.functional.normalize(torch.randn(m, DIM, generator=g, device=dev), dim=1))
    torch.cuda.synchronize()
    nq4, k4 = 1024, 100
    gq4 = torch.Generator(device=dev).manual_seed(0xBEEF4)     # replicated queries
    Q4 = [torch.nn.functional.normalize(torch.randn(nq4, DIM, generator=gq4, device=dev), dim=1) for _ in range(2)]
    o4 = sh3.alloc_out(nq4, k4)

    def step_c4(i):
        sh3.query_device(Q4[i % 2], k4, o4)

    for i in range(2):
        step_c4(i)
    K4 = max(3, min(K, 5))
    s = summarize(timed(step_c4, K4), K4)
    per_rank_ms = [None] * world
    dist.all_gather_object(per_rank_ms, s["ms_per_step"])
    fallbacks = int(sum_over_ranks(sh3.fallbacks()))
    # self-check: the same exchange + merge fed by the exact fp64 scan (K5) must give the same rows for a sample of the batch
    ns = 8
    sh3.query_device(Q4[0], k4, o4)
    fast_rows = o4["m_rows"][:ns].clone()
    fast_dist = o4["m_dist"][:ns].clone()
    _lib.check(lib.b2r_set_path(sh3.h, 3))
    o5 = sh3.alloc_out(ns, k4)
    sh3.query_device(Q4[0][:ns].contiguous(), k4, o5)
    torch.cuda.synchronize()
    _lib.check(lib.b2r_set_path(sh3.h, 0))
    ok = bool(torch.equal(fast_rows, o5["m_rows"])) and bool(torch.allclose(fast_dist, o5["m_dist"], rtol=1e-6, atol=0))
    vflag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(vflag, op=dist.ReduceOp.MIN)
    if int(vflag.item()) != 1:
        raise SystemExit("bench: config-4 merged result differs from the exact fp64 scan on the sampled queries")
    flops4 = 2.0 * nq4 * per * DIM
    tf = flops4 / (s["ms_per_step"] * 1e-3) / 1e12
    out = {"workload": f"configs[3]: {per * world} x {DIM} bf16 rows row-sharded over {world} GPUs, batch {nq4}, top_k {k4}",
           "rows_total": per * world, "rows_per_gpu": per, "scaled_down_to_fit": scaled,
           "qps": nq4 / (s["ms_per_step"] * 1e-3), **s, "ms_per_step_per_rank": per_rank_ms,
           "tflops_per_gpu": tf,
           "roofline": {"bound": "tensor", "achieved": tf, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                        "frac": tf / pk["bf16_tflops_sustained"], "peak_kind": "sustained (kernel timed inside a long step)",
                        "frac_of_burst": tf / pk["bf16_tflops"]},
           "roofline_qps": {"burst": nq4 / (flops4 / (pk["bf16_tflops"] * 1e12)), "sustained": nq4 / (flops4 / (pk["bf16_tflops_sustained"] * 1e12))},
           "collective": "ONE nccl all_gather of the packed per-rank top-k block",
           "bytes_gathered_per_step": world * o4["layout"][0],
           "exact_fallbacks_all_ranks": fallbacks,
           "verified": True, "verified_how": f"merged rows/distances of {ns} sampled queries == the same exchange fed by the exact fp64 scan (K5) on every shard"}
    sh3.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-hnsw", action="store_true", help="skip the HNSW restatement inside the CPU legs")
    ap.add_argument("--no-c4", action="store_true", help="N > 1: skip the 100M-row config-4 leg")
    ap.add_argument("--no-configs", action="store_true", help="N = 1: skip the config3 / config5 legs")
    ap.add_argument("--p2p-exchange", action="store_true", help="N > 1: the library's own exchange (b2r_xchg_push / b2r_xchg_merge: peer-to-peer stores + "
                    "stream memory operations) instead of the NCCL all_gather; measured slower at the headline shape, see profiles/r2_exchange.md")
    ap.add_argument("--c4-rows", type=int, default=100_000_000, help="total rows of the config-4 leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
