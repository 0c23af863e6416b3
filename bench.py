#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on its config: queries/sec, 10M-doc Zipfian corpus, 3-term AND,
BM25 top-100 (config C2), at 1/2/4/8 B200.

    python bench.py --gpus N --steps K --warmup W            our CUDA path (torchrun for N>1)
    python bench.py --impl reference --gpus N --steps K ...  the reference's own CPU Enquire::get_mset

A "step" is one pass of the hot path over one batch of BATCH synthetic queries.
  value  — whole-job queries/s with the batch's plan already resident in HBM (device-timed with CUDA
           events on the searcher's stream, barrier + synchronize on both sides, max over ranks).
  e2e    — the same metric through the reference-facing C-ABI call (xgm_search_submit/wait) with HOST
           buffers: host planning, H2D of the plan, kernels, D2H of the MSets all inside the timed region.
  roofline — decode+intersect+score kernel: algorithmic bytes / its CUDA-event time.
  cpu_baseline — the compiled reference (oracle/_ref) on the box's host cores, rank 0, N=1 only.
N>1 is strong scaling: the same 10M-doc corpus split into N interleaved docid shards (Xapian's own
scheme, backends/multi.h:37-70), every query runs on every shard with global statistics, and the
per-GPU top-k are merged after a single NCCL all-gather (Matcher::merge_mset semantics).
"""
from __future__ import annotations

import argparse
import json
import os
import random
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NDOCS = int(os.environ.get("XGM_BENCH_DOCS", 10_000_000))
VOCAB = int(os.environ.get("XGM_BENCH_VOCAB", 1_000_000))
SEED = 12345
QSEED = 777
TOPRANKS = 1000
NTERMS = 3
TOPK = 100
BATCH = int(os.environ.get("XGM_BENCH_BATCH", 4096))
REF_QUERIES_PER_STEP = int(os.environ.get("XGM_BENCH_REF_QUERIES", 1024))
METRIC = "queries/sec, 10M-doc 3-term AND BM25 top-100"
UNIT = "queries/s"


def gen_query_terms(step: int, n: int):
    rng = random.Random(QSEED * 1000003 + step)
    return [rng.sample(range(TOPRANKS), NTERMS) for _ in range(n)]


def term_name(r: int) -> str:
    return f"T{r:06d}"


# ---------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe)
# ---------------------------------------------------------------------------------------------

class ClockSampler:
    """SM clock / throttle-reason samples taken DURING the timed regions.  NVML is polled from a thread every
    few milliseconds (the regions last tens of milliseconds; polling faster contends with the CUDA driver
    and slows the host side of the end-to-end loop), plus one sample taken by the main thread while the
    device-resident region's work is queued; `nvidia-smi -lms` is the fallback."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"),
               (0x4, "sw_power_cap"), (0x80, "hw_power_brake_slowdown"))

    def __init__(self, gpu_index: int, uuid: str | None = None):
        self.gpu = gpu_index
        self.uuid = uuid
        self.samples = []      # (perf_counter, sm_mhz, reason_bits)
        self.windows = []      # [t0, t1] of the timed regions
        self.max_mhz = None
        self.proc = None
        self.rows = []
        self.handle = None
        self.stop_flag = False
        self.thread = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = None
            if self.uuid:
                try:
                    h = pynvml.nvmlDeviceGetHandleByUUID(self.uuid)
                except Exception:
                    h = None
            if h is None:
                h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
            self.handle = h
            self.nvml = pynvml
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.handle = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "20"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
            time.sleep(0.3)
        except Exception:
            self.proc = None

    def sample_now(self):
        if self.handle is None:
            return
        n = self.nvml
        try:
            mhz = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
            try:
                bits = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
            except Exception:
                bits = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
            self.samples.append((time.perf_counter(), mhz, bits))
        except Exception:
            pass

    def _poll(self):
        while not self.stop_flag:
            self.sample_now()
            time.sleep(0.004)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def begin(self):
        self.windows.append([time.perf_counter(), None])

    def end(self):
        self.windows[-1][1] = time.perf_counter()

    def _inside(self, t):
        return any(w[0] <= t <= (w[1] if w[1] is not None else t) for w in self.windows)

    def stop(self):
        self.stop_flag = True
        if self.handle is None and not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["nvml and nvidia-smi unavailable"]}
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
            mx = []
            for t, r in self.rows:
                p = [x.strip() for x in r.split(",")]
                if len(p) < 9:
                    continue
                try:
                    mhz = float(p[1]); mx.append(float(p[2]))
                except ValueError:
                    continue
                bits = 0
                for (bit, _), v in zip(self.REASONS[:4], p[5:9]):
                    if v.lower().startswith("active"):
                        bits |= bit
                self.samples.append((t, mhz, bits))
            self.max_mhz = max(mx) if mx else None
        elif self.thread:
            self.thread.join(timeout=1)
        inside = [s for s in self.samples if self._inside(s[0])]
        # the regions are short; if the poller never fell inside one, use the samples bracketing them
        use = inside or self.samples
        bits = 0
        for s in use:
            bits |= s[2]
        reasons = sorted(name for bit, name in self.REASONS if bits & bit)
        return {"sm_mhz": statistics.median(s[1] for s in use) if use else None, "sm_max_mhz": self.max_mhz,
                "samples": len(use), "samples_in_timed_regions": len(inside), "reasons": reasons,
                "source": "nvml" if self.handle is not None else "nvidia-smi"}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the compiled reference's Enquire::get_mset on host cores
# ---------------------------------------------------------------------------------------------

def ref_db_dir():
    base = os.environ.get("XGM_REF_DB_DIR") or ("/dev/shm" if os.path.isdir("/dev/shm") else "/tmp")
    return os.path.join(base, f"xgm_refdb_{NDOCS}_{VOCAB}_{SEED}")


def ref_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def build_reference_db():
    from oracle import oracle as O
    if not O.have_reference():
        raise RuntimeError("oracle/_ref missing: the reference was not built (oracle/build_ref.sh)")
    procs = min(ref_cores(), 128)
    return O.ref_build_parallel(ref_db_dir(), NDOCS, VOCAB, seed=SEED, procs=procs)


def run_reference_queries(dbs, nqueries: int, steps: int, warmup: int, threads: int):
    """Each step = nqueries queries of the bench workload, all host threads, timing only get_mset."""
    from oracle import oracle as O
    work = os.path.join(ref_db_dir(), "work")
    terms = gen_query_terms(0, nqueries)
    lines = [O.query_line("AND", [term_name(t) for t in q], 0, TOPK) for q in terms]
    info, _ = O.ref_query(dbs, lines, work, threads=threads, repeat=steps, warmup=max(1, warmup), dump=False)
    return info


def single_thread_baseline(dbs):
    """SURVEY.md §8(d) asks for the reference at (i) one thread and (ii) all cores: the one-thread leg, on a
    small bounded sample (a few seconds)."""
    n = int(os.environ.get("XGM_BENCH_REF_QUERIES_1T", 256))
    try:
        info = run_reference_queries(dbs, n, 1, 1, 1)
        return {"value": info["qps"], "unit": UNIT, "cores": 1, "p50_ms": info["p50_ms"], "p99_ms": info["p99_ms"],
                "sample": f"1 pass of {n} queries of the same workload, one thread"}
    except Exception as e:  # reported, never required
        return {"value": None, "unit": UNIT, "cores": 1, "sample": f"unavailable: {e}"}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    t0 = time.time()
    binfo = build_reference_db()
    cores = ref_cores()
    info = run_reference_queries(binfo["dbs"], REF_QUERIES_PER_STEP, args.steps, args.warmup, cores)
    qps = info["qps"]
    ms_per_step = info["wall_s"] / args.steps * 1e3
    sample = (f"{REF_QUERIES_PER_STEP} queries/step of the same workload on the full {NDOCS}-doc glass DB "
              f"(built by {binfo['procs']} parallel WritableDatabase writers + Database::compact), "
              f"{cores} threads each with its own Xapian::Database+Enquire, timing Enquire::get_mset only")
    line = {"impl": "reference", "metric": METRIC, "value": qps, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C2: 10M docs, V=1M Zipf(1) terms, 3-term OP_AND, BM25, get_mset(0,100)",
                       "docs": NDOCS, "vocab": VOCAB, "queries_per_step": REF_QUERIES_PER_STEP, "topk": TOPK},
            "cpu_baseline": {"value": qps, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample,
                             "p50_ms": info["p50_ms"], "p99_ms": info["p99_ms"],
                             "single_thread": single_thread_baseline(binfo["dbs"])},
            "e2e": {"value": qps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "setup_s": round(time.time() - t0, 1)}
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------

class CudaArray:
    """Zero-copy view of a raw device pointer for torch.as_tensor (CUDA array interface v2)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2}


def ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from xapiand_b200 import xgm

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libxgm has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    t_setup = time.time()
    ix = xgm.Index.synthetic(NDOCS, VOCAB, seed=SEED, nshards=world, shard=rank, device=local_rank)
    info = ix.info()
    build_s = time.time() - t_setup

    # ---- global statistics (phase 1 of Xapiand's two-phase scheme, handler.cc:1532-1538) ----
    def make_batch(step: int):
        terms = gen_query_terms(step, BATCH)
        stats_list = None
        if world > 1:
            uniq = sorted({t for q in terms for t in q})
            local_tf = torch.tensor([ix.term_stats(term_name(t)).termfreq for t in uniq], dtype=torch.int64, device="cuda")
            totals = torch.tensor([info.doccount, info.total_length], dtype=torch.int64, device="cuda")
            dist.all_reduce(local_tf)
            dist.all_reduce(totals)
            gtf = dict(zip(uniq, local_tf.tolist()))
            coll, tlen = totals.tolist()
            stats_list = [(coll, tlen, [gtf[t] for t in q]) for q in terms]
        qs = [xgm.Query(xgm.OP_AND, [term_name(t) for t in q], first=0, maxitems=TOPK,
                        stats=None if stats_list is None else stats_list[i]) for i, q in enumerate(terms)]
        return xgm.QueryBatch(qs)

    NSEARCH = 3  # batches in flight in the end-to-end loop (host planning / GPU / result scatter overlap)
    searchers = [xgm.Searcher(ix, max_batch=BATCH, max_topk=TOPK) for _ in range(NSEARCH)]
    streams = [torch.cuda.ExternalStream(s.stream(), device=torch.device("cuda", local_rank)) for s in searchers]
    L = xgm.lib()

    # merge buffers for N > 1: ONE all-gather of each GPU's result slab (weights | docids | counts of its
    # per-query top-k), then the merge kernel (Matcher::merge_mset) on every rank
    if world > 1:
        gathered = []
        for s in searchers:
            base, nbytes, off_d, off_c, stride = s.device_slab()
            local = torch.as_tensor(CudaArray(base, (nbytes,), "|u1"), device="cuda")
            g = torch.empty(world * nbytes, dtype=torch.uint8, device="cuda")
            ow = torch.empty(BATCH * TOPK, dtype=torch.float64, device="cuda")
            od = torch.empty(BATCH * TOPK, dtype=torch.int32, device="cuda")
            on = torch.empty(BATCH, dtype=torch.int32, device="cuda")
            gathered.append((local, g, nbytes, off_d, off_c, ow, od, on, stride))
        host_out = [(torch.empty(BATCH * TOPK, dtype=torch.float64).pin_memory(),
                     torch.empty(BATCH * TOPK, dtype=torch.int32).pin_memory(),
                     torch.empty(BATCH, dtype=torch.int32).pin_memory()) for _ in searchers]

    def merge_step(si: int, to_host: bool):
        """all-gather + merge on the searcher's stream; optionally copy the merged MSets to the host."""
        local, g, nbytes, off_d, off_c, ow, od, on, stride = gathered[si]
        with torch.cuda.stream(streams[si]):
            dist.all_gather_into_tensor(g, local)
            st = L.xgm_merge_topk_device_slab(g.data_ptr(), nbytes, off_d, off_c, world, BATCH, stride, TOPK,
                                              ow.data_ptr(), od.data_ptr(), on.data_ptr(), searchers[si].stream())
            if st != 0:
                raise RuntimeError(L.xgm_last_error().decode())
            if to_host:
                hw, hd, hn = host_out[si]
                hw.copy_(ow, non_blocking=True)
                hd.copy_(od, non_blocking=True)
                hn.copy_(on, non_blocking=True)

    # ---- warm-up: W steps through the full API (also makes the plan of batch 0 resident) ----
    batches = [make_batch(i) for i in range(max(args.warmup, 1) + args.steps + 1)]
    for w in range(max(args.warmup, 3)):
        for si, srch in enumerate(searchers):  # every searcher (staging buffers, worker thread) is warmed up
            srch.submit(batches[w % len(batches)], background=True)
            if world > 1:
                srch.launched()
                merge_step(si, True)
            srch.wait_raw()
    if world > 1:  # the device-resident loop alternates between two searchers holding the same resident plan
        searchers[1].submit(batches[0])
        searchers[1].wait_raw()
    searchers[0].submit(batches[0])
    _, _, _, inf0 = searchers[0].wait_raw()
    bad = sum(1 for i in range(BATCH) if inf0[i].status != 0)
    if bad:
        raise SystemExit(f"{bad} queries of the bench batch were not answered on the device")
    st0 = searchers[0].last_stats()
    barrier()

    # ---- device-resident timed region: K replays of the resident plan ----
    try:
        uuid = "GPU-" + str(torch.cuda.get_device_properties(local_rank).uuid)
    except Exception:
        uuid = None
    sampler = ClockSampler(local_rank, uuid)
    sampler.start()
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    match_ms = []
    barrier()
    sampler.begin()
    ev0.record(streams[0])
    for k in range(args.steps):
        # N > 1: the all-gather + merge of step k (searcher k%2's stream) overlaps the kernels of step k+1
        # (the index's compute stream), consecutive steps being independent batches
        si = k % 2 if world > 1 else 0
        searchers[si].replay()
        if world > 1:
            merge_step(si, False)
    if world > 1:
        tail = torch.cuda.Event()
        tail.record(streams[1])
        streams[0].wait_event(tail)
    ev1.record(streams[0])
    sampler.sample_now()  # the K replays are queued and running
    barrier()
    sampler.end()
    dev_ms = ev0.elapsed_time(ev1)
    # per-launch time of the dominant kernel: K more replays, reading each launch's own events
    for k in range(args.steps):
        searchers[0].replay()
        match_ms.append(searchers[0].last_stats().match_kernel_ms)
    torch.cuda.synchronize()
    t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    value = BATCH * args.steps / (dev_ms / 1e3)

    # ---- end-to-end timed region: K steps through submit/wait with host buffers, NSEARCH searchers ----
    barrier()
    sampler.begin()
    t0 = time.perf_counter()
    inflight = []
    h2d = d2h = 0
    prev = None
    for k in range(args.steps):
        si = k % NSEARCH
        if len(inflight) == NSEARCH:
            searchers[inflight.pop(0)].wait_raw()
        # the searcher's worker thread plans + enqueues batch k while this thread scatters an earlier one
        searchers[si].submit(batches[1 + k], background=True)
        if world > 1 and prev is not None:
            # all-gather + merge of the previous batch: its kernels were enqueued while we were busy above
            searchers[prev].launched()
            merge_step(prev, True)
        prev = si
        inflight.append(si)
    if world > 1:
        searchers[prev].launched()
        merge_step(prev, True)
    pending = inflight[-1]
    for si in inflight:
        searchers[si].wait_raw()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    sampler.end()
    e2e_s = t1 - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    bs = searchers[pending].last_stats()
    h2d, d2h = int(bs.h2d_bytes), int(bs.d2h_bytes)
    if world > 1:
        d2h += BATCH * TOPK * 12 + BATCH * 4
    clocks = sampler.stop()
    e2e_value = BATCH * args.steps / e2e_s

    # ---- p50 latency at batch = 1 through the C-ABI (rank-local) ----
    lat = []
    one = xgm.Searcher(ix, max_batch=1, max_topk=TOPK)
    singles = [xgm.QueryBatch([xgm.Query(xgm.OP_AND, [term_name(t) for t in q], maxitems=TOPK)])
               for q in gen_query_terms(999, 200)]
    for b in singles[:20]:
        one.submit(b); one.wait_raw()
    for b in singles:
        a = time.perf_counter()
        one.submit(b)
        one.wait_raw()
        lat.append((time.perf_counter() - a) * 1e3)
    lat.sort()

    # ---- roofline of the dominant kernel ----
    peak, peak_src = measured_peak_gbs()
    alg_bytes = int(st0.algorithmic_bytes)
    kern_ms = statistics.mean(match_ms)
    achieved = alg_bytes / 1e9 / (kern_ms / 1e3)
    launches_per_step = int(st0.kernel_launches) + (1 if world > 1 else 0)

    traffic = None
    tp = os.path.join(ROOT, "profiles", "r1_and_bm_kernel.json")
    if os.path.exists(tp) and world == 1:
        try:
            prof = json.load(open(tp))
            if prof.get("queries_per_launch") == BATCH:
                traffic = prof["dram_bytes_read"] + prof["dram_bytes_write"]
        except Exception:
            pass
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C2: 10M docs, V=1M Zipf(1) terms, 3-term OP_AND, BM25, get_mset(0,100)",
                       "docs": NDOCS, "vocab": VOCAB, "queries_per_step": BATCH, "topk": TOPK,
                       "shards": world, "shard_docs": int(info.doccount),
                       "cache": "inputs larger than L2: one step streams %.0f MB of posting columns (L2 = 126 MB)" % (alg_bytes / 1e6),
                       "index_bytes": int(info.bytes_docids + info.bytes_wdfs + info.bytes_headers + info.bytes_doclen),
                       "index_build_s": round(build_s, 1)},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_s / args.steps * 1e3, "p50_ms_batch1": lat[len(lat) // 2],
                    "p99_ms_batch1": lat[int(len(lat) * 0.99)]},
            "gpu_launches": launches_per_step * args.steps,
            "roofline": {"bound": "hbm", "kernel": "xgm_and_bm_kernel (decode driver + bitmap intersect + BM25)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "traffic_source": "profiles/r1_and_bm_kernel.json (ncu dram__bytes_read+write of one launch)" if traffic else None,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
                         "kernel_ms_per_launch": kern_ms},
            "clocks": clocks}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            binfo = build_reference_db()
            cores = ref_cores()
            cinfo = run_reference_queries(binfo["dbs"], REF_QUERIES_PER_STEP, 3, 1, cores)
            line["cpu_baseline"] = {
                "value": cinfo["qps"], "unit": UNIT, "cores": cores, "kind": "reference",
                "sample": (f"3 passes of {REF_QUERIES_PER_STEP} queries of the same workload on the full {NDOCS}-doc "
                           f"glass DB, {cores} threads (one Xapian::Database+Enquire each), timing get_mset only"),
                "p50_ms": cinfo["p50_ms"], "p99_ms": cinfo["p99_ms"],
                "single_thread": single_thread_baseline(binfo["dbs"])}
        except Exception as e:  # the baseline is reported, never required for the GPU numbers
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": ref_cores(), "kind": "reference",
                                    "sample": f"unavailable: {e}"}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    return ours(args)


if __name__ == "__main__":
    sys.exit(main())
