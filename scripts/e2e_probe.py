"""Scratch probe of the end-to-end loop (bench.py's e2e region at N=1): where the host time goes."""
import json
import os
import random
import sys
import time

sys.path.insert(0, ".")
import torch
from xapiand_b200 import xgm

NDOCS, VOCAB, BATCH, TOPK = 10_000_000, 1_000_000, 4096, 100
ix = xgm.Index.synthetic(NDOCS, VOCAB, seed=1)
rng = random.Random(5)
batches = [xgm.QueryBatch([xgm.Query(xgm.OP_AND, [f"T{t:06d}" for t in rng.sample(range(1000), 3)], maxitems=TOPK)
                           for _ in range(BATCH)]) for _ in range(24)]
for nsearch, threads in [(3, 8), (2, 8), (4, 8), (3, 4), (3, 16), (3, 6)]:
    os.environ["XGM_HOST_THREADS"] = str(threads)
    ss = [xgm.Searcher(ix, max_batch=BATCH, max_topk=TOPK) for _ in range(nsearch)]
    for s in ss:
        s.submit(batches[0]); s.wait_raw()
    for rep in range(2):
        torch.cuda.synchronize()
        t_sub = t_wait = 0.0
        plan = waitms = 0.0
        t0 = time.perf_counter()
        inflight = []
        K = 20
        for k in range(K):
            si = k % nsearch
            if len(inflight) == nsearch:
                a = time.perf_counter()
                j = inflight.pop(0)
                ss[j].wait_raw()
                t_wait += time.perf_counter() - a
                st = ss[j].last_stats(); plan += st.host_plan_ms; waitms += st.host_wait_ms
            a = time.perf_counter()
            ss[si].submit(batches[1 + k], background=True)
            t_sub += time.perf_counter() - a
            inflight.append(si)
        for j in inflight:
            ss[j].wait_raw()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    n = K - nsearch
    print(json.dumps(dict(nsearch=nsearch, threads=threads, ms_per_step=round(dt / K * 1e3, 3), qps=round(BATCH * K / dt),
                          submit_ms=round(t_sub / K * 1e3, 3), wait_call_ms=round(t_wait / n * 1e3, 3),
                          plan_ms=round(plan / n, 3), scatter_ms=round(waitms / n, 3))))
    del ss
