"""Scratch probe: GPU-side span of each batch (H2D + kernels + D2H) when several searchers are in flight."""
import json
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
for nsearch in (1, 3):
    ss = [xgm.Searcher(ix, max_batch=BATCH, max_topk=TOPK) for _ in range(nsearch)]
    st = [torch.cuda.ExternalStream(s.stream()) for s in ss]
    for s in ss:
        s.submit(batches[0]); s.wait_raw()
    K = 18
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    torch.cuda.synchronize()
    inflight = []
    kern = []
    t0 = time.perf_counter()
    for k in range(K):
        si = k % nsearch
        if len(inflight) == nsearch:
            j = inflight.pop(0)
            ss[j].wait_raw()
            x = ss[j].last_stats(); kern.append((x.match_kernel_ms, x.topk_kernel_ms))
        ev[k][0].record(st[si])
        ss[si].submit(batches[1 + k])
        ev[k][1].record(st[si])
        inflight.append(si)
    for j in inflight:
        ss[j].wait_raw()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    spans = [ev[k][0].elapsed_time(ev[k][1]) for k in range(K)]
    ends = [ev[0][0].elapsed_time(ev[k][1]) for k in range(K)]
    gaps = [ends[k] - ends[k - 1] for k in range(1, K)]
    print(json.dumps(dict(nsearch=nsearch, ms_per_step=round(dt / K * 1e3, 3), span_ms=[round(x, 2) for x in spans[4:12]],
                          end_gaps=[round(x, 2) for x in gaps[4:12]], kern=[(round(a, 2), round(b, 2)) for a, b in kern[4:10]])))
    del ss
