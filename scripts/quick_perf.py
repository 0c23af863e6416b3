"""Scratch perf probe (not the bench contract): build an index, time batches through the C-ABI."""
import json
import random
import sys
import time

sys.path.insert(0, ".")
import numpy as np
from xapiand_b200 import xgm

ndocs = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
vocab = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
nq = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
t0 = time.time()
ix = xgm.Index.synthetic(ndocs, vocab)
t1 = time.time()
info = ix.info()
print(json.dumps(dict(build_s=round(t1 - t0, 2), docs=info.doccount, postings=info.npostings, blocks=info.nblocks,
                      bytes_docids=info.bytes_docids, bytes_wdfs=info.bytes_wdfs, bytes_headers=info.bytes_headers)))
rng = random.Random(777)
nq = max(nq, 1)
queries = [xgm.Query(xgm.OP_AND, [f"T{r:06d}" for r in rng.sample(range(1000), 3)], maxitems=100) for _ in range(nq)]
batch = xgm.QueryBatch(queries)
s = xgm.Searcher(ix, max_batch=nq, max_topk=100)
for it in range(5):
    t0 = time.time()
    s.submit(batch)
    d, w, k, inf = s.wait_raw()
    dt = time.time() - t0
    st = s.last_stats()
    nover = sum(1 for i in range(nq) if inf[i].status != 0)
    print(json.dumps(dict(iter=it, e2e_ms=round(dt * 1e3, 3), qps=round(nq / dt), match_ms=round(st.match_kernel_ms, 3),
                          topk_ms=round(st.topk_kernel_ms, 3), items=st.work_items, alg_MB=round(st.algorithmic_bytes / 1e6, 1),
                          alg_GBps=round(st.algorithmic_bytes / 1e6 / max(st.match_kernel_ms, 1e-6), 1), overflow=nover,
                          plan_ms=round(st.host_plan_ms, 3), wait_ms=round(st.host_wait_ms, 3),
                          mean_hits=float(np.mean([inf[i].exact_matches for i in range(nq)])))))
# OR 5-term top-1000 (config C3)
nqo = 200
oq = [xgm.Query(xgm.OP_OR, [f"T{r:06d}" for r in rng.sample(range(1000), 5)], maxitems=1000) for _ in range(nqo)]
ob = xgm.QueryBatch(oq)
so = xgm.Searcher(ix, max_batch=nqo, max_topk=1000)
for it in range(3):
    t0 = time.time()
    so.submit(ob)
    d2, w2, k2, inf2 = so.wait_raw()
    dt = time.time() - t0
    st = so.last_stats()
    print(json.dumps(dict(or_iter=it, e2e_ms=round(dt * 1e3, 3), qps=round(nqo / dt), match_ms=round(st.match_kernel_ms, 3),
                          topk_ms=round(st.topk_kernel_ms, 3), items=st.work_items, bad=sum(1 for i in range(nqo) if inf2[i].status != 0),
                          approx=sum(1 for i in range(nqo) if inf2[i].flags & 1), second_pass=st.second_pass_queries,
                          mean_hits=float(np.mean([inf2[i].exact_matches for i in range(nqo)])))))
# replay timing (device only)
import ctypes
for it in range(3):
    s.replay()
    st = s.last_stats()
    print(json.dumps(dict(replay=it, match_ms=round(st.match_kernel_ms, 3), topk_ms=round(st.topk_kernel_ms, 3),
                          alg_GBps=round(st.algorithmic_bytes / 1e6 / max(st.match_kernel_ms, 1e-6), 1))))
