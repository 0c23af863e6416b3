"""Scratch perf probe of BASELINE config C3 (5-term OR, top-1000) — not the bench contract."""
import json
import random
import sys

sys.path.insert(0, ".")
import numpy as np
from xapiand_b200 import xgm

ndocs = int(sys.argv[1]) if len(sys.argv) > 1 else 10000000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 512
ix = xgm.Index.synthetic(ndocs, 1000000)
rng = random.Random(777)
oq = [xgm.Query(xgm.OP_OR, [f"T{r:06d}" for r in rng.sample(range(1000), 5)], maxitems=1000) for _ in range(nq)]
ob = xgm.QueryBatch(oq)
so = xgm.Searcher(ix, max_batch=nq, max_topk=1000)
for it in range(4):
    so.submit(ob)
    d2, w2, k2, inf2 = so.wait_raw()
    st = so.last_stats()
    print(json.dumps(dict(or_iter=it, nq=nq, match_ms=round(st.match_kernel_ms, 3), topk_ms=round(st.topk_kernel_ms, 3),
                          items=st.work_items, bad=sum(1 for i in range(nq) if inf2[i].status != 0),
                          approx=sum(1 for i in range(nq) if inf2[i].flags & 1), second_pass=st.second_pass_queries,
                          alg_GBps=round(st.algorithmic_bytes / 1e6 / max(st.match_kernel_ms, 1e-6), 1),
                          mean_hits=float(np.mean([inf2[i].exact_matches for i in range(nq)])))))
