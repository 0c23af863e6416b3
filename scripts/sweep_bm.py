"""Scratch sweep (not the bench contract): match-kernel time of the bench's C2 batch for work-item granularity
(XGM_BPI) and docid-range width (XGM_RANGE_BITS).  One index build, one searcher per setting."""
import json
import os
import statistics
import sys

sys.path.insert(0, ".")
import bench
from xapiand_b200 import xgm

cfg = bench.config("C2")
ix = xgm.Index.synthetic(cfg["docs"], bench.VOCAB, seed=bench.SEED, host_threads=16)
batches = [xgm.QueryBatch([bench.xgm_query(cfg, q) for q in bench.gen_queries(cfg, step, cfg["batch"])]) for step in (0, 16)]
settings = [(rb, bpi) for rb in (17, 18, 19, 20) for bpi in (8, 12, 16, 24, 32)]
for rb, bpi in settings:
    os.environ["XGM_RANGE_BITS"] = str(rb)
    os.environ["XGM_BPI"] = str(bpi)
    s = xgm.Searcher(ix, max_batch=cfg["batch"], max_topk=cfg["topk"])
    out = {"range_bits": rb, "bpi": bpi}
    for bi, b in enumerate(batches):
        s.submit(b); s.wait_raw()
        ms, tk = [], []
        for _ in range(8):
            s.replay()
            st = s.last_stats()
            ms.append(st.match_kernel_ms); tk.append(st.topk_kernel_ms)
        out[f"match_ms_b{bi}"] = round(statistics.mean(ms), 4)
        out[f"min_b{bi}"] = round(min(ms), 4)
        out[f"topk_b{bi}"] = round(statistics.mean(tk), 4)
        out["items"] = st.work_items
    print(json.dumps(out), flush=True)
    del s
