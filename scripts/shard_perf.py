"""Scratch probe: per-GPU step cost of the bench batch when the 10M-doc corpus is cut into N shards (one GPU runs
shard 0 of N) — where the strong-scaling time goes: match kernel, top-k kernel, everything else, merge."""
import json
import random
import sys

sys.path.insert(0, ".")
import torch
from xapiand_b200 import xgm

NDOCS, VOCAB, BATCH, TOPK = 10_000_000, 1_000_000, 4096, 100


class CudaArray:
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


rng = random.Random(20240611)
terms = [rng.sample(range(1000), 3) for _ in range(BATCH)]
L = xgm.lib()
for ns in [int(a) for a in sys.argv[1:]] or [1, 4, 8]:
    ix = xgm.Index.synthetic(NDOCS, VOCAB, seed=1, nshards=ns, shard=0)
    s = xgm.Searcher(ix, max_batch=BATCH, max_topk=TOPK)
    st_ = torch.cuda.ExternalStream(s.stream())
    b = xgm.QueryBatch([xgm.Query(xgm.OP_AND, [f"T{t:06d}" for t in q], maxitems=TOPK) for q in terms])
    for _ in range(3):
        s.submit(b); s.wait_raw()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(st_)
    for _ in range(20):
        s.replay()
    e1.record(st_)
    torch.cuda.synchronize()
    step = e0.elapsed_time(e1) / 20
    mm, tt = [], []
    for _ in range(10):
        s.replay()
        st = s.last_stats()
        mm.append(st.match_kernel_ms); tt.append(st.topk_kernel_ms)
    base, nbytes, off_d, off_c, stride = s.device_slab()
    local = torch.as_tensor(CudaArray(base, (nbytes,), "|u1"), device="cuda")
    g = local.repeat(ns)
    ow = torch.empty(BATCH * TOPK, dtype=torch.float64, device="cuda")
    od = torch.empty(BATCH * TOPK, dtype=torch.int32, device="cuda")
    on = torch.empty(BATCH, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    e0.record(st_)
    for _ in range(20):
        L.xgm_merge_topk_device_slab(g.data_ptr(), nbytes, off_d, off_c, ns, BATCH, stride, TOPK, ow.data_ptr(), od.data_ptr(),
                                     on.data_ptr(), s.stream())
    e1.record(st_)
    torch.cuda.synchronize()
    merge = e0.elapsed_time(e1) / 20
    print(json.dumps(dict(nshards=ns, step_ms=round(step, 4), match_ms=round(sum(mm) / len(mm), 4),
                          topk_ms=round(sum(tt) / len(tt), 4), merge_ms=round(merge, 4), slab_MB=round(nbytes / 1e6, 2),
                          items=st.work_items, alg_MB=round(st.algorithmic_bytes / 1e6, 1))))
    del s, ix
