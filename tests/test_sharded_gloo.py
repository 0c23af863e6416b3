"""CPU test of the N>1 host logic with world_size 2 over gloo: statistics all-reduce, per-shard
matching with the global statistics (the oracle stands in for the kernels, this box has no GPU),
unshard + all-gather + Matcher::merge_mset through the C-ABI's host merge — checked against a golden
fixture produced by the compiled reference running Xapiand's two-phase scheme over 2 shards."""
import os
import socket
import struct
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from oracle import oracle as O
    from tests.golden_util import load
    from xapiand_b200 import sharded, xgm
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fx = load("shard2_20k")
        assert fx["nshards"] == world
        shard = O.Index.synthetic(fx["ndocs"], fx["vocab"], fx["seed"], nshards=world, shard=rank)
        bad = []
        for i, gq in enumerate(fx["queries"]):
            ltf = [shard.termfreq(t) for t in gq["terms"]]
            coll, tlen, gtf = sharded.global_stats(ltf, shard.doccount, shard.total_length)
            m = shard.match(O.Query(op=O.OP_AND if gq["op"] == "AND" else O.OP_OR, terms=gq["terms"], first=0,
                                    maxitems=gq["first"] + gq["maxitems"], check_at_least=gq["check_at_least"],
                                    stats=(coll, tlen, gtf)))
            local = xgm.MSet(m.docids, m.weights, None, 0, m.lb, m.est, m.ub, m.max_possible, m.max_attained,
                             m.percent_scale_factor, m.exact, 0, 0)
            merged = sharded.merge_over_ranks(local, gq["first"], gq["maxitems"])
            if rank == 0:
                ok = (list(merged.docids) == gq["docids"]
                      and all(struct.pack("<d", a) == struct.pack("<d", b) for a, b in zip(merged.weights, gq["weights"]))
                      and struct.pack("<d", merged.max_possible) == struct.pack("<d", gq["max_possible"])
                      and struct.pack("<d", merged.max_attained) == struct.pack("<d", gq["max_attained"])
                      and (merged.matches_lower_bound, merged.get_matches_estimated(), merged.matches_upper_bound)
                      == (gq["lb"], gq["est"], gq["ub"])
                      # MSetIterator::get_percent of the merged MSet: the percent scale of the shard with the
                      # highest max_attained (MSet::Internal::merge_stats, api/mset.cc:376-395)
                      and [O.convert_to_percent(w, merged.percent_scale_factor) for w in merged.weights] == gq["percents"])
                if not ok:
                    bad.append(i)
        if rank == 0:
            q.put(bad)
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_matches_reference_twophase():
    import __graft_entry__ as g
    g.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    bad = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert bad == [], f"queries differing from the reference: {bad}"


def test_unshard_matches_reference_formula():
    from xapiand_b200 import xgm
    d = np.array([1, 2, 3, 1000], np.uint32)
    # unshard(shard_did, shard, n) = (shard_did - 1) * n + shard + 1, backends/multi.h:66-70
    assert list(xgm.unshard(d, 2, 8)) == [3, 11, 19, 7995]
