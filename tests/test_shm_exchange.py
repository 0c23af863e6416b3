"""The single-node phase-1 statistics exchange (xapiand_b200/shm_exchange.py): 8 processes, the calling pattern of
bench.py's end-to-end loop (post one exchange ahead, collect in order, uneven pacing), every sum checked."""
import multiprocessing as mp
import os
import random
import time

import numpy as np

from xapiand_b200.shm_exchange import ShmExchange

WORLD, NVALS, ROUNDS = 8, 1002, 400


def _vec(rank, xid):
    v = np.arange(NVALS, dtype=np.int64) * (rank + 1) + xid * 7 + rank
    v[-1] = (1 << 40) + rank * xid  # total_length-sized values survive
    return v


def _worker(name, rank, ready, go, out):
    try:
        if rank == 0:
            x = ShmExchange(name, 0, WORLD, NVALS, create=True)
            ready.set()
        else:
            ready.wait(30)
            x = ShmExchange(name, rank, WORLD, NVALS)
        go.wait(30)
        rng = random.Random(rank)
        expect_cache = {}
        posted = -1
        for xid in range(ROUNDS):
            if posted < xid:
                x.post(xid, _vec(rank, xid)); posted = xid
            got = x.collect(xid, timeout=60)
            want = expect_cache.get(xid)
            if want is None:
                want = sum(_vec(r, xid) for r in range(WORLD))
            assert (got == want).all(), f"rank {rank} exchange {xid}"
            if xid + 1 < ROUNDS:  # the next exchange is posted before the "planning" of this batch
                x.post(xid + 1, _vec(rank, xid + 1)); posted = xid + 1
            if rng.random() < 0.05:
                time.sleep(rng.random() * 0.002)  # a rank that falls behind
        x.close()
        out.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        out.put((rank, repr(e)))


def test_eight_ranks_exchange_in_lockstep():
    ctx = mp.get_context("fork")
    name = f"xgm_test_p1_{os.getpid()}"
    ready, go, out = ctx.Event(), ctx.Event(), ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(name, r, ready, go, out)) for r in range(WORLD)]
    for p in ps:
        p.start()
    go.set()
    res = [out.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(30)
    assert sorted(res) == [(r, "ok") for r in range(WORLD)], res
    assert not os.path.exists(os.path.join("/dev/shm", name))


def _bench_pattern_worker(name, rank, world, ready, go, out):
    """The exact call sequence of bench.py's N > 1 path: warm-up (every batch index three times in a row, prefetches
    that are collected late or never), exchange_stats(0), then the end-to-end loop over consecutive batches."""
    try:
        nvals, nbatches = 1002, 32
        if rank == 0:
            x = ShmExchange(name, 0, world, nvals, create=True)
            ready.set()
        else:
            ready.wait(30)
            x = ShmExchange(name, rank, world, nvals)
        go.wait(30)
        pending, next_xid = {}, [0]

        def local(bi):
            return np.full(nvals, (rank + 1) * 1000 + bi, np.int64)

        def start(bi):
            if bi in pending or bi >= nbatches:
                return
            xid = next_xid[0]; next_xid[0] += 1
            x.post(xid, local(bi)); pending[bi] = xid

        def exchange(bi):
            start(bi)
            got = x.collect(pending.pop(bi), timeout=60)
            want = sum((r + 1) * 1000 + bi for r in range(world))
            assert (got == want).all(), f"rank {rank} batch {bi}: {got[0]} != {want}"
            start(bi + 1)

        rng = random.Random(100 + rank)
        for w in range(3):
            for _ in range(3):
                exchange(w % nbatches)
                if rng.random() < 0.3:
                    time.sleep(rng.random() * 0.001)
        exchange(0)
        for k in range(24):
            exchange(3 + 1 + k)
            if rng.random() < 0.1:
                time.sleep(rng.random() * 0.002)
        x.close()
        out.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        out.put((rank, repr(e)))


def test_bench_call_pattern_eight_ranks():
    ctx = mp.get_context("fork")
    name = f"xgm_test_p1b_{os.getpid()}"
    ready, go, out = ctx.Event(), ctx.Event(), ctx.Queue()
    ps = [ctx.Process(target=_bench_pattern_worker, args=(name, r, WORLD, ready, go, out)) for r in range(WORLD)]
    for p in ps:
        p.start()
    go.set()
    res = [out.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(30)
    assert sorted(res) == [(r, "ok") for r in range(WORLD)], res
