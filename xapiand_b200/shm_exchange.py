"""Phase-1 statistics exchange between the ranks of ONE node through shared memory.

Xapiand's two-phase scheme sums, per request, every shard's term frequencies, document count and total length
before any shard weighs a document (`src/database/handler.cc:1532-1538`, `src/xapian/weight/weightinternal.cc:54-72`).
With one process per GPU on one box that sum is host data on both sides (term dictionary in, query planner out) and a
few kilobytes per batch: a collective library is the wrong tool for it — a gloo all-reduce between 8 local processes
was measured at 2.4 ms per batch, six times the GPU time of the batch.  Here every rank writes its vector into its own
row of a slab in /dev/shm and publishes a sequence number; every rank sums the rows once all sequence numbers are there.
Tens of microseconds, no GPU stream, no ordering against the result exchange on the NCCL communicator.

Ranks on several nodes need a network collective instead (`bench.py` falls back to a gloo group).

Ordering: the row is written before its sequence number and read after it.  x86's store-store / load-load ordering
(TSO) is what makes that sufficient; on a weaker memory model the two numpy stores would need a fence between them.
"""
import mmap
import os
import time

import numpy as np


class ShmExchange:
    """post(xid, values) / collect(xid) with xid = 0, 1, 2, ... identical on all ranks; `slots` exchanges may be in
    flight (a rank never runs more than two posts ahead of the slowest rank's collect)."""

    def __init__(self, name: str, rank: int, world: int, nvals: int, slots: int = 8, create: bool = False):
        self.rank, self.world, self.nvals, self.slots = rank, world, nvals, slots
        self.path = os.path.join("/dev/shm", name)
        nbytes = 8 * slots * world * (1 + nvals)
        if create:
            try:
                os.unlink(self.path)  # a leftover of a run that died
            except FileNotFoundError:
                pass
            fd = os.open(self.path, os.O_CREAT | os.O_EXCL | os.O_RDWR, 0o600)
            os.ftruncate(fd, nbytes)  # zero-filled: no sequence number matches yet
        else:
            fd = os.open(self.path, os.O_RDWR)
        try:
            self.map = mmap.mmap(fd, nbytes)
        finally:
            os.close(fd)
        flat = np.frombuffer(self.map, dtype=np.int64)
        self.seq = flat[: slots * world].reshape(slots, world)
        self.data = flat[slots * world:].reshape(slots, world, nvals)
        self.owner = create

    def post(self, xid: int, values) -> None:
        s = xid % self.slots
        self.data[s, self.rank, :] = values
        self.seq[s, self.rank] = xid + 1

    def collect(self, xid: int, timeout: float = 120.0) -> np.ndarray:
        s = xid % self.slots
        want = xid + 1
        seq = self.seq[s]
        if not (seq == want).all():
            deadline = time.perf_counter() + timeout
            spins = 0
            while not (seq == want).all():
                spins += 1
                if spins > 2000:  # ~1 ms of polling: stop burning the core the other ranks' planners need
                    time.sleep(0.00005)
                    if time.perf_counter() > deadline:
                        raise TimeoutError(f"statistics exchange {xid}: ranks {np.nonzero(seq != want)[0].tolist()} did not post")
        return self.data[s].sum(axis=0)

    def close(self) -> None:
        self.seq = self.data = None
        try:
            self.map.close()
        except BufferError:
            pass
        if self.owner:
            try:
                os.unlink(self.path)
            except FileNotFoundError:
                pass
