"""Full-size checks (BASELINE.json C2/C3 sizes: 10M documents, 1M-term Zipf vocabulary).

The sequential oracle is too slow to build a 10M-document index inside a test, so parity at this size
goes through size-independent properties plus a vectorised numpy restatement of the scoring applied to the
posting lists the index itself decodes (round-trip-checked elsewhere):

* AND (C2): for sampled queries the complete MSet is recomputed — intersection of the three decoded lists,
  BM25Weight::get_sumpart (weight/bm25weight.cc:170-181) in the reference's operation and summation order
  (MultiAndPostList::get_weight, matcher/multiandpostlist.cc:149-159), ranking by (weight desc, docid asc)
  (msetcmp.cc:54-61) — and must agree bit for bit, together with the exact match count;
* OR (C3): same with the union and the OrPostList tree order (queryinternal.cc:440-489, oracle's program);
* every MSet is sorted under the reference's order, top-10 is a prefix of top-100, results do not depend
  on the batch they travel in, nor on the membership bitmaps (index rebuilt with XGM_BITMAP_K=0);
* sharding invariance (C4's data path at C2 size): two 5M-document shards with global statistics, one
  gathered slab per shard, xgm_merge_topk_device_slab == the unsharded index's MSets.
"""
import ctypes as C
import os
import random

import numpy as np
import pytest

from oracle import oracle as O
from xapiand_b200 import xgm

pytestmark = pytest.mark.gpu

NDOCS, VOCAB, SEED = 10_000_000, 1_000_000, 20240611
K1, K3, B, MNL = 1.0, 1.0, 0.5, 0.5


def term(r):
    return f"T{r:06d}"


@pytest.fixture(scope="module")
def full():
    ix = xgm.Index.synthetic(NDOCS, VOCAB, seed=SEED)
    info = ix.info()
    assert info.doccount == NDOCS
    doclen = ix.doclengths()
    cache = {}

    def postings(rank):
        if rank not in cache:
            cache[rank] = ix.decode_term(ix.term_stats(term(rank)).term_id)
        return cache[rank]

    return ix, info, doclen, postings


def bm25_init(N, total_length, tf):
    tw, lf = C.c_double(), C.c_double()
    O.lib().orc_bm25_init(N, total_length, tf, 1, 1.0, K1, K3, B, C.byref(tw), C.byref(lf))
    return tw.value, lf.value


def sumpart(tw, len_factor, wdf, dlen):
    """BM25Weight::get_sumpart over arrays, one IEEE operation per numpy call (no contraction)."""
    normlen = np.maximum(dlen.astype(np.float64) * len_factor, MNL)
    denom = K1 * (normlen * B + (1 - B)) + wdf.astype(np.float64)
    return tw * (wdf.astype(np.float64) / denom)


def rank_topk(docids, weights, k):
    order = np.lexsort((docids, -weights))[:k]
    return docids[order], weights[order]


def expected_and(ranks, info, doclen, postings, k):
    lists = [postings(r) for r in ranks]
    tfs = [len(d) for d, _ in lists]
    order = O.and_order(tfs)
    docs = lists[order[0]][0]
    for j in order[1:]:
        other = lists[j][0]
        pos = np.minimum(np.searchsorted(other, docs), len(other) - 1)
        docs = docs[other[pos] == docs]
    w = np.zeros(len(docs), np.float64)
    for j in order:
        d, wdf = lists[j]
        pos = np.searchsorted(d, docs)
        tw, lf = bm25_init(info.doccount, info.total_length, tfs[j])
        w = w + sumpart(tw, lf, wdf[pos], doclen[docs])
    return (len(docs),) + rank_topk(docs, w, k)


def expected_or(ranks, info, doclen, postings, k):
    lists = [postings(r) for r in ranks]
    tfs = [len(d) for d, _ in lists]
    docs = np.unique(np.concatenate([d for d, _ in lists]))
    leafw, leafhas = [], []
    for (d, wdf), tf in zip(lists, tfs):
        pos = np.minimum(np.searchsorted(d, docs), len(d) - 1)
        has = d[pos] == docs
        tw, lf = bm25_init(info.doccount, info.total_length, tf)
        leafw.append(np.where(has, sumpart(tw, lf, wdf[pos], doclen[docs]), 0.0))
        leafhas.append(has)
    stack = []
    for op in O.or_program(tfs):  # postfix: leaf index, or -1 = OrPostList node (l, r or l + r)
        if op >= 0:
            stack.append((leafw[op], leafhas[op]))
        else:
            rw, rh = stack.pop()
            lw, lh = stack.pop()
            stack.append((np.where(lh & rh, lw + rw, np.where(rh, rw, lw)), lh | rh))
    w = stack[0][0]
    return (len(docs),) + rank_topk(docs, w, k)


def assert_sorted(m, ctx):
    w = np.asarray(m.weights)
    d = np.asarray(m.docids).astype(np.int64)
    assert np.all((w[:-1] > w[1:]) | ((w[:-1] == w[1:]) & (d[:-1] < d[1:]))), f"{ctx}: not in (weight desc, docid asc) order"


def test_c2_and_fullsize_parity_and_properties(full):
    ix, info, doclen, postings = full
    rng = random.Random(777)
    terms = [rng.sample(range(1000), 3) for _ in range(1024)]
    s = xgm.Searcher(ix, max_batch=1024, max_topk=100)
    res = s.search([xgm.Query(xgm.OP_AND, [term(t) for t in q], maxitems=100) for q in terms])
    for i, m in enumerate(res):
        assert m.status == 0
        assert_sorted(m, f"and[{i}]")
    # complete recomputation for a sample (hot ranks make the lists long: keep it to a few dozen queries)
    for i in range(0, 1024, 43):
        n, d, w = expected_and(terms[i], info, doclen, postings, 100)
        m = res[i]
        assert m.exact_matches == n, f"and[{i}] {terms[i]}: {m.exact_matches} matches, expected {n}"
        assert list(m.docids) == list(d), f"and[{i}] {terms[i]} docids"
        assert np.asarray(m.weights).tobytes() == w.tobytes(), f"and[{i}] {terms[i]} weights"
        # ProtoMSet::finalise (protomset.h:497-505): a result set that did not fill up is known exactly
        assert m.matches_upper_bound == (min(len(postings(t)[0]) for t in terms[i]) if n > 100 else n)
    # top-10 is a prefix of top-100; single-query calls give what the batch gave
    res10 = s.search([xgm.Query(xgm.OP_AND, [term(t) for t in q], maxitems=10) for q in terms[:256]])
    for i, (a, b) in enumerate(zip(res10, res)):
        assert list(a.docids) == list(b.docids[:10]) and np.asarray(a.weights).tobytes() == np.asarray(b.weights[:10]).tobytes(), i
    one = xgm.Searcher(ix, max_batch=1, max_topk=100)
    for i in range(0, 64, 7):
        m = one.search([xgm.Query(xgm.OP_AND, [term(t) for t in terms[i]], maxitems=100)])[0]
        assert list(m.docids) == list(res[i].docids) and np.asarray(m.weights).tobytes() == np.asarray(res[i].weights).tobytes()
        assert m.exact_matches == res[i].exact_matches


def test_c3_or_fullsize_parity_and_properties(full):
    ix, info, doclen, postings = full
    rng = random.Random(778)
    terms = [rng.sample(range(1000), 5) for _ in range(96)]
    s = xgm.Searcher(ix, max_batch=96, max_topk=1000)
    res = s.search([xgm.Query(xgm.OP_OR, [term(t) for t in q], maxitems=1000) for q in terms])
    for i, m in enumerate(res):
        assert m.status == 0 and m.size() == 1000
        assert_sorted(m, f"or[{i}]")
    for i in range(0, 96, 16):
        n, d, w = expected_or(terms[i], info, doclen, postings, 1000)
        m = res[i]
        assert list(m.docids) == list(d), f"or[{i}] {terms[i]} docids"
        assert np.asarray(m.weights).tobytes() == w.tobytes(), f"or[{i}] {terms[i]} weights"
        if m.flags & xgm.MSET_COUNT_LOWER_BOUND:  # MaxScore skipped whole posting-list segments
            assert m.exact_matches <= n
        else:
            assert m.exact_matches == n, f"or[{i}]: {m.exact_matches} vs {n}"
    res100 = s.search([xgm.Query(xgm.OP_OR, [term(t) for t in q], maxitems=100) for q in terms])
    for i, (a, b) in enumerate(zip(res100, res)):
        assert list(a.docids) == list(b.docids[:100]), i


def test_fullsize_results_do_not_depend_on_bitmaps(full):
    ix, info, doclen, postings = full
    rng = random.Random(779)
    qs = [xgm.Query(xgm.OP_AND, [term(t) for t in rng.sample(range(1000), 3)], maxitems=100) for _ in range(256)]
    qs += [xgm.Query(xgm.OP_OR, [term(t) for t in rng.sample(range(1000), 5)], maxitems=200) for _ in range(16)]
    a = xgm.Searcher(ix, max_batch=len(qs), max_topk=200).search(qs)
    old = os.environ.get("XGM_BITMAP_K")
    os.environ["XGM_BITMAP_K"] = "0"
    try:
        plain = xgm.Index.synthetic(NDOCS, VOCAB, seed=SEED)
    finally:
        if old is None:
            del os.environ["XGM_BITMAP_K"]
        else:
            os.environ["XGM_BITMAP_K"] = old
    assert plain.info().nbitmaps == 0 and info.nbitmaps > 0
    b = xgm.Searcher(plain, max_batch=len(qs), max_topk=200).search(qs)
    for i, (x, y) in enumerate(zip(a, b)):
        assert x.status == 0 and y.status == 0
        assert list(x.docids) == list(y.docids), i
        assert np.asarray(x.weights).tobytes() == np.asarray(y.weights).tobytes(), i
        assert x.matches_upper_bound == y.matches_upper_bound
        if not ((x.flags | y.flags) & xgm.MSET_COUNT_LOWER_BOUND):
            assert x.exact_matches == y.exact_matches, i


class _CudaArray:
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def test_fullsize_sharding_invariance_through_device_merge(full):
    """C4's data path at C2 size: shard s of 2 holds global docid (local-1)*2 + s + 1 (backends/multi.h:66-70);
    global statistics (Xapiand's phase 1, handler.cc:1532-1538) make the weights shard-independent, so
    gathering the shards' result slabs and merging them (Matcher::merge_mset) must reproduce the MSets of
    the unsharded index."""
    import torch
    ix, info, doclen, postings = full
    n, K, nq = 2, 100, 512
    rng = random.Random(780)
    terms = [rng.sample(range(1000), 3) for _ in range(nq)]
    whole = xgm.Searcher(ix, max_batch=nq, max_topk=K).search(
        [xgm.Query(xgm.OP_AND, [term(t) for t in q], maxitems=K) for q in terms])
    shards = [xgm.Index.synthetic(NDOCS, VOCAB, seed=SEED, nshards=n, shard=s) for s in range(n)]
    sinfo = [x.info() for x in shards]
    assert sum(i.doccount for i in sinfo) == info.doccount and sum(i.total_length for i in sinfo) == info.total_length
    slabs = []
    keep = []
    for sh in shards:
        s = xgm.Searcher(sh, max_batch=nq, max_topk=K)
        keep.append(s)
        batch = []
        for q in terms:
            gtf = [sum(x.term_stats(term(t)).termfreq for x in shards) for t in q]
            assert gtf == [len(postings(t)[0]) for t in q]
            batch.append(xgm.Query(xgm.OP_AND, [term(t) for t in q], maxitems=K,
                                   stats=(info.doccount, info.total_length, gtf)))
        s.search(batch)
        base, nbytes, off_d, off_c, stride = s.device_slab()
        slabs.append(torch.as_tensor(_CudaArray(base, (nbytes,), "|u1"), device="cuda").clone())
    G = torch.cat(slabs)
    ow = torch.zeros(nq * K, dtype=torch.float64, device="cuda")
    od = torch.zeros(nq * K, dtype=torch.int32, device="cuda")
    on = torch.zeros(nq, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    st = xgm.lib().xgm_merge_topk_device_slab(G.data_ptr(), nbytes, off_d, off_c, n, nq, stride, K, ow.data_ptr(),
                                              od.data_ptr(), on.data_ptr(), None)
    assert st == 0, xgm.lib().xgm_last_error()
    torch.cuda.synchronize()
    ow = ow.cpu().numpy().reshape(nq, K)
    od = od.cpu().numpy().view(np.uint32).reshape(nq, K)
    on = on.cpu().numpy()
    same_order = 0
    for i, m in enumerate(whole):
        c = int(on[i])
        assert c == m.size(), i
        # each shard sums the leaves in ITS termfreq order (as each reference shard's MultiAndPostList does);
        # when that order equals the unsharded one the f64 sums, hence the MSets, are bit-identical
        gorder = O.and_order([len(postings(t)[0]) for t in terms[i]])
        if all(O.and_order([x.term_stats(term(t)).termfreq for t in terms[i]]) == gorder for x in shards):
            same_order += 1
            assert list(od[i, :c]) == list(m.docids), f"shard merge[{i}] {terms[i]}"
            assert ow[i, :c].tobytes() == np.asarray(m.weights).tobytes(), f"shard merge[{i}] weights"
        else:
            assert np.allclose(ow[i, :c], np.asarray(m.weights), rtol=1e-12, atol=0), f"shard merge[{i}] weights"
    assert same_order > nq // 2
