"""GPU parity tests proper: the CUDA path (through the C-ABI) against the oracle on seeded inputs."""
import random

import numpy as np
import pytest

from oracle import oracle as O
from tests.util import assert_mset_equal, gen_queries
from xapiand_b200 import xgm

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def small():
    nd, V = 30000, 20000
    return xgm.Index.synthetic(nd, V), O.Index.synthetic(nd, V), nd, V


@pytest.fixture(scope="module")
def medium():
    nd, V = 300000, 100000
    return xgm.Index.synthetic(nd, V), O.Index.synthetic(nd, V), nd, V


def to_x(q):
    return xgm.Query(xgm.OP_AND if q["op"] == "AND" else xgm.OP_OR, q["terms"], first=q["first"],
                     maxitems=q["maxitems"], check_at_least=q["check_at_least"])


def to_o(q):
    return O.Query(op=O.OP_AND if q["op"] == "AND" else O.OP_OR, terms=q["terms"], first=q["first"],
                   maxitems=q["maxitems"], check_at_least=q["check_at_least"])


def run_and_compare(ix, orc, queries, max_topk=128, check_counts=True):
    s = xgm.Searcher(ix, max_batch=max(1, len(queries)), max_topk=max_topk)
    res = s.search([to_x(q) for q in queries])
    napprox = 0
    for i, (q, m) in enumerate(zip(queries, res)):
        ref = orc.match(to_o(q))
        approx = bool(m.flags & 1)
        napprox += approx
        assert_mset_equal(m, ref, ctx=f"query {i} {q}", check_counts=check_counts and not approx)
        # always exact, pruned or not (the match count only when no posting-list segment was skipped)
        assert m.matches_upper_bound == ref.ub
        if m.flags & 2:
            assert m.exact_matches <= ref.exact
        else:
            assert m.exact_matches == ref.exact
        if approx:  # conservative but valid bounds
            assert m.matches_lower_bound <= ref.lb and m.matches_lower_bound <= m.matches_estimated_raw <= ref.ub
    s.close()
    return res, napprox


def test_index_info_and_roundtrip(small):
    ix, orc, nd, V = small
    info = ix.info()
    assert info.doccount == nd and info.lastdocid == nd
    assert info.total_length == orc.total_length
    assert info.doclen_lower_bound == orc.doclen_lb and info.doclen_upper_bound == orc.doclen_ub
    assert info.npostings == int(orc.offsets()[-1])
    for t in [0, 1, 7, 100, 999, 5000, V - 1]:
        d, w = ix.decode_term(t)
        rd, rw = orc.postings(t)
        assert np.array_equal(d, rd) and np.array_equal(w, rw), f"term {t} round trip"
        st = ix.term_stats(orc.name(t))
        assert st.term_id == t and st.termfreq == len(rd)
        assert st.collfreq == int(orc.collfreq()[t]) and st.wdf_upper_bound == int(orc.wdf_ub()[t])
    assert ix.term_stats("nosuchterm").termfreq == 0


def test_and_3term_top100(small):
    ix, orc, nd, V = small
    rng = random.Random(1)
    run_and_compare(ix, orc, gen_queries(rng, 200, 1000, nd))


def test_and_various_shapes(small):
    ix, orc, nd, V = small
    rng = random.Random(2)
    qs = gen_queries(rng, 200, 300, nd, ks=(1, 2, 3, 4, 6), maxitems=(1, 5, 10, 100), first=(0, 0, 3, 10))
    qs += gen_queries(rng, 100, 300, nd, ks=(2, 3), maxitems=(10, 50), check_all=True)
    run_and_compare(ix, orc, qs)


def test_and_rare_terms_and_absent(small):
    ix, orc, nd, V = small
    rng = random.Random(3)
    qs = [dict(op="AND", terms=[rng.randrange(0, 50), rng.randrange(5000, V)], first=0, maxitems=10, check_at_least=0)
          for _ in range(100)]
    run_and_compare(ix, orc, qs)
    # absent term by name
    s = xgm.Searcher(ix, 4, 16)
    r = s.search([xgm.Query(xgm.OP_AND, ["T000001", "nosuchterm"], maxitems=10)])[0]
    assert r.status == 0 and r.size() == 0 and r.matches_upper_bound == 0


def test_and_medium_index(medium):
    ix, orc, nd, V = medium
    rng = random.Random(4)
    run_and_compare(ix, orc, gen_queries(rng, 150, 1000, nd))


def test_single_query_calls_match_batch(small):
    ix, orc, nd, V = small
    rng = random.Random(5)
    qs = gen_queries(rng, 20, 500, nd)
    s = xgm.Searcher(ix, 32, 128)
    batch = s.search([to_x(q) for q in qs])
    for q, b in zip(qs, batch):
        one = s.search([to_x(q)])[0]
        assert np.array_equal(one.docids, b.docids) and one.weights.tobytes() == b.weights.tobytes()


def test_or_5term_top1000(small):
    ix, orc, nd, V = small
    rng = random.Random(6)
    qs = gen_queries(rng, 60, 1000, nd, ops=("OR",), ks=(5,), maxitems=(1000,))
    run_and_compare(ix, orc, qs, max_topk=1000)


def test_or_various_shapes_exact_counts(small):
    ix, orc, nd, V = small
    rng = random.Random(7)
    qs = gen_queries(rng, 120, 2000, nd, ops=("OR",), ks=(2, 3, 5, 8), maxitems=(1, 10, 100), first=(0, 0, 5))
    qs += gen_queries(rng, 60, 2000, nd, ops=("OR",), ks=(2, 4), maxitems=(10, 100), check_all=True)
    res, napprox = run_and_compare(ix, orc, qs)


def test_dense_terms_force_pruning(small):
    """Queries over the most frequent terms have far more matches than the candidate buffer:
    the threshold pruning must still return the exact top-k."""
    ix, orc, nd, V = small
    rng = random.Random(8)
    qs = []
    for _ in range(40):
        qs.append(dict(op="AND", terms=rng.sample(range(8), rng.choice([1, 2, 3])), first=0,
                       maxitems=rng.choice([10, 100]), check_at_least=0))
        qs.append(dict(op="OR", terms=rng.sample(range(12), rng.choice([2, 3, 5])), first=rng.choice([0, 4]),
                       maxitems=rng.choice([10, 100]), check_at_least=0))
    res, napprox = run_and_compare(ix, orc, qs)
    assert napprox > 0  # the pruned path was really exercised


def test_or_medium_index(medium):
    ix, orc, nd, V = medium
    rng = random.Random(9)
    run_and_compare(ix, orc, gen_queries(rng, 40, 1000, nd, ops=("OR",), ks=(5,), maxitems=(1000,)), max_topk=1000)


def test_filter_andnot_andmaybe_against_oracle(small):
    """§8(f)-1 shapes on a 30k-document index (terms with and without membership bitmaps): every MSet field
    incl. exact match count and percentage scale against the oracle."""
    ix, orc, nd, V = small
    rng = random.Random(91)
    xq, oq = [], []
    for i in range(300):
        nb = rng.choice([1, 2, 2, 3])
        pool = rng.sample(range(400 if i % 3 else 4000), nb + 9)  # every third query uses rare terms too
        nf, nx, nm = rng.choice([0, 0, 1, 2]), rng.choice([0, 0, 1, 2, 3]), rng.choice([0, 0, 1, 2, 3])
        if nf + nx + nm == 0:
            nm = 2
        base, f, x, m = pool[:nb], pool[nb:nb + nf], pool[nb + 3:nb + 3 + nx], pool[nb + 6:nb + 6 + nm]
        kw = dict(first=rng.choice([0, 0, 2]), maxitems=rng.choice([10, 100]), check_at_least=rng.choice([0, 0, nd]))
        xq.append(xgm.Query(xgm.OP_AND, base, filter_terms=f, not_terms=x, maybe_terms=m, **kw))
        oq.append(O.Query(op=O.OP_AND, terms=base, filter_terms=f, not_terms=x, maybe_terms=m, **kw))
    res = xgm.Searcher(ix, max_batch=len(xq), max_topk=128).search(xq)
    for i, (m, q) in enumerate(zip(res, oq)):
        ref = orc.match(q)
        assert_mset_equal(m, ref, ctx=f"ops[{i}] {q.terms} F{q.filter_terms} N{q.not_terms} M{q.maybe_terms}",
                          check_counts=not (m.flags & 1))
        assert m.exact_matches == ref.exact and m.matches_upper_bound == ref.ub

