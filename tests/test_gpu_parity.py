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
    noverflow = 0
    for i, (q, m) in enumerate(zip(queries, res)):
        ref = orc.match(to_o(q))
        assert_mset_equal(m, ref, ctx=f"query {i} {q}", check_counts=check_counts)
    s.close()
    return res


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
