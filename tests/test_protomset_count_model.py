"""CPU model of the parallel form of ProtoMSet's known_matching_docs that xgm_topk_kernel implements
(DESIGN.md §3.3 / §8.1), checked against the sequential restatement in the oracle (itself pinned against
the compiled reference):

    walking the matches in docid order, a match at position i is counted iff i <= r or fewer than k earlier
    matches have a strictly greater weight, where r is the position at which ProtoMSet first raises
    min_weight: the heap-build position k when check_at_least <= k + 1, otherwise the first position
    >= check_at_least - 1 whose match enters the top-k (fewer than k earlier matches have a weight >= its own)
    — min_weight only moves when the heap is built or an item is replaced once known_matching_docs has
    reached check_at_least (protomset.h:340-400 with the `weight < min_weight -> continue` of
    matcher.cc:496-498).  With check_at_least <= k + 1 (what Xapiand passes by default) this is
    "counted iff among the first k + 1 or fewer than k earlier matches are strictly greater".

Two evaluations of "earlier and strictly greater": the O(n^2) pair count of the current kernel and the
O(n log^2 n) merge-sort dominance count planned for it — both must give the oracle's count, and the merge
sort's final order must be the MSet order (weight desc, docid asc)."""
import random

import numpy as np

from oracle import oracle as O


def first_raise(ge_before, k, cal):
    """Position r at which min_weight is first raised (len = never)."""
    n = len(ge_before)
    if cal <= k + 1:
        return k
    for i in range(max(cal - 1, k), n):
        if ge_before[i] < k:
            return i
    return n


def known_by_pairs(w, k, cal):
    n = len(w)
    gb = np.array([int(np.sum(w[:i] > w[i])) for i in range(n)])
    ge = np.array([int(np.sum(w[:i] >= w[i])) for i in range(n)])
    r = first_raise(ge, k, cal)
    idx = np.arange(n)
    return int(np.sum((idx <= r) | (gb < k)))


def known_by_merge_sort(w, k, cal):
    """Bottom-up merge sort of the docid-ordered weights by (weight desc, position asc); while run A (earlier
    positions) is merged with run B every b in B adds #{a in A : w_a > w_b} to its counter."""
    n = len(w)
    # (weights desc, positions, earlier-and-greater, earlier-and-not-less)
    runs = [([float(x)], [i], [0], [0]) for i, x in enumerate(w)]
    while len(runs) > 1:
        nxt = []
        for a in range(0, len(runs), 2):
            if a + 1 == len(runs):
                nxt.append(runs[a])
                continue
            wa, pa, ga, ea = runs[a]
            wb, pb, gb, eb = runs[a + 1]
            neg_a = [-x for x in wa]  # ascending for bisect
            neg_b = [-x for x in wb]
            out = [None] * (len(wa) + len(wb))
            for i, x in enumerate(wa):
                # b's strictly greater than a go first; on equal weights the earlier position (a) wins
                pos = i + int(np.searchsorted(neg_b, -x, side="left"))
                out[pos] = (x, pa[i], ga[i], ea[i])
            for j, x in enumerate(wb):
                greater = int(np.searchsorted(neg_a, -x, side="left"))      # a's with w_a > w_b
                not_less = int(np.searchsorted(neg_a, -x, side="right"))    # a's with w_a >= w_b precede b
                out[j + not_less] = (x, pb[j], gb[j] + greater, eb[j] + not_less)
            nxt.append(tuple([o[c] for o in out] for c in range(4)))
        runs = nxt
    ws, ps, gs, es = runs[0]
    order_ok = all((ws[i] > ws[i + 1]) or (ws[i] == ws[i + 1] and ps[i] < ps[i + 1]) for i in range(n - 1))
    ge_by_pos = [0] * n
    for p, e in zip(ps, es):
        ge_by_pos[p] = e
    r = first_raise(ge_by_pos, k, cal)
    known = sum(1 for p, g in zip(ps, gs) if p <= r or g < k)
    return known, order_ok, ps


def test_parallel_count_rule_matches_sequential_protomset():
    nd, V = 4000, 600
    ix = O.Index.synthetic(nd, V)
    rng = random.Random(3)
    checked = 0
    for _ in range(120):
        op = rng.choice([O.OP_AND, O.OP_OR])
        terms = rng.sample(range(60), rng.choice([2, 3]) if op == O.OP_AND else rng.choice([2, 4]))
        k = rng.choice([1, 3, 10, 40])
        cal = rng.choice([0, 0, 7, 25, 60, nd])
        got = ix.match(O.Query(op=op, terms=terms, first=0, maxitems=k, check_at_least=cal))
        allm = ix.match(O.Query(op=op, terms=terms, first=0, maxitems=nd, check_at_least=nd))  # every match
        order = np.argsort(allm.docids, kind="stable")
        w = np.asarray(allm.weights)[order]  # weights in docid order
        kk = min(k, nd)
        ccal = max(min(cal, nd), kk)  # Enquire::get_mset clamping, api/enquire.cc:420-426
        if len(w) == 0:
            continue
        a = known_by_pairs(w, kk, ccal)
        b, order_ok, ps = known_by_merge_sort(w, kk, ccal)
        assert a == b == got.known, (op, terms, k, cal, a, b, got.known)
        assert order_ok
        # the merge sort's final order is the MSet order
        ranked = np.asarray(allm.docids)[order][ps]
        assert list(ranked[:len(got.docids)]) == list(got.docids)
        checked += 1
    assert checked > 100
