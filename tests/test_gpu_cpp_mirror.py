"""GPU test of the C++ host mirror (include/xgm_enquire.hpp): a program written against the
Xapian-shaped classes is compiled with g++, linked to libxgm.so and compared with the oracle."""
import os
import struct
import subprocess

import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_enquire_mirror_matches_oracle(tmp_path):
    exe = str(tmp_path / "enquire_mirror")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "enquire_mirror.cc"), "-o", exe,
                           "-L" + os.path.join(ROOT, "xapiand_b200"), "-lxgm",
                           "-Wl,-rpath," + os.path.join(ROOT, "xapiand_b200")])
    nd, V = 20000, 3000
    orc = O.Index.synthetic(nd, V)
    cases = [("AND", 0, 10, 0, [30, 170, 400], {}), ("OR", 2, 25, 0, [5, 120, 700], {}), ("AND", 0, 5, nd, [250], {}),
             # OP_AND_MAYBE(OP_AND_NOT(OP_FILTER(AND, boolean term), OR of terms), OR of terms)
             ("AND", 0, 20, nd, [12, 60], dict(filter_terms=[3], not_terms=[40, 9], maybe_terms=[75, 110]))]
    for op, first, maxitems, cal, terms, groups in cases:
        env = dict(os.environ)
        for key, var in (("filter_terms", "XGM_MIRROR_FILTER"), ("not_terms", "XGM_MIRROR_NOT"), ("maybe_terms", "XGM_MIRROR_MAYBE")):
            if groups.get(key):
                env[var] = ",".join(f"T{t:06d}" for t in groups[key])
        out = subprocess.check_output([exe, str(nd), str(V), op, str(first), str(maxitems), str(cal)] +
                                      [f"T{t:06d}" for t in terms], timeout=240, env=env).decode().splitlines()
        ref = orc.match(O.Query(op=O.OP_AND if op == "AND" else O.OP_OR, terms=terms, first=first, maxitems=maxitems,
                                check_at_least=cal, **groups))
        head = out[0].split()
        n = int(head[1])
        assert n == len(ref.docids)
        assert int(head[4]) == ref.ub
        if int(head[7]) == 0:  # bounds not flagged approximate
            assert int(head[2]) == ref.lb
            assert int(head[3]) == O.round_estimate(ref.lb, ref.ub, ref.est)
        else:
            assert int(head[2]) <= ref.lb
        assert struct.pack("<d", float(head[5])) == struct.pack("<d", ref.max_possible)
        assert struct.pack("<d", float(head[6])) == struct.pack("<d", ref.max_attained)
        for line, d, w in zip(out[1:1 + n], ref.docids, ref.weights):
            p = line.split()
            assert int(p[0]) == int(d) and struct.pack("<d", float(p[1])) == struct.pack("<d", float(w))
            expect_pct = 100 if ref.percent_scale_factor == 0 else max(1, min(100, int(float(w) * ref.percent_scale_factor + 100.0 * 2.220446049250313e-16)))
            assert int(p[2]) == expect_pct
        assert out[1 + n] == "DECLINED"
