"""oracle_execute_generated (the oracle over a never-materialised, counter-generated table with one output buffer per worker:
the reference's multi-fragment kernel + ResultSetStorage::reduce) must be bit-identical to the per-fragment executor over the
same rows materialised on the host — it is what bench.py checks the timed 1e9-row results against."""
import numpy as np
import pytest

import oracle_lib
import sqlmini
from heavydb_b200 import abi

SEED = 0x5EED
FR = 1 << 17
SPEC = [("c0", abi.kBIGINT, 0, 0, 10**6, 1), ("c1", abi.kBIGINT, 1, 0, 10**6, 1), ("g", abi.kINT, 2, 0, 10**4, 1), ("v", abi.kDOUBLE, 3, 0, 1, 1),
        ("s", abi.kBIGINT, 4, 0, 5000, 900_000_000_007)]   # s: sparse keys -> baseline hash


def tables(n, frag_ids):
    t = abi.Table([(ty, True) for _, ty, *_ in SPEC])
    r = abi.Table([(ty, True) for _, ty, *_ in SPEC])
    left = n
    for fid in frag_ids:
        m = min(FR, left)
        left -= m
        t.add_host_fragment([oracle_lib.gen_column(ty, SEED, tag, fid * FR, m, lo, span, stride=st) for _, ty, tag, lo, span, st in SPEC], fragment_id=fid)
        r.add_remote_fragment(m, t.fragments[-1].stats, fragment_id=fid)
    return t, r


QUERIES = [
    ("SELECT g, SUM(c1), COUNT(*) FROM t WHERE c0 < 500000 GROUP BY g;", 0),
    ("SELECT g, AVG(v), MIN(c1), MAX(v) FROM t WHERE c0 >= 250000 GROUP BY g;", 0),
    ("SELECT c0, SUM(c1) FROM t GROUP BY c0;", 0),
    ("SELECT COUNT(*), SUM(c1), MIN(v) FROM t WHERE c0 < 1000;", 0),
    ("SELECT s, SUM(c1), COUNT(*) FROM t GROUP BY s;", 8000),
]


@pytest.mark.parametrize("threads", [1, 3, 8])
def test_generated_equals_materialised(threads):
    frag_ids = [0, 1, 2, 5, 9]      # ids need not be dense: the global row of a tuple is fragment_id * FR + i
    t, r = tables(4 * FR + 777, frag_ids)
    names = [s[0] for s in SPEC]
    gen = [(ty, tag, lo, span, st) for _, ty, tag, lo, span, st in SPEC]
    for sql, guess in QUERIES:
        unit = sqlmini.parse(sql, t, names)
        a = oracle_lib.execute(unit, t, entry_guess=guess, has_card=guess > 0, num_threads=2)
        b = oracle_lib.execute_generated(unit, r, gen, SEED, FR, entry_guess=guess, has_card=guess > 0, num_threads=threads)
        assert a.plan.as_dict() == b.plan.as_dict()
        if guess:   # baseline hash: slots depend on the insertion order, the rows do not
            assert sorted(a.rows()) == sorted(b.rows())
        elif "AVG(v)" in sql:   # double sums: the per-thread partition of the rows changes the order of additions
            ra, rb = sorted(a.rows()), sorted(b.rows())
            assert len(ra) == len(rb)
            for x, y in zip(ra, rb):
                assert x[0] == y[0] and x[2:] == y[2:] and abs(x[1] - y[1]) <= 1e-9 * abs(y[1])
        else:
            assert np.array_equal(a.buffer(), b.buffer())
