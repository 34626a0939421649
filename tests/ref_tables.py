"""The reference's own golden table `test` (numeric columns), Tests/ExecuteTest.cpp:142-186 (schema) and
:30063-30115 (three INSERT templates: x10, x5, x5; fragment_size = 2 => 10 fragments), plus a SQLite copy used
as the comparator exactly like the reference's SQLiteComparator (ExecuteTest.cpp:383-520)."""
from __future__ import annotations

import math
import sqlite3

import numpy as np

from heavydb_b200 import abi

# name, sql type, notnull
TEST_COLS = [
    ("x", abi.kINT, True),
    ("w", abi.kTINYINT, False),
    ("y", abi.kINT, False),
    ("z", abi.kSMALLINT, False),
    ("t", abi.kBIGINT, False),
    ("d", abi.kDOUBLE, False),
    ("dn", abi.kDOUBLE, False),
    ("u", abi.kINT, False),
    ("ofd", abi.kINT, False),
    ("ufd", abi.kINT, True),
    ("ofq", abi.kBIGINT, False),
    ("ufq", abi.kBIGINT, True),
    ("smallint_nulls", abi.kSMALLINT, False),
    ("f", abi.kFLOAT, False),
    ("ff", abi.kFLOAT, False),
    ("fn", abi.kFLOAT, False),
]
TEST_NAMES = [c[0] for c in TEST_COLS]

# values per INSERT template, in TEST_COLS order (None = NULL)
_T1 = (7, -8, 42, 101, 1001, 2.2, None, None, 2147483647, -2147483648, None, -1, 32767, 1.1, 1.1, None)
_T2 = (8, -7, 43, -78, 1002, 2.4, -2002.4, None, None, -2147483647, 9223372036854775807, -9223372036854775808, None, 1.2, 101.2, -101.2)
_T3 = (7, -7, 43, 102, 1002, 2.6, -220.6, None, 1, -1, 1, -9223372036854775808, 1, 1.3, 1000.3, -1000.3)
G_NUM_ROWS = 10  # ExecuteTest.cpp:605


def test_rows(num_rows: int = G_NUM_ROWS):
    return [_T1] * num_rows + [_T2] * (num_rows // 2) + [_T3] * (num_rows // 2)


def to_columns(rows, cols=TEST_COLS):
    out = []
    for c, (_, t, _nn) in enumerate(cols):
        null = abi.NULL_OF[t]
        out.append(np.array([null if r[c] is None else r[c] for r in rows], dtype=abi.NUMPY_OF[t]))
    return out


def make_table(rows, cols=TEST_COLS, fragment_size: int = 2) -> abi.Table:
    t = abi.Table([(ty, nn) for _, ty, nn in cols])
    arrays = to_columns(rows, cols)
    n = len(rows)
    for b in range(0, max(n, 0), fragment_size):
        t.add_host_fragment([a[b:b + fragment_size] for a in arrays])
    return t


def make_sqlite(rows, cols=TEST_COLS, name="test"):
    con = sqlite3.connect(":memory:")
    decl = ", ".join(f"{n} {'double' if t in (abi.kDOUBLE, abi.kFLOAT) else 'bigint'}" for n, t, _ in cols)
    con.execute(f"CREATE TABLE {name}({decl})")
    con.executemany(f"INSERT INTO {name} VALUES({','.join('?' * len(cols))})", rows)
    return con


EPS = 1.25e-5  # ExecuteTest.cpp:311


# SUM / AVG over a FLOAT argument: the reference adds in float (agg_sum_float, RuntimeFunctions.cpp; atomicAdd(float) on
# its GPU, in no fixed order), so two correct executions agree to float rounding of the partial sums only.  Relative
# part for sums far from zero; absolute part = the same fraction of a TYPICAL partial sum of the test columns
# (|x| ~ 50, random walk over n rows) for the sums that cancel.
FLOAT_SUM_RTOL = 2e-4


def float_sum_atol(n_rows: int) -> float:
    return FLOAT_SUM_RTOL * 50.0 * math.sqrt(max(n_rows, 1))


def assert_rows_match(ours, ref, fp_tol=EPS, fp_abs=0.0):
    """SQLiteComparator::compare_impl semantics (ExecuteTest.cpp:383-520): integers exact, fp within
    EPS*|ref|, NULL <-> NULL; rows compared as sorted multisets (our queries carry no ORDER BY)."""
    def key(r):
        return tuple((0, 0) if v is None else (1, v) for v in r)
    ours_s, ref_s = sorted(ours, key=key), sorted(ref, key=key)
    assert len(ours_s) == len(ref_s), f"row count {len(ours_s)} != {len(ref_s)}\nours={ours_s}\nref={ref_s}"
    for a, b in zip(ours_s, ref_s):
        assert len(a) == len(b)
        for va, vb in zip(a, b):
            if vb is None or va is None:
                assert va is None and vb is None, f"NULL mismatch {a} vs {b}"
            elif isinstance(vb, str) or isinstance(va, str):
                assert va == vb, f"{a} vs {b}"
            elif isinstance(vb, float) or isinstance(va, float):
                assert math.isclose(float(va), float(vb), rel_tol=fp_tol, abs_tol=fp_abs) or va == vb, f"{a} vs {b}"
            else:
                assert int(va) == int(vb), f"{a} vs {b}"
