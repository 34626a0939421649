"""The reference's golden join pair: `test` (ref_tables.py) and `test_inner` (Tests/ExecuteTest.cpp:29719-29739:
x int not null, y int, xx smallint, dt DATE, dt32 DATE ENCODING FIXED(32), dt16 DATE ENCODING FIXED(16), ts; rows
(7, 43, 7, 1999-09-09 ...) and (-9, 72, -9, 2014-12-13 ...)) with the join queries of the Select.Joins_* tests that
fall on this path (one equi-join level on a unique inner key)."""
import sqlite3

import numpy as np

from heavydb_b200 import abi

DAY = 86400
INNER_COLS = [
    ("x", abi.kINT, True, 0), ("y", abi.kINT, False, 0), ("xx", abi.kSMALLINT, False, 0),
    ("dt", abi.kDATE, False, -4), ("dt32", abi.kDATE, False, -4), ("dt16", abi.kDATE, False, -2), ("ts", abi.kTIMESTAMP, False, 0),
]
INNER_NAMES = [c[0] for c in INNER_COLS]
INNER_ROWS = [
    (7, 43, 7, 10843 * DAY, 10843 * DAY, 10843 * DAY, 1418509395),          # '1999-09-09' x3, '2014-12-13 22:23:15'
    (-9, 72, -9, 16417 * DAY, 16417 * DAY, 16417 * DAY, 10843 * DAY + 14 * 3600 + 15 * 60 + 16),   # '2014-12-13' x3, '1999-09-09 14:15:16'
]


def inner_table() -> abi.Table:
    t = abi.Table([(ty, nn) for _, ty, nn, _ in INNER_COLS], encoded_sizes=[e for *_, e in INNER_COLS])
    arrays = []
    for c, (_, _ty, _nn, enc) in enumerate(INNER_COLS):
        vals = [r[c] // DAY if enc < 0 else r[c] for r in INNER_ROWS]
        arrays.append(np.array(vals, dtype=t.physical_dtype(c)))
    t.add_host_fragment(arrays)          # the hash-join builder sees ONE concatenated fragment
    return t


def add_inner_to_sqlite(con: sqlite3.Connection):
    con.execute(f"CREATE TABLE test_inner({', '.join(n + ' bigint' for n in INNER_NAMES)})")
    con.executemany(f"INSERT INTO test_inner VALUES({','.join('?' * len(INNER_NAMES))})", INNER_ROWS)
    return con


JOIN_GOLDEN = [
    "SELECT COUNT(*) FROM test JOIN test_inner ON test.x = test_inner.x;",                                              # :12627 / :12857
    "SELECT test_inner.x, COUNT(*) FROM test JOIN test_inner ON test.x = test_inner.x GROUP BY test_inner.x ORDER BY 2;",   # :12633
    "SELECT COUNT(*) FROM test JOIN test_inner ON test.x = test_inner.xx;",                                             # :12713
    "SELECT test_inner.xx, COUNT(*) FROM test JOIN test_inner ON test.x = test_inner.xx GROUP BY test_inner.xx ORDER BY 2;",  # :12714
    "SELECT COUNT(*) FROM test LEFT JOIN test_inner ON test.x = test_inner.x WHERE test.y > 42;",                       # :13371
    "SELECT test.x, test_inner.x, COUNT(*) FROM test LEFT JOIN test_inner ON test.x = test_inner.x GROUP BY test.x, test_inner.x;",   # :13441 aggregated
    "SELECT test_inner.x, COUNT(*) FROM test LEFT JOIN test_inner ON test.x = test_inner.x WHERE test_inner.x IS NOT NULL GROUP BY test_inner.x;",   # :13448
    "SELECT test_inner.y, COUNT(*) FROM test LEFT JOIN test_inner ON test_inner.x = test.x WHERE test_inner.y = 43 GROUP BY test_inner.y ORDER BY 2 DESC;",   # :13509 (str = 'foo' is the y = 43 row)
    "SELECT test_inner.dt, COUNT(*), MAX(test_inner.ts), MIN(test.t) FROM test JOIN test_inner ON test.x = test_inner.x WHERE test_inner.dt16 = test_inner.dt32 GROUP BY test_inner.dt;",
]
# test_inner on its own (Select.ColumnWidths :8345-8349: DISTINCT == GROUP BY without aggregates)
INNER_GOLDEN = [
    "SELECT x, COUNT(*) FROM test_inner GROUP BY x ORDER BY 1;",
    "SELECT x, xx, y, COUNT(*) FROM test_inner GROUP BY x, xx, y ORDER BY 1, 2, 3;",
    "SELECT dt, COUNT(*), MIN(ts), MAX(dt16) FROM test_inner WHERE dt32 = dt16 GROUP BY dt;",   # 5575 day entries, two of them touched
]
