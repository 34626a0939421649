"""Star-schema tables for the one-level INNER hash join: a fact table probing a dimension table whose
key is unique, so that the reference builds a one-to-one PerfectJoinHashTable."""
import numpy as np

from heavydb_b200 import abi

FACT_COLS = [("fk32", abi.kINT, False), ("fk64", abi.kBIGINT, True), ("x", abi.kINT, True), ("v", abi.kBIGINT, False),
             ("d", abi.kDOUBLE, True), ("fk16", abi.kINT, False)]     # fk16: INT ENCODING FIXED(16)
FACT_ENC = [0, 0, 0, 0, 0, 2]
FACT_NAMES = [c[0] for c in FACT_COLS]
DIM_COLS = [("id32", abi.kINT, True), ("id64", abi.kBIGINT, False), ("attr", abi.kINT, False), ("attr8", abi.kTINYINT, True),
            ("w", abi.kDOUBLE, False), ("big", abi.kBIGINT, True)]
DIM_NAMES = [c[0] for c in DIM_COLS]
DIM_ROWS = 1000


def dim_table(seed=7, rows=DIM_ROWS):
    rng = np.random.default_rng(seed)
    id32 = rng.permutation(rows).astype(np.int32) + 3               # unique, NOT NULL, range [3, rows + 3)
    id64 = (rng.permutation(rows).astype(np.int64) * 2) - 100       # unique, sparse by 2, nullable
    id64[rng.random(rows) < 0.05] = abi.NULL_BIGINT                  # NULL inner keys never match
    attr = rng.integers(0, 20, rows).astype(np.int32)
    attr[rng.random(rows) < 0.1] = abi.NULL_INT
    w = rng.normal(0, 10, rows)
    w[rng.random(rows) < 0.1] = abi.NULL_DOUBLE
    cols = [id32, id64, attr, rng.integers(-100, 100, rows).astype(np.int8), w, rng.integers(-2**50, 2**50, rows).astype(np.int64)]
    t = abi.Table([(ty, nn) for _, ty, nn in DIM_COLS])
    t.add_host_fragment(cols)                                        # ONE concatenated fragment
    return t


def fact_table(n, seed, frag_rows, dim_rows=DIM_ROWS):
    rng = np.random.default_rng(seed)
    fk32 = rng.integers(-5, dim_rows + 60, n).astype(np.int32)       # some keys fall outside the dimension's range
    if n:
        fk32[rng.random(n) < 0.08] = abi.NULL_INT
    fk16 = rng.integers(0, dim_rows + 10, n).astype(np.int16)
    if n:
        fk16[rng.random(n) < 0.05] = -2**15                           # physical NULL of FIXED(16)
    v = rng.integers(-10**6, 10**6, n).astype(np.int64)
    if n:
        v[rng.random(n) < 0.1] = abi.NULL_BIGINT
    cols = [fk32, (rng.integers(-60, dim_rows + 20, n).astype(np.int64) * 2) - 100 + rng.integers(0, 2, n), rng.integers(0, 100, n).astype(np.int32),
            v, rng.random(n), fk16]
    t = abi.Table([(ty, nn) for _, ty, nn in FACT_COLS], encoded_sizes=FACT_ENC)
    for b in range(0, max(n, 1), frag_rows):
        t.add_host_fragment([c[b:b + frag_rows] for c in cols])
    return t


JOIN_QUERIES = [
    "SELECT d.attr, COUNT(*), SUM(t.v) FROM t JOIN d ON t.fk32 = d.id32 GROUP BY d.attr;",
    "SELECT COUNT(*), SUM(d.big), AVG(d.w), MIN(d.attr8), MAX(t.v) FROM t JOIN d ON t.fk32 = d.id32 WHERE t.x < 50;",
    "SELECT t.x, COUNT(*), SUM(d.attr8), AVG(t.d) FROM t JOIN d ON d.id32 = t.fk32 WHERE d.attr < 10 AND t.x > 5 GROUP BY t.x;",
    "SELECT d.attr8, COUNT(*), MIN(d.w), MAX(d.w), COUNT(d.w) FROM t JOIN d ON t.fk64 = d.id64 GROUP BY d.attr8;",     # int64 keys, NULL inner keys
    "SELECT d.attr, t.x, COUNT(*), SUM(t.v) FROM t JOIN d ON t.fk32 = d.id32 WHERE t.x < 30 GROUP BY d.attr, t.x;",     # composite key over both tables
    "SELECT d.attr, SUM(t.v), COUNT(t.v) FROM t JOIN d ON t.fk16 = d.id32 WHERE d.w IS NOT NULL AND NOT (d.attr8 BETWEEN -10 AND 10) GROUP BY d.attr;",  # FIXED(16) outer key
    "SELECT d.big, COUNT(*) FROM t JOIN d ON t.fk32 = d.id32 WHERE d.attr = 3 GROUP BY d.big;",                         # baseline hash on an inner column
    "SELECT d.attr, COUNT(*), AVG(d.w) FROM t JOIN d ON t.fk32 = d.id32 GROUP BY d.attr ORDER BY 2 DESC, 1 ASC NULLS FIRST LIMIT 5;",
    "SELECT COUNT(*) FROM t JOIN d ON t.fk32 = d.id32 WHERE d.attr IS NULL OR t.fk64 > 500;",
    "SELECT MIN(d.id64), MAX(d.id64), COUNT(d.id64), COUNT(*) FROM t JOIN d ON t.fk32 = d.id32;",
    "SELECT d.attr, COUNT(*) FROM t JOIN d ON t.fk32 = d.id32 WHERE t.x < d.attr8 OR d.w > t.d GROUP BY d.attr;",   # column OP column across the tables
]


# LEFT joins: an outer row without a match stays, its inner columns read NULL (codegenOuterJoinNullPlaceholder)
LEFT_JOIN_QUERIES = [
    "SELECT d.attr, COUNT(*), SUM(t.v), COUNT(d.id32) FROM t LEFT JOIN d ON t.fk32 = d.id32 GROUP BY d.attr;",
    "SELECT COUNT(*), COUNT(d.w), SUM(d.big), MIN(d.attr8), AVG(d.w), MAX(d.id64) FROM t LEFT JOIN d ON t.fk32 = d.id32 WHERE t.x < 50;",
    "SELECT t.x, COUNT(*), COUNT(d.attr8), SUM(d.attr8) FROM t LEFT JOIN d ON t.fk64 = d.id64 WHERE d.attr8 IS NULL OR d.attr8 > 0 GROUP BY t.x;",
    "SELECT d.attr8, COUNT(*), AVG(t.d) FROM t LEFT JOIN d ON t.fk16 = d.id32 GROUP BY d.attr8;",
    "SELECT d.attr, t.x, COUNT(*), MIN(d.w) FROM t LEFT JOIN d ON t.fk32 = d.id32 WHERE t.x < 20 GROUP BY d.attr, t.x;",
    "SELECT d.big, COUNT(*) FROM t LEFT JOIN d ON t.fk32 = d.id32 WHERE NOT (d.attr >= 2) OR d.attr IS NULL GROUP BY d.big;",        # baseline hash, NULL key
    "SELECT d.attr, COUNT(*), AVG(d.w) FROM t LEFT JOIN d ON t.fk32 = d.id32 GROUP BY d.attr ORDER BY 2 DESC, 1 ASC NULLS FIRST LIMIT 6;",
    "SELECT COUNT(*), COUNT(d.id32) FROM t LEFT JOIN d ON t.fk32 = d.id32 WHERE d.id32 IS NULL;",                                  # anti-join shape
]


def logical_rows(table, cols):
    out = []
    arrays = [np.concatenate([f.host_cols[c] for f in table.fragments]) for c in range(len(cols))]
    for i in range(len(arrays[0])):
        r = []
        for c, (_, t, nn) in enumerate(cols):
            v = arrays[c][i]
            if not nn and v == table.physical_null(c):
                r.append(None)
            else:
                r.append(float(v) if t == abi.kDOUBLE else int(v))
        out.append(tuple(r))
    return out
