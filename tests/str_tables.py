"""Tables with dictionary-encoded string columns (int32 / uint8 / uint16 ids) and TIME-family columns — the column
kinds of the reference's own boundary test (Tests/GroupByTest.cpp:60-63 groups by a dictionary string)."""
import numpy as np

from heavydb_b200 import abi

STR_COLS = [
    # name, logical type, notnull, encoded physical bytes (0 = none)
    ("x", abi.kINT, True, 0),
    ("str", abi.kTEXT, False, 0),          # TEXT ENCODING DICT(32), nullable
    ("s8", abi.kTEXT, False, 1),           # TEXT ENCODING DICT(8): uint8 ids, NULL = 255
    ("s16", abi.kVARCHAR, True, 2),        # VARCHAR(n) NOT NULL ENCODING DICT(16): uint16 ids
    ("ts", abi.kTIMESTAMP, False, 0),
    ("dt", abi.kDATE, True, 4),            # DATE NOT NULL ENCODING FIXED(32)
    ("v", abi.kBIGINT, False, 0),
    ("d", abi.kDOUBLE, True, 0),
]
STR_NAMES = [c[0] for c in STR_COLS]


def str_table(n, seed, frag_rows):
    rng = np.random.default_rng(seed)

    def nulls(a, null, p=0.12):
        a = a.copy()
        if n:
            a[rng.random(n) < p] = null
        return a
    cols = [
        rng.integers(1, 4, n).astype(np.int32),
        nulls(rng.integers(0, 50, n).astype(np.int32), abi.NULL_INT),
        nulls(rng.integers(0, 230, n).astype(np.uint8), 255),             # ids above 127: the unsigned decode matters
        rng.integers(0, 40000, n).astype(np.uint16),
        nulls(1_600_000_000 + rng.integers(0, 300, n).astype(np.int64), abi.NULL_BIGINT),
        (18_000 + rng.integers(0, 40, n)).astype(np.int32) * 86400 // 86400 * 86400 // 86400 + 1_555_000_000,
        nulls(rng.integers(-10**9, 10**9, n).astype(np.int64), abi.NULL_BIGINT),
        rng.random(n),
    ]
    t = abi.Table([(ty, nn) for _, ty, nn, _ in STR_COLS], encoded_sizes=[e for *_, e in STR_COLS])
    for b in range(0, max(n, 1), frag_rows):
        t.add_host_fragment([c[b:b + frag_rows] for c in cols])
    return t


def logical_rows(table):
    """Python rows (None = NULL) with logical values, for SQLite."""
    out = []
    ncol = len(STR_COLS)
    arrays = [np.concatenate([f.host_cols[c] for f in table.fragments]) for c in range(ncol)]
    for i in range(len(arrays[0])):
        r = []
        for c, (_, t, nn, _e) in enumerate(STR_COLS):
            v = arrays[c][i]
            if not nn and v == table.physical_null(c):
                r.append(None)
            else:
                r.append(float(v) if t == abi.kDOUBLE else int(v))
        out.append(tuple(r))
    return out


STR_QUERIES = [
    "SELECT str, COUNT(*) FROM s WHERE x = 1 GROUP BY str;",       # the shape of GroupByTest.cpp:100-130
    "SELECT COUNT(*) FROM s WHERE x = 1 GROUP BY str;",            # verbatim target list of PerfectHashNoFallback
    "SELECT s8, COUNT(*), SUM(v), COUNT(str) FROM s GROUP BY s8;",
    "SELECT s16, MIN(ts), MAX(ts), COUNT(ts) FROM s WHERE s8 <> 7 GROUP BY s16;",
    "SELECT ts, COUNT(*), AVG(d) FROM s WHERE str = 3 GROUP BY ts;",
    "SELECT dt, s8, COUNT(*) FROM s GROUP BY dt, s8;",
    "SELECT COUNT(str), COUNT(s8), MIN(dt), MAX(ts), COUNT(*) FROM s WHERE ts > 1600000050;",
    "SELECT str, s8, MIN(v), MAX(v) FROM s WHERE s16 <> 5 GROUP BY str, s8;",
    "SELECT s8, COUNT(*) FROM s WHERE s8 = 200 OR s8 = 3 OR str = 7 GROUP BY s8;",
    "SELECT dt, COUNT(*), SUM(v) FROM s WHERE dt >= 1555000010 GROUP BY dt ORDER BY 2 DESC, 1 LIMIT 5;",
]

STR_REJECTED = [
    "SELECT s8, SUM(ts) FROM s GROUP BY s8;",
    "SELECT s8, MIN(str) FROM s GROUP BY s8;",
    "SELECT COUNT(*) FROM s WHERE str < 3;",
    "SELECT str, COUNT(*) FROM s GROUP BY str ORDER BY 1;",
    "SELECT AVG(dt) FROM s;",
]
