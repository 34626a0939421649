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
    ("ts", abi.kTIMESTAMP, False, 4),      # TIMESTAMP ENCODING FIXED(32)
    ("dt", abi.kDATE, True, 0),            # DATE NOT NULL ENCODING NONE: int64 seconds on the day grid
    ("v", abi.kBIGINT, False, 0),
    ("d", abi.kDOUBLE, True, 0),
    ("dd", abi.kDATE, False, -4),          # DATE ENCODING DAYS(32) (the default for DATE): int32 days, NULL = INT32_MIN
    ("dd16", abi.kDATE, True, -2),         # DATE NOT NULL ENCODING DAYS(16): int16 days
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
        nulls(1_600_000_000 + rng.integers(0, 300, n).astype(np.int32), abi.NULL_INT),
        (18_000 + rng.integers(0, 40, n)).astype(np.int64) * 86400,
        nulls(rng.integers(-10**9, 10**9, n).astype(np.int64), abi.NULL_BIGINT),
        rng.random(n),
        nulls((18_000 + rng.integers(0, 60, n)).astype(np.int32), abi.NULL_INT),
        (10_000 + rng.integers(-15, 15, n)).astype(np.int16),
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
            if _e < 0:   # days-encoded DATE: the physical minimum is NULL, values decode to seconds
                r.append(None if v == table.physical_null(c) else int(v) * 86400)
            elif not nn and v == table.physical_null(c):
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
    "SELECT dt, COUNT(*), SUM(v) FROM s WHERE dt >= 1555286410 GROUP BY dt ORDER BY 2 DESC, 1 LIMIT 5;",   # narrowed min off the day grid
    # DATE keys carry the day bucket (ExpressionRange.cpp:622); days-encoded chunks decode as days * 86400
    "SELECT dd, COUNT(*), MIN(dd16), MAX(dd), COUNT(dd) FROM s GROUP BY dd;",
    "SELECT dd16, COUNT(*), SUM(v) FROM s WHERE dd > 1555286400 GROUP BY dd16;",
    "SELECT dd, dt, COUNT(*), MAX(dd16) FROM s WHERE dd <= 1557000000 GROUP BY dd, dt;",
    "SELECT COUNT(*), MIN(dd), MAX(dd16), COUNT(dd), MIN(dt) FROM s WHERE dd = 1555286400 OR dd16 <> 864000000 OR dd = 5;",
    "SELECT dt, COUNT(*), MIN(ts) FROM s WHERE dt < 1557100000 AND dt <> 1555372800 GROUP BY dt;",
    "SELECT s8, COUNT(*), COUNT(dd) FROM s WHERE dd < dd16 OR dd IS NULL OR NOT (dd16 >= 864000000) GROUP BY s8;",
    "SELECT dd, COUNT(*) FROM s WHERE x <> 2 GROUP BY dd ORDER BY 1 DESC LIMIT 7;",
    "SELECT dd16, MIN(d), AVG(v) FROM s WHERE dd16 BETWEEN 863308800 AND 864950400 AND dd IS NOT NULL GROUP BY dd16 ORDER BY 1;",
    # COUNT(DISTINCT) of dictionary ids (ExecuteTest.cpp:3921 `COUNT(distinct str)`), DICT(8) / DICT(16) chunks and a TIMESTAMP
    "SELECT x, COUNT(DISTINCT str), COUNT(DISTINCT s8) FROM s GROUP BY x;",
    "SELECT COUNT(DISTINCT str), COUNT(DISTINCT s16), COUNT(DISTINCT ts) FROM s WHERE s8 <> 7;",
]

STR_REJECTED = [
    "SELECT x, COUNT(DISTINCT dd) FROM s GROUP BY x;",           # COUNT(DISTINCT) of a days-encoded DATE (count_distinct_on_encoded_date_arg_)
    "SELECT x, SUM(DISTINCT v) FROM s GROUP BY x;",               # DISTINCT other than COUNT
    "SELECT COUNT(*) FROM s WHERE dd < dt;",                      # days-encoded vs seconds chunk
    "SELECT x, SUM(dd) FROM s GROUP BY x;",
    "SELECT s8, SUM(ts) FROM s GROUP BY s8;",
    "SELECT s8, MIN(str) FROM s GROUP BY s8;",
    "SELECT COUNT(*) FROM s WHERE str < 3;",
    "SELECT str, COUNT(*) FROM s GROUP BY str ORDER BY 1;",
    "SELECT AVG(dt) FROM s;",
]
