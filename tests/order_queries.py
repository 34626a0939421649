"""ORDER BY / LIMIT / OFFSET query sets shared by the CPU (oracle vs SQLite) and GPU (CUDA vs oracle) tests.

Every ORDER BY below ends in the GROUP BY key(s), so the order is total and the reference's unspecified tie order
(std::sort / std::partial_sort in ResultSet::topPermutationImpl, ResultSet.cpp:1501-1527) cannot matter."""
import numpy as np

from heavydb_b200 import abi

# on the reference's golden table `test` (ExecuteTest.cpp:30063-30115); adapted from ExecuteTest.cpp:2019, :2026,
# :2028, :2841, :2868, :2885 (aliases replaced by positions, NULL placement spelled out as the tests do at :2456-2460)
GOLDEN_ORDER_QUERIES = [
    "SELECT x, w, COUNT(*) FROM test GROUP BY x, w ORDER BY 1, 3, 2;",
    "SELECT x, MAX(dn) FROM test GROUP BY x ORDER BY 2 ASC NULLS FIRST, 1;",
    "SELECT COUNT(*), x, y, w FROM test GROUP BY x, y, w ORDER BY 1, 2, 3 ASC NULLS FIRST, 4;",
    "SELECT x, AVG(u), COUNT(*) FROM test GROUP BY x ORDER BY 3 DESC, 1;",
    "SELECT x, SUM(z) FROM test WHERE z <> 101 GROUP BY x ORDER BY x;",
    "SELECT smallint_nulls, COUNT(*) FROM test GROUP BY smallint_nulls ORDER BY smallint_nulls ASC NULLS FIRST;",
    "SELECT smallint_nulls, COUNT(*) FROM test GROUP BY smallint_nulls ORDER BY smallint_nulls DESC NULLS LAST;",
    "SELECT y, SUM(t), AVG(d) FROM test GROUP BY y ORDER BY 3 DESC, y ASC NULLS LAST LIMIT 2;",
    "SELECT y, COUNT(*) FROM test GROUP BY y ORDER BY y DESC NULLS FIRST LIMIT 1 OFFSET 1;",
    "SELECT z, COUNT(*), MIN(dn), MAX(w) FROM test GROUP BY z ORDER BY 3 DESC NULLS LAST, z;",
    "SELECT t, x, COUNT(*) FROM test GROUP BY t, x ORDER BY 3, t DESC, x;",
    "SELECT COUNT(*), SUM(x) FROM test ORDER BY 1;",
    "SELECT x, COUNT(*) FROM test GROUP BY x LIMIT 1;",          # LIMIT without ORDER BY: first entries in buffer order
    "SELECT x, COUNT(*) FROM test WHERE x > 100 GROUP BY x ORDER BY 2 DESC, 1 LIMIT 3;",   # empty result
    "SELECT x, COUNT(*) FROM test GROUP BY x ORDER BY 2 DESC LIMIT 0;",     # RelSort::isEmptyResult(): LIMIT 0 is an empty result, not "no limit"
    "SELECT y, SUM(t) FROM test GROUP BY y LIMIT 0 OFFSET 1;",
]

# on the mixed-type random table of test_gpu_parity.random_table
RAND_ORDER_QUERIES = [
    "SELECT k8, COUNT(*), SUM(a16) FROM r GROUP BY k8 ORDER BY 2 DESC, 1 ASC NULLS FIRST;",
    "SELECT k16, COUNT(*), MIN(a64), MAX(a64) FROM r WHERE k16 > 150 GROUP BY k16 ORDER BY 3 ASC NULLS LAST, 1 DESC NULLS LAST LIMIT 17;",
    "SELECT nn32, SUM(a32), AVG(a32) FROM r GROUP BY nn32 ORDER BY 3 DESC NULLS LAST, 1 LIMIT 25 OFFSET 10;",
    "SELECT nn64, AVG(d), MIN(d), MAX(d) FROM r GROUP BY nn64 ORDER BY 2 ASC NULLS FIRST, 1;",
    "SELECT nn64, SUM(dnn) FROM r WHERE a8 <> 3 GROUP BY nn64 ORDER BY 2 DESC, 1 LIMIT 5;",
    "SELECT k32, COUNT(*), SUM(a64) FROM r WHERE nn32 < 200 GROUP BY k32 ORDER BY 3 DESC NULLS LAST, 1 ASC NULLS LAST LIMIT 40;",
    "SELECT k64, COUNT(*), SUM(a32) FROM r GROUP BY k64 ORDER BY 2 DESC, 1 DESC NULLS FIRST LIMIT 100 OFFSET 3;",
    "SELECT sparse, COUNT(*), SUM(a64) FROM r GROUP BY sparse ORDER BY 2 DESC, 1 LIMIT 10;",                      # baseline hash
    "SELECT sparse, MAX(d) FROM r WHERE nn32 < 150 GROUP BY sparse ORDER BY 2 DESC NULLS LAST, sparse DESC;",    # baseline, key target from the key column
    "SELECT k8, nn64, COUNT(*), AVG(d) FROM r GROUP BY k8, nn64 ORDER BY 3 DESC, 1 ASC NULLS FIRST, 2 LIMIT 50;",  # composite key
    "SELECT nn64, k8, MIN(big) FROM r GROUP BY nn64, k8 ORDER BY 3 ASC NULLS LAST, 1, 2 DESC NULLS LAST;",
    "SELECT nn32, MIN(a8), MAX(a16) FROM r GROUP BY nn32 ORDER BY 2 ASC NULLS FIRST, 3 DESC, 1 LIMIT 1000;",
    "SELECT k16, COUNT(a8), COUNT(*) FROM r GROUP BY k16 ORDER BY 2, 3 DESC, 1 ASC NULLS FIRST LIMIT 100000 OFFSET 7;",
    "SELECT nn32, COUNT(*) FROM r GROUP BY nn32 ORDER BY 1 DESC LIMIT 3;",
]


# Reference quirk, reproduced on purpose: ORDER BY + OFFSET without LIMIT sorts with top_n = 0 + offset
# (RelAlgExecutor.cpp:3590-3594: top_n = get_limit_value(limit) + offset), keeps those `offset` rows, then drops them.
OFFSET_WITHOUT_LIMIT_QUIRK = "SELECT k16, COUNT(*) FROM r GROUP BY k16 ORDER BY 2, 1 ASC NULLS FIRST OFFSET 7;"


def rows_of(table: abi.Table, cols):
    """Python rows (None = NULL) of a host table, for SQLite."""
    arrays = [np.concatenate([f.host_cols[c] for f in table.fragments]) if table.fragments else np.zeros(0)
              for c in range(len(cols))]
    out = []
    for i in range(len(arrays[0]) if arrays else 0):
        r = []
        for c, (_, t, nn) in enumerate(cols):
            v = arrays[c][i]
            if not nn and v == abi.NULL_OF[t]:
                r.append(None)
            else:
                r.append(float(v) if t in (abi.kDOUBLE, abi.kFLOAT) else int(v))
        out.append(tuple(r))
    return out


def sqlite_sql(sql: str, unit: abi.BuiltUnit, table_name: str) -> str:
    """The same query for SQLite with every ORDER BY item's NULL placement made explicit (HeavyDB's default treats
    NULL as the largest value, SQLite's as the smallest; ExecuteTest.cpp:2456-2460 spells it out the same way)."""
    s = sql.rstrip(";")
    up = s.upper()
    if " ORDER BY " not in up:
        return s
    head, tail = s[:up.index(" ORDER BY ")], s[up.index(" ORDER BY ") + 10:]
    rest = ""
    for kw in (" LIMIT ", " OFFSET "):
        if kw in tail.upper():
            k = tail.upper().index(kw)
            tail, rest = tail[:k], tail[k:]
            break
    u = unit.unit
    items = []
    for i in range(u.num_order_entries):
        oe = u.order_entries[i]
        items.append(f"{oe.tle_no} {'DESC' if oe.is_desc else 'ASC'} NULLS {'FIRST' if oe.nulls_first else 'LAST'}")
    if " OFFSET " in rest.upper() and " LIMIT " not in rest.upper():
        rest = " LIMIT -1" + rest       # SQLite has no bare OFFSET
    return f"{head} ORDER BY {', '.join(items)}{rest}"
