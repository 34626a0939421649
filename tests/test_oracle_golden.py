"""Pins the CPU oracle against the reference's own SQL golden tests: table `test`
(Tests/ExecuteTest.cpp:30063-30115) and the query strings of Select.FilterAndSimpleAggregation (:1885-2233),
Select.FilterAndGroupBy (:2815-2872), Select.GroupByKeylessAndNotKeyless (:3146-3169), with SQLite as the
comparator (ExecuteTest.cpp:383-520) exactly as the reference does."""
import pytest

import oracle_lib
import ref_tables as rt
import sqlmini
from heavydb_b200 import abi

# Query strings taken verbatim from Tests/ExecuteTest.cpp where the shape is inside the path's subset.
REFERENCE_QUERIES = [
    "SELECT COUNT(*) FROM test;",
    "SELECT COUNT(smallint_nulls), COUNT(*), COUNT(dn) FROM test;",
    "SELECT MIN(x) FROM test;",
    "SELECT MAX(x) FROM test;",
    "SELECT MIN(z) FROM test;",
    "SELECT MAX(z) FROM test;",
    "SELECT MIN(t) FROM test;",
    "SELECT MAX(t) FROM test;",
    "SELECT COUNT(*) FROM test WHERE x > 6 AND x < 8;",
    "SELECT COUNT(*) FROM test WHERE x > 6 AND x < 8 AND z > 100 AND z < 102;",
    "SELECT COUNT(*) FROM test WHERE x > 6 AND x < 8 OR (z > 100 AND z < 103);",
    "SELECT COUNT(*) FROM test WHERE x <> 7;",
    "SELECT COUNT(*) FROM test WHERE z <> 102;",
    "SELECT COUNT(*) FROM test WHERE t <> 1002;",
    "SELECT MIN(x) FROM test WHERE x = 7;",
    "SELECT MIN(z) FROM test WHERE z = 101;",
    "SELECT MIN(t) FROM test WHERE t = 1001;",
    "SELECT AVG(y) FROM test WHERE x > 6 AND x < 8;",
    "SELECT AVG(y) FROM test WHERE z > 100 AND z < 102;",
    "SELECT AVG(y) FROM test WHERE t > 1000 AND t < 1002;",
    "SELECT x, AVG(u), COUNT(*) FROM test GROUP BY x;",           # :2587 Select.GroupBy (ORDER BY dropped)
    "SELECT COUNT(*) FROM test WHERE d > 2.3;",
    "SELECT SUM(d) FROM test;",
    "SELECT SUM(dn) FROM test;",
    "SELECT MIN(dn) FROM test;",
    "SELECT MAX(dn) FROM test;",
    "SELECT AVG(dn) FROM test;",
]

# Multi-column GROUP BY (perfect hash over the cardinality product)
MULTI_KEY_QUERIES = [
    "SELECT x, y, COUNT(*) FROM test GROUP BY x, y;",                      # verbatim, Select.GroupBy ExecuteTest.cpp:2590
    "SELECT x, z, COUNT(*), SUM(t), MIN(d), MAX(dn), AVG(y) FROM test GROUP BY x, z;",
    "SELECT y, smallint_nulls, COUNT(*), SUM(x) FROM test GROUP BY y, smallint_nulls;",
    "SELECT w, x, y, COUNT(*), AVG(t) FROM test GROUP BY w, x, y;",
    "SELECT y, w, COUNT(ofq), SUM(ofd) FROM test WHERE x = 7 GROUP BY y, w;",
    "SELECT z, y, MIN(t), MAX(t) FROM test WHERE z > 0 GROUP BY z, y;",
    "SELECT t, w, COUNT(*), COUNT(u) FROM test GROUP BY t, w;",
    "SELECT x, y, z, w, SUM(t) FROM test GROUP BY x, y, z, w;",
    "SELECT SUM(t), x, COUNT(*), y FROM test GROUP BY x, y;",
]

# Same table, more shapes of the path (filter + GROUP BY + every aggregate, nullable keys and arguments).
PATH_QUERIES = [
    "SELECT x, COUNT(*) FROM test GROUP BY x;",
    "SELECT y, COUNT(*) FROM test GROUP BY y;",
    "SELECT z, COUNT(*), SUM(t), MIN(t), MAX(t), AVG(t) FROM test GROUP BY z;",
    "SELECT w, SUM(x), SUM(y), SUM(z) FROM test GROUP BY w;",
    "SELECT t, SUM(d), AVG(d), MIN(d), MAX(d) FROM test GROUP BY t;",
    "SELECT t, SUM(dn), AVG(dn), MAX(dn), COUNT(dn) FROM test GROUP BY t;",
    "SELECT smallint_nulls, COUNT(*), SUM(x) FROM test GROUP BY smallint_nulls;",
    "SELECT smallint_nulls, AVG(smallint_nulls) FROM test GROUP BY smallint_nulls;",
    "SELECT ofd, COUNT(*), MIN(ofd), MAX(ofd) FROM test GROUP BY ofd;",
    "SELECT x, COUNT(ofq), MIN(ofq), MAX(ofq) FROM test GROUP BY x;",
    "SELECT x, SUM(ofd), AVG(ofd), COUNT(ofd) FROM test GROUP BY x;",
    "SELECT x, SUM(u), MIN(u), MAX(u), COUNT(u) FROM test GROUP BY x;",
    "SELECT y, SUM(t) FROM test WHERE x > 7 GROUP BY y;",
    "SELECT y, SUM(t), COUNT(*) FROM test WHERE z < 102 GROUP BY y;",
    "SELECT y, SUM(t), COUNT(*) FROM test WHERE y < 43 GROUP BY y;",
    "SELECT y, SUM(t), COUNT(*) FROM test WHERE y >= 43 AND y <= 43 GROUP BY y;",
    "SELECT z, MAX(y) FROM test WHERE x = 7 GROUP BY z;",
    "SELECT x, MAX(z) FROM test WHERE z > -100 GROUP BY x;",
    "SELECT x, SUM(z) FROM test WHERE z <> 101 GROUP BY x;",
    "SELECT x, AVG(d) FROM test WHERE dn < -300.5 OR d <= 2.2 GROUP BY x;",
    "SELECT SUM(x), SUM(y), SUM(z), SUM(t), SUM(w) FROM test;",
    "SELECT MIN(w), MAX(w), MIN(y), MAX(y), MIN(smallint_nulls), MAX(smallint_nulls) FROM test;",
    "SELECT SUM(u), MIN(u), MAX(u), AVG(u), COUNT(u) FROM test;",
    "SELECT SUM(ofd), MIN(ofd), MAX(ofd), COUNT(ofd) FROM test WHERE x = 7;",
    "SELECT COUNT(*), SUM(t) FROM test WHERE x > 100;",
    "SELECT x, COUNT(*) FROM test WHERE x > 100 GROUP BY x;",
    "SELECT MIN(ufd), MAX(ufd) FROM test WHERE ufd > -2147483648;",
    "SELECT x, MAX(ufd), SUM(ufd) FROM test GROUP BY x;",
]


# NOT / IS [NOT] NULL / IN / BETWEEN (Analyzer::UOper kNOT, kISNULL; IN lists and BETWEEN as the OR / AND they expand to)
NULL_LOGIC_QUERIES = [
    "SELECT SUM(z) FROM test WHERE z IS NOT NULL;",                          # verbatim, ExecuteTest.cpp:1955
    "SELECT COUNT(*) FROM test WHERE u IS NOT NULL;",                        # :1989
    "SELECT COUNT(*) FROM test WHERE ofq >= 0 OR ofq IS NULL;",              # :2016
    "SELECT MAX(dn) FROM test WHERE dn IS NOT NULL;",                        # :2025
    "SELECT x, MAX(dn) FROM test WHERE dn IS NOT NULL GROUP BY x;",          # :2026 without the ORDER BY
    "SELECT x, SUM(z) FROM test WHERE z IS NOT NULL GROUP BY x;",            # :2868
    "SELECT COUNT(*) FROM test WHERE x IN (7, 8);",                          # :7364
    "SELECT COUNT(*) FROM test WHERE x IN (9, 10);",                         # :7365
    "SELECT COUNT(*) FROM test WHERE z IN (101, 102);",                      # :7366
    "SELECT COUNT(*) FROM test WHERE z IN (201, 202);",                      # :7367
    "SELECT t, COUNT(*) FROM test WHERE t NOT IN (1001, 1003, 1005, 1007, 1009, -10) GROUP BY t;",   # :2521
    "SELECT COUNT(*) FROM test WHERE x IS NULL;",                            # NOT NULL column: constant false
    "SELECT COUNT(*) FROM test WHERE x IS NOT NULL;",
    "SELECT COUNT(*), SUM(t) FROM test WHERE dn IS NULL OR y IS NULL;",
    "SELECT y, COUNT(*) FROM test WHERE NOT (y = 42 OR z > 101) GROUP BY y;",        # NULL y: NOT(NULL OR FALSE) = NULL -> dropped
    "SELECT COUNT(*) FROM test WHERE NOT (dn < -300.5 AND ofd IS NULL);",
    "SELECT COUNT(*) FROM test WHERE NOT (NOT (smallint_nulls IS NULL)) AND NOT x = 8;",
    "SELECT z, COUNT(*) FROM test WHERE z BETWEEN 100 AND 101 GROUP BY z;",
    "SELECT COUNT(*) FROM test WHERE z NOT BETWEEN 100 AND 101 OR w IN (-8);",
    "SELECT COUNT(*), MIN(ofd) FROM test WHERE NOT (ofd IN (1, 2, 3) OR ofd IS NULL);",
    "SELECT COUNT(*) FROM test WHERE d IS NULL OR NOT dn IS NOT NULL;",
    # column OP column (both sides cast to the common type; NULL on either side is not TRUE)
    "SELECT COUNT(*) FROM test WHERE x < y;",
    "SELECT COUNT(*), SUM(t) FROM test WHERE y >= z OR w = x;",
    "SELECT x, COUNT(*) FROM test WHERE z <> smallint_nulls AND NOT (d > dn) GROUP BY x;",
    "SELECT y, MIN(dn) FROM test WHERE dn < y OR t > ofq GROUP BY y;",        # double vs int, bigint vs bigint with NULLs
    "SELECT COUNT(*) FROM test WHERE NOT (x <> w) OR ufd <= ofd;",
]


# COUNT(DISTINCT column) on per-group bitmaps (init_count_distinct_descriptors, GroupByAndAggregate.cpp:650-855;
# agg_count_distinct_bitmap[_skip_val], RuntimeFunctions.cpp:366-376,1201-1210)
COUNT_DISTINCT_QUERIES = [
    "SELECT x, COUNT(DISTINCT x) FROM test GROUP BY x;",                      # verbatim, ExecuteTest.cpp:2781
    "SELECT x, y, COUNT(DISTINCT x) FROM test GROUP BY x, y;",                # :2783
    "SELECT COUNT(DISTINCT x) FROM test;",                                    # :3919
    "SELECT COUNT(*), MIN(x), MAX(x), AVG(y), SUM(z), COUNT(DISTINCT x) FROM test;",          # :3924 without the alias
    "SELECT y, COUNT(DISTINCT z) FROM test GROUP BY y;",
    "SELECT z, AVG(z), COUNT(DISTINCT z) FROM test GROUP BY z;",              # :3931 on the integer key
    "SELECT y, AVG(z), COUNT(DISTINCT x) FROM test GROUP BY y;",              # :3934 without HAVING
    "SELECT COUNT(DISTINCT y), COUNT(DISTINCT w), COUNT(DISTINCT smallint_nulls), COUNT(y) FROM test;",       # nullable arguments: NULLs are skipped
    "SELECT t, COUNT(DISTINCT y), COUNT(*), SUM(x) FROM test WHERE z > 0 GROUP BY t;",
    "SELECT x, COUNT(DISTINCT ofd), COUNT(DISTINCT u) FROM test GROUP BY x;",   # u: all NULL -> empty range -> 64-bit bitmap, count 0
    "SELECT ofq, COUNT(DISTINCT x), COUNT(*) FROM test WHERE ofq < 100 OR x > 100 GROUP BY ofq;",   # baseline-hash groups
    "SELECT x, COUNT(DISTINCT y) FROM test WHERE y < 43 GROUP BY x;",         # simple qual narrows the bitmap's range
]


# FLOAT columns (fixed_width_float_decode, DecodersImpl.h:112-123; agg_*_float on 4-byte slots, RuntimeFunctions.cpp:1491-1596)
FLOAT_QUERIES = [
    "SELECT COUNT(smallint_nulls), COUNT(*), COUNT(fn) FROM test;",          # verbatim, ExecuteTest.cpp:1890
    "SELECT MIN(ff) FROM test;",                                              # :1897
    "SELECT MIN(fn) FROM test;",                                              # :1898
    "SELECT SUM(ff) FROM test;",                                              # :1899
    "SELECT SUM(fn) FROM test;",                                              # :1900
    "SELECT MAX(f), MIN(f), AVG(f), AVG(fn), MAX(fn) FROM test;",
    "SELECT COUNT(*) FROM test WHERE f > 1.1;",
    "SELECT COUNT(*) FROM test WHERE f > 1.0 AND f < 1.2;",
    "SELECT COUNT(*), SUM(x) FROM test WHERE fn < -500 OR fn IS NULL;",
    "SELECT x, AVG(ff), COUNT(*) FROM test GROUP BY x;",                      # :2022 without the ORDER BY
    "SELECT x, MAX(fn) FROM test WHERE fn IS NOT NULL GROUP BY x;",           # :2023
    "SELECT x, SUM(f), MIN(ff), MAX(ff), COUNT(fn), SUM(fn) FROM test GROUP BY x;",
    "SELECT y, MIN(ff), MAX(fn), AVG(fn) FROM test GROUP BY y;",              # a group whose fn are all NULL (MIN(fn) first would
                                                                              # trip the keyless-MIN quirk, see test_reference_quirks)
    "SELECT z, SUM(ff), AVG(f) FROM test WHERE ff > 100 GROUP BY z;",
    "SELECT t, MIN(f), MAX(f) FROM test WHERE f <= d GROUP BY t;",            # float vs double column
    "SELECT ofq, SUM(f), COUNT(ff) FROM test WHERE ofq < 100 OR x > 100 GROUP BY ofq;",   # baseline-hash groups
    "SELECT x, y, MAX(ff), AVG(fn) FROM test GROUP BY x, y;",
]


@pytest.fixture(scope="module")
def env():
    rows = rt.test_rows()
    return rt.make_table(rows), rt.make_sqlite(rows)


@pytest.mark.parametrize("sql", REFERENCE_QUERIES + PATH_QUERIES + MULTI_KEY_QUERIES + NULL_LOGIC_QUERIES + COUNT_DISTINCT_QUERIES + FLOAT_QUERIES)
def test_oracle_vs_sqlite(env, sql):
    table, con = env
    unit = sqlmini.parse(sql, table, rt.TEST_NAMES)
    res = oracle_lib.execute(unit, table, entry_guess=64, has_card=True)
    ours = res.rows()
    ref = [tuple(r) for r in con.execute(sql.rstrip(";")).fetchall()]
    rt.assert_rows_match(ours, ref)
    assert res.row_count() == len(ref)


COLUMNAR_EXTRA = [   # 8-byte stored keys: keyed (non-keyless) perfect hash and baseline hash in the columnar layout
    "SELECT t, MIN(y), MAX(dn) FROM test GROUP BY t;",
    "SELECT t, z, MIN(y) FROM test GROUP BY t, z;",
    "SELECT ofq, COUNT(*), SUM(x), AVG(d) FROM test WHERE ofq < 100 OR x > 100 GROUP BY ofq;",   # baseline hash
    "SELECT ufq, MIN(y) FROM test WHERE ufq > 0 OR x > 100 GROUP BY ufq;",
]


@pytest.mark.parametrize("sql", REFERENCE_QUERIES + PATH_QUERIES + MULTI_KEY_QUERIES + COLUMNAR_EXTRA)
def test_oracle_columnar_vs_sqlite(env, sql):
    """--enable-columnar-output / the columnar hint (eo.output_columnar_hint): same answers from the columnar buffer
    (ResultSet.h:72-84), including the host reduce over the 2-row fragments."""
    table, con = env
    unit = sqlmini.parse(sql, table, rt.TEST_NAMES)
    try:
        res = oracle_lib.execute(unit, table, entry_guess=64, has_card=True, output_columnar=True)
    except oracle_lib.OracleError as e:
        # keyed layout whose first stored key is narrower than 8 bytes: refused (see oracle.cpp, make_plan)
        assert e.code == abi.ERR_UNSUPPORTED and "narrower than 8 bytes" in str(e)
        rw = oracle_lib.plan(unit, table, entry_guess=64, has_card=True)
        assert not rw.keyless_hash and rw.group_col_widths[0] < 8
        return
    p = res.plan
    assert p.output_columnar == 1
    ref = [tuple(r) for r in con.execute(sql.rstrip(";")).fetchall()]
    rt.assert_rows_match(res.rows(), ref)
    assert res.row_count() == len(ref)
    if p.query_desc_type != abi.NonGroupedAggregate:   # the layout itself: slot columns back to back, 8-byte aligned
        n, off = p.entry_count, (0 if p.keyless_hash else p.num_group_cols * ((8 * p.entry_count + 7) // 8 * 8))
        for s in range(p.num_slots):
            assert p.slot_offset[s] == off
            off += (p.slot_padded_width[s] * n + 7) // 8 * 8
        assert p.buffer_size == off


def test_multi_column_baseline_is_rejected():
    """A cardinality product above g_baseline_groupby_threshold (1e6, Execute.cpp:111) means baseline hash in the
    reference; multi-column baseline keys are outside this path and must be refused, not mis-executed."""
    table = rt.make_table(rt.test_rows())
    for sql in ["SELECT y, ofd, COUNT(*) FROM test GROUP BY y, ofd;", "SELECT smallint_nulls, z, COUNT(*) FROM test GROUP BY smallint_nulls, z;"]:
        unit = sqlmini.parse(sql, table, rt.TEST_NAMES)
        with pytest.raises(oracle_lib.OracleError) as ei:
            oracle_lib.execute(unit, table)
        assert ei.value.code == abi.ERR_UNSUPPORTED


def test_known_answers():
    """Hand-checkable values from the reference test comments: 15 rows have x = 7 (10 + 5)."""
    rows = rt.test_rows()
    table = rt.make_table(rows)
    unit = sqlmini.parse("SELECT COUNT(*) FROM test WHERE x > 6 AND x < 8;", table, rt.TEST_NAMES)
    assert oracle_lib.execute(unit, table).rows() == [(15,)]
    unit = sqlmini.parse("SELECT x, COUNT(*), SUM(t) FROM test GROUP BY x;", table, rt.TEST_NAMES)
    assert sorted(oracle_lib.execute(unit, table).rows()) == [(7, 15, 10 * 1001 + 5 * 1002), (8, 5, 5 * 1002)]


def test_reference_quirks():
    """Places where a faithful restatement of the reference differs from SQLite.  The oracle follows the
    reference's code, and the CUDA path must follow the oracle.

    (1) get_keyless_info's kMIN case (GroupByAndAggregate.cpp:561-589) does not look at has_nulls (kMAX does,
        :598-603): for a NULLABLE DOUBLE argument whose max is below NULL_DOUBLE (= DBL_MIN, i.e. all values
        negative) it picks MIN's slot as the "is this entry empty" marker, so a group whose arguments are all
        NULL keeps the init value and is reported as an empty entry — the row disappears.
    (2) A NOT NULL column that holds the type's NULL sentinel (ufd = -2147483648, the "underflow detection"
        column of the reference's own test table) reads back as NULL: makeTargetValue compares against the
        sentinel regardless of nullability (ResultSetIteration.cpp:2184-2188)."""
    rows = rt.test_rows()
    table = rt.make_table(rows)
    unit = sqlmini.parse("SELECT t, SUM(dn), AVG(dn), MIN(dn), MAX(dn), COUNT(dn) FROM test GROUP BY t;", table, rt.TEST_NAMES)
    res = oracle_lib.execute(unit, table)
    assert res.plan.keyless_hash == 1 and res.plan.idx_target_as_key == 4
    got = res.rows()
    assert len(got) == 1 and got[0][0] == 1002 and got[0][5] == 10      # the t = 1001 group (all dn NULL) is dropped
    # (1b) the same for a FLOAT argument, and wider: get_keyless_info reads the 4-byte init pattern (0x00800000 = NULL_FLOAT) as a
    #      DOUBLE (`*reinterpret_cast<const double*>(&init_max)`, GroupByAndAggregate.cpp:575-578) — a denormal of ~4e-317 — so any
    #      float argument whose maximum is <= 0 makes MIN the marker
    unit = sqlmini.parse("SELECT y, MIN(fn), MAX(fn), AVG(fn) FROM test GROUP BY y;", table, rt.TEST_NAMES)
    res = oracle_lib.execute(unit, table)
    assert res.plan.keyless_hash == 1 and res.plan.idx_target_as_key == 1
    got = res.rows()
    assert len(got) == 1 and got[0][0] == 43            # the y = 42 group (all fn NULL) is dropped
    unit = sqlmini.parse("SELECT x, MIN(ufd), MAX(ufd), SUM(ufd) FROM test GROUP BY x;", table, rt.TEST_NAMES)
    got = sorted(oracle_lib.execute(unit, table).rows(), key=lambda r: r[0])
    assert got == [(7, None, -1, 10 * -2147483648 + 5 * -1), (8, -2147483647, -2147483647, 5 * -2147483647)]


def test_empty_table():
    """test_empty (ExecuteTest.cpp:30117+): aggregates over no rows — COUNT 0, others NULL; GROUP BY yields no rows."""
    table = rt.make_table([])
    for sql, exp in [("SELECT COUNT(*) FROM test;", [(0,)]), ("SELECT SUM(x), MIN(y), MAX(t), AVG(d) FROM test;", [(None, None, None, None)])]:
        unit = sqlmini.parse(sql, table, rt.TEST_NAMES)
        assert oracle_lib.execute(unit, table).rows() == exp
    unit = sqlmini.parse("SELECT x, COUNT(*) FROM test GROUP BY x;", table, rt.TEST_NAMES)
    with pytest.raises(oracle_lib.OracleError) as ei:   # empty range => baseline => needs an estimate
        oracle_lib.execute(unit, table)
    assert ei.value.code == abi.ERR_CARDINALITY_ESTIMATION_REQUIRED
    assert oracle_lib.execute(unit, table, entry_guess=16, has_card=True).rows() == []


# constrained_not_null (OutputBufferInitialization.cpp:301-324): a top-level `arg IS NOT NULL` qual makes the aggregate's
# target NOT NULL for (1) get_keyless_info's kSUM case (GroupByAndAggregate.cpp:531), (2) the init values
# (OutputBufferInitialization.cpp:287) and (3) the skip-val choice (TargetExprBuilder.cpp:690).  Expected plans derived BY HAND
# from those three places and the golden table's values (z in {101, -78, 102}, t in {1001, 1002}, dn in {NULL, -2002.4, -220.6}).
#   sql, keyless, idx_target_as_key, init_vals, skip_null_val per target
I64_MIN, I64_MAX = -(1 << 63), (1 << 63) - 1
CONSTRAINED_NOT_NULL_PLANS = [
    # :2868 — z spans zero, so a NOT NULL SUM cannot mark emptiness: keyed layout, SUM starts at 0 and adds without a skip test
    ("SELECT x, SUM(z) FROM test WHERE z IS NOT NULL GROUP BY x;", 0, -1, [0, 0], [0, 0]),
    # the same without the qual: nullable argument without NULLs in the chunk stats -> keyless on SUM, NULL-sentinel init
    ("SELECT x, SUM(z) FROM test GROUP BY x;", 1, 1, [0, I64_MIN], [0, 1]),
    # t > 0 everywhere: NOT NULL SUM over a strictly positive range IS a marker; init 0
    ("SELECT x, SUM(t) FROM test WHERE t IS NOT NULL GROUP BY x;", 1, 1, [0, 0], [0, 0]),
    ("SELECT x, SUM(t) FROM test WHERE NOT (t IS NULL) AND x > 0 GROUP BY x;", 1, 1, [0, 0], [0, 0]),
    # the qual must be a top-level conjunct over the SAME column
    ("SELECT x, SUM(t) FROM test WHERE t IS NOT NULL OR x > 7 GROUP BY x;", 1, 1, [0, I64_MIN], [0, 1]),
    ("SELECT x, SUM(t) FROM test WHERE y IS NOT NULL GROUP BY x;", 1, 1, [0, I64_MIN], [0, 1]),
    # :2026 — kMAX's keyless test does not consult the quals (dn has NULLs -> keyed); init becomes -DBL_MAX, no skip test
    ("SELECT x, MAX(dn) FROM test WHERE dn IS NOT NULL GROUP BY x;", 0, -1, [0, "-dblmax"], [0, 0]),
    ("SELECT x, MAX(dn) FROM test GROUP BY x;", 0, -1, [0, "nulldbl"], [0, 1]),
    # MIN / COUNT / AVG over a constrained argument: 8-byte slot -> INT64_MAX; COUNT(z) counts every row
    ("SELECT x, MIN(z), COUNT(z), AVG(z) FROM test WHERE z IS NOT NULL GROUP BY x;", 1, 2, [0, I64_MAX, 0, 0, 0], [0, 0, 0, 0]),
    # non-grouped (:1955, :2025): set_notnull(target, false) wins for MIN / MAX / SUM / AVG; the code generator forces skip_null_val
    ("SELECT SUM(z) FROM test WHERE z IS NOT NULL;", 0, -1, [I64_MIN], [1]),
    ("SELECT MAX(dn) FROM test WHERE dn IS NOT NULL;", 0, -1, ["nulldbl"], [1]),
]


@pytest.mark.parametrize("case", CONSTRAINED_NOT_NULL_PLANS, ids=[c[0][:60] for c in CONSTRAINED_NOT_NULL_PLANS])
def test_constrained_not_null_plan(env, case):
    import struct
    sql, keyless, idx, inits, skips = case
    table, con = env
    unit = sqlmini.parse(sql, table, rt.TEST_NAMES)
    res = oracle_lib.execute(unit, table, entry_guess=64, has_card=True)
    p = res.plan
    bits = {"-dblmax": struct.unpack("<q", struct.pack("<d", -1.7976931348623157e308))[0],
            "nulldbl": struct.unpack("<q", struct.pack("<d", 2.2250738585072014e-308))[0]}
    assert p.keyless_hash == keyless
    if keyless:   # KeylessInfo::target_index only means something for a keyless layout
        assert p.idx_target_as_key == idx
    assert [p.init_vals[s] for s in range(p.num_slots)] == [bits.get(v, v) for v in inits]
    assert [p.targets[i].skip_null_val for i in range(p.num_targets)] == skips
    ref = [tuple(r) for r in con.execute(sql.rstrip(";")).fetchall()]
    rt.assert_rows_match(res.rows(), ref)


def test_slot_widths_by_hand(env):
    """pick_target_compact_width (QueryMemoryDescriptor.cpp:748-842) read line by line: 4-byte slots ONLY for a single-column GROUP BY
    whose targets are COUNT(*) and projections of integers of at most 4 bytes / dictionary strings, over at most 2^32 tuples, without
    g_bigint_count; `groupby_exprs.size() != 1 || !groupby_exprs.front()` (:775-778) keeps 8 bytes for non-grouped units AND for
    every multi-column GROUP BY — round 2 found planner and oracle both compacting the latter — any aggregate with an argument (:786-789)
    and any wider projection (:813) too.  Oracle and product planner, same answers."""
    table, _ = env
    cases = [
        ("SELECT x, COUNT(*) FROM test GROUP BY x;", False, [4, 4]),
        ("SELECT COUNT(*), x FROM test GROUP BY x;", False, [4, 4]),
        ("SELECT x, COUNT(*) FROM test GROUP BY x;", True, [8, 8]),               # g_bigint_count (:752-754)
        ("SELECT x, w, COUNT(*) FROM test GROUP BY x, w;", False, [8, 8, 8]),      # two GROUP BY columns (:775-778)
        ("SELECT w, x, z, COUNT(*) FROM test GROUP BY x, w, z;", False, [8, 8, 8, 8]),
        ("SELECT COUNT(*) FROM test;", False, [8]),                               # groupby_exprs == {nullptr}
        ("SELECT x, COUNT(y) FROM test GROUP BY x;", False, [8, 8]),              # an aggregate with an argument (:786-789)
        ("SELECT x, COUNT(*), MIN(w) FROM test GROUP BY x;", False, [8, 8, 8]),
        ("SELECT t, COUNT(*) FROM test GROUP BY t;", False, [8, 8]),              # BIGINT projection: not `no bigger than 4` (:813)
        ("SELECT z, COUNT(*) FROM test GROUP BY z;", False, [4, 4]),              # SMALLINT key
    ]
    for sql, bigint_count, widths in cases:
        unit = sqlmini.parse(sql, table, rt.TEST_NAMES, bigint_count=bigint_count)
        p = oracle_lib.plan(unit, table, entry_guess=64, has_card=True, bigint_count=bigint_count)
        assert list(p.slot_padded_width[: p.num_slots]) == widths, sql
        from heavydb_b200 import executor
        g = executor.Executor().plan(unit, table, eo=executor.execution_options(bigint_count=bigint_count),
                                     max_groups_buffer_entry_guess=64, has_cardinality_estimation=True)
        assert list(g.slot_padded_width[: g.num_slots]) == widths and g.as_dict() == p.as_dict(), sql


def test_count_distinct_descriptors_by_hand(env):
    """CountDistinctDescriptor{Bitmap, min_val, bitmap_sz_bits} as init_count_distinct_descriptors derives them from the golden
    table's chunk stats (x in [7, 8], y in [42, 43] with NULLs, z in [-78, 102], u all NULL), and what the reference refuses on a GPU."""
    table, _ = env
    def plan(sql):
        return oracle_lib.plan(sqlmini.parse(sql, table, rt.TEST_NAMES), table, entry_guess=64, has_card=True)
    p = plan("SELECT x, COUNT(DISTINCT x), COUNT(DISTINCT z), COUNT(*) FROM test GROUP BY x;")
    assert [(p.count_distinct_min[i], p.count_distinct_bits[i]) for i in range(4)] == [(0, 0), (7, 2), (-78, 181), (0, 0)]
    assert [p.targets[i].is_distinct for i in range(4)] == [0, 1, 1, 0]
    assert p.keyless_hash == 1 and p.idx_target_as_key == 3          # distinct targets never mark emptiness: COUNT(*) does
    assert list(p.slot_padded_width[:4]) == [8, 8, 8, 8] and [p.init_vals[i] for i in range(4)] == [0, 0, 0, 0]
    p = plan("SELECT COUNT(DISTINCT u), COUNT(DISTINCT y) FROM test WHERE y < 43;")
    assert [(p.count_distinct_min[i], p.count_distinct_bits[i]) for i in range(2)] == [(0, 64), (42, 1)]
    for sql in ["SELECT COUNT(DISTINCT d) FROM test;",            # fp argument: std::set implementation, CPU only (:3912-3913 run with dt = CPU)
                "SELECT x, COUNT(DISTINCT ofq) FROM test GROUP BY x;"]:   # 2^63-wide range: no bitmap
        with pytest.raises(oracle_lib.OracleError) as ei:
            plan(sql)
        assert ei.value.code == abi.ERR_UNSUPPORTED
