"""Product planner (b2q_plan: restated GroupByAndAggregate / QueryMemoryDescriptor decisions, host-only) against
the oracle's planner on the reference's golden table and on synthetic shapes of BASELINE.json's configs."""
import numpy as np
import pytest

import oracle_lib
import ref_tables as rt
import sqlmini
from heavydb_b200 import abi, executor
from test_oracle_golden import CONSTRAINED_NOT_NULL_PLANS, COUNT_DISTINCT_QUERIES, FLOAT_QUERIES, MULTI_KEY_QUERIES, NULL_LOGIC_QUERIES, PATH_QUERIES, REFERENCE_QUERIES

EXTRA = [
    "SELECT t, SUM(dn), AVG(dn), MIN(dn), MAX(dn), COUNT(dn) FROM test GROUP BY t;",
    "SELECT x, MIN(ufd), MAX(ufd), SUM(ufd) FROM test GROUP BY x;",
    "SELECT ofq, COUNT(*) FROM test GROUP BY ofq;",       # range too big for perfect hash -> baseline
    "SELECT ufq, COUNT(*), SUM(x) FROM test GROUP BY ufq;",
    # apply_int_qual's const_val +/- 1 at the ends of int64 (ExpressionRange.cpp:103-111 wraps): the range is left alone
    "SELECT ofq, COUNT(*) FROM test WHERE ofq > 9223372036854775807 GROUP BY ofq;",
    "SELECT ufq, COUNT(*) FROM test WHERE ufq < -9223372036854775808 GROUP BY ufq;",
] + COUNT_DISTINCT_QUERIES + FLOAT_QUERIES + ["SELECT y, MIN(fn), MAX(fn), AVG(fn) FROM test GROUP BY y;"] + [c[0] for c in CONSTRAINED_NOT_NULL_PLANS]   # `arg IS NOT NULL` quals: the plans are additionally pinned by hand in test_oracle_golden


@pytest.fixture(scope="module")
def table():
    return rt.make_table(rt.test_rows())


@pytest.mark.parametrize("bigint_count", [False, True])
@pytest.mark.parametrize("sql", REFERENCE_QUERIES + PATH_QUERIES + EXTRA + MULTI_KEY_QUERIES + NULL_LOGIC_QUERIES)
def test_plan_matches_oracle(table, sql, bigint_count):
    unit = sqlmini.parse(sql, table, rt.TEST_NAMES, bigint_count=bigint_count)
    ex = executor.Executor()
    eo = executor.execution_options(bigint_count=bigint_count)
    want = oracle_lib.plan(unit, table, entry_guess=48, has_card=True, bigint_count=bigint_count).as_dict()
    got = ex.plan(unit, table, eo=eo, max_groups_buffer_entry_guess=48, has_cardinality_estimation=True).as_dict()
    assert got == want


@pytest.mark.parametrize("sql", REFERENCE_QUERIES + PATH_QUERIES + EXTRA + MULTI_KEY_QUERIES)
def test_columnar_plan_matches_oracle(table, sql):
    """eo.output_columnar_hint: same layout decision (ResultSet.h:72-84 offsets) or the same refusal on both sides."""
    unit = sqlmini.parse(sql, table, rt.TEST_NAMES)
    eo = executor.execution_options(output_columnar_hint=True)
    kw = dict(max_groups_buffer_entry_guess=48, has_cardinality_estimation=True)
    try:
        want = oracle_lib.plan(unit, table, entry_guess=48, has_card=True, output_columnar=True).as_dict()
    except oracle_lib.OracleError as e:
        assert e.code == abi.ERR_UNSUPPORTED
        with pytest.raises(executor.UnsupportedOnThisPath):
            executor.Executor().plan(unit, table, eo=eo, **kw)
        return
    got = executor.Executor().plan(unit, table, eo=eo, **kw).as_dict()
    assert got == want and got["output_columnar"] == 1


def test_cardinality_estimation_required(table):
    unit = sqlmini.parse("SELECT ofq, COUNT(*) FROM test GROUP BY ofq;", table, rt.TEST_NAMES)
    with pytest.raises(executor.CardinalityEstimationRequired):
        executor.Executor().plan(unit, table)
    with pytest.raises(oracle_lib.OracleError) as ei:
        oracle_lib.plan(unit, table)
    assert ei.value.code == abi.ERR_CARDINALITY_ESTIMATION_REQUIRED


def test_unsupported_is_rejected_not_ignored(table):
    for field, val in (("num_join_quals", 2), ("has_union_all", 1), ("has_window_function", 1)):
        b = abi.UnitBuilder(table)
        b.target(b.agg(abi.kCOUNT))
        b.unsupported[field] = val
        with pytest.raises(executor.UnsupportedOnThisPath):
            executor.Executor().plan(b.build(), table)
    co = executor.compilation_options(device_type=abi.DEVICE_CPU)   # no CPU execution on this path
    b = abi.UnitBuilder(table)
    b.target(b.agg(abi.kCOUNT))
    with pytest.raises(executor.UnsupportedOnThisPath):
        executor.Executor().plan(b.build(), table, co=co)


def _config_table(kind):
    """Tiny stand-ins with the chunk statistics of BASELINE.json's configs (stats are what the planner reads)."""
    if kind == "C2":   # c0..c3 int64 in [0,1e6), g int32 in [0,1e4)
        t = abi.Table([(abi.kBIGINT, True)] * 4 + [(abi.kINT, True)])
        cols = [np.array([0, 999999], dtype=np.int64)] * 4 + [np.array([0, 9999], dtype=np.int32)]
        return t.add_host_fragment(cols), ["c0", "c1", "c2", "c3", "g"]
    if kind == "C3":   # f int64, g int32 in [0,256), v double not null
        t = abi.Table([(abi.kBIGINT, True), (abi.kINT, True), (abi.kDOUBLE, True)])
        return t.add_host_fragment([np.array([0, 999999]), np.array([0, 255]), np.array([0.0, 0.999])]), ["f", "g", "v"]
    if kind == "C4":   # key int64 in [0,1e7), v int64 in [0,1e6)
        t = abi.Table([(abi.kBIGINT, True), (abi.kBIGINT, True)])
        return t.add_host_fragment([np.array([0, 9999999]), np.array([0, 999999])]), ["key", "v"]
    raise ValueError(kind)


@pytest.mark.parametrize("kind,sql,expect", [
    ("C2", "SELECT g, SUM(c1), COUNT(*) FROM t WHERE c0 < 500000 GROUP BY g;",
     dict(query_desc_type=abi.GroupByPerfectHash, entry_count=10000, keyless_hash=1, idx_target_as_key=2, row_size=24)),
    ("C2", "SELECT g, SUM(c1), SUM(c2), SUM(c3), COUNT(*) FROM t WHERE c0 < 500000 GROUP BY g;",
     dict(query_desc_type=abi.GroupByPerfectHash, entry_count=10000, keyless_hash=1, idx_target_as_key=4, row_size=40)),
    ("C3", "SELECT g, AVG(v) FROM t WHERE f < 500000 GROUP BY g;",
     dict(query_desc_type=abi.GroupByPerfectHash, entry_count=256, keyless_hash=1, idx_target_as_key=2, row_size=24)),
    ("C4", "SELECT key, SUM(v) FROM t GROUP BY key;",
     dict(query_desc_type=abi.GroupByPerfectHash, entry_count=10000000, keyless_hash=0, row_size=24)),
])
def test_baseline_config_plans(kind, sql, expect):
    """the layouts the reference picks for BASELINE.json's configs."""
    table, names = _config_table(kind)
    unit = sqlmini.parse(sql, table, names)
    want = oracle_lib.plan(unit, table).as_dict()
    got = executor.Executor().plan(unit, table).as_dict()
    assert got == want
    for k, v in expect.items():
        assert got[k] == v, (k, got[k], v)


def test_remote_fragments_carry_stats_only():
    """Multi-GPU: a device is handed its own fragments plus the OTHER devices' fragments as chunk stats (col_buffers =
    NULL), so every device derives the plan of the whole table — position-aligned partial tables (include/b2q.h)."""
    from test_gpu_parity import RAND_NAMES, random_table
    table = random_table(4000, seed=2, frag_rows=500)      # 8 fragments with different per-fragment ranges
    view = abi.Table(table.col_types)
    for f in table.fragments:
        if f.fragment_id % 2 == 0:
            view.fragments.append(f)
        else:
            view.add_remote_fragment(f.num_tuples, f.stats, f.fragment_id)
    for sql in ["SELECT k32, COUNT(*), SUM(a64) FROM r GROUP BY k32;", "SELECT k8, k16, MIN(d) FROM r WHERE nn32 < 100 GROUP BY k8, k16;"]:
        unit = sqlmini.parse(sql, table, RAND_NAMES)
        whole = executor.Executor().plan(unit, table).as_dict()
        assert executor.Executor().plan(unit, view).as_dict() == whole
        local_only = abi.Table(table.col_types)
        local_only.fragments = [f for f in table.fragments if f.fragment_id % 2 == 0]
        assert executor.Executor().plan(unit, local_only).as_dict()["buffer_size"] > 0   # plans, but not necessarily the same ranges


def _stats_only_table(col_types, stats_per_col, tuples=1000, encoded_sizes=None):
    t = abi.Table(col_types, encoded_sizes=encoded_sizes)
    stats = []
    for lo, hi, has_nulls in stats_per_col:
        s = abi.ChunkStats()
        s.int_min, s.int_max, s.has_nulls = lo, hi, int(has_nulls)
        stats.append(s)
    t.add_remote_fragment(tuples, stats, fragment_id=0)
    return t


def _both_plans(unit, table, **kw):
    outs = []
    for side in ("oracle", "product"):
        try:
            if side == "oracle":
                outs.append(oracle_lib.plan(unit, table, entry_guess=kw.get("guess", 0), has_card=kw.get("has_card", False)).as_dict())
            else:
                outs.append(executor.Executor().plan(unit, table, max_groups_buffer_entry_guess=kw.get("guess", 0),
                                                     has_cardinality_estimation=kw.get("has_card", False)).as_dict())
        except oracle_lib.OracleError as e:
            outs.append(("error", e.code))
        except executor.CardinalityEstimationRequired:
            outs.append(("error", abi.ERR_CARDINALITY_ESTIMATION_REQUIRED))
    return outs


def test_dictionary_key_range_too_big_for_perfect_hash():
    """getColRangeInfo's string-key branch (GroupByAndAggregate.cpp:311-356): dictionary ids are dense, so a range past
    max_entry_count stays perfect hash unless a filter may thin it out; with filters and no sort, baseline when no
    estimate exists yet or 2 * estimate < range."""
    big = 60_000_000     # > 2^30 / (2 cols * 8)
    t = _stats_only_table([(abi.kTEXT, False), (abi.kINT, True)], [(0, big, True), (0, 9, False)])
    names = ["str", "x"]
    o, p = _both_plans(sqlmini.parse("SELECT str, COUNT(*) FROM t GROUP BY str;", t, names), t)
    assert o == p and o["query_desc_type"] == abi.GroupByPerfectHash and o["entry_count"] == big + 2
    unit = sqlmini.parse("SELECT str, COUNT(*) FROM t WHERE x < 3 GROUP BY str;", t, names)
    assert _both_plans(unit, t) == [("error", abi.ERR_CARDINALITY_ESTIMATION_REQUIRED)] * 2
    o, p = _both_plans(unit, t, guess=1000, has_card=True)
    assert o == p and o["query_desc_type"] == abi.GroupByBaselineHash and o["entry_count"] == 1000
    o, p = _both_plans(unit, t, guess=big, has_card=True)          # 2 * estimate >= range: perfect hash after all
    assert o == p and o["query_desc_type"] == abi.GroupByPerfectHash
    o, p = _both_plans(sqlmini.parse("SELECT str, COUNT(*) FROM t WHERE x < 3 GROUP BY str ORDER BY 2 DESC LIMIT 3;", t, names), t)
    assert o == p and o["query_desc_type"] == abi.GroupByPerfectHash
    # an integer key of the same range goes baseline whatever the filters
    ti = _stats_only_table([(abi.kINT, False), (abi.kINT, True)], [(0, big, True), (0, 9, False)])
    assert _both_plans(sqlmini.parse("SELECT k, COUNT(*) FROM t GROUP BY k;", ti, ["k", "x"]), ti) == [("error", abi.ERR_CARDINALITY_ESTIMATION_REQUIRED)] * 2


def test_date_keys_carry_the_day_bucket():
    """getLeafColumnRange gives DATE columns bucket = 86400 (ExpressionRange.cpp:622-624): perfect hash over
    (max - min) / 86400 + 1 entries, never keyless (QueryMemoryDescriptor.cpp:327-333), never 'too big' (:357)."""
    day = 86400
    for enc in (0, -4):
        t = _stats_only_table([(abi.kDATE, False), (abi.kINT, True)], [(18000 * day, 58000 * day, True), (0, 9, False)], encoded_sizes=[enc, 0])
        o, p = _both_plans(sqlmini.parse("SELECT d, COUNT(*) FROM t GROUP BY d;", t, ["d", "x"]), t)
        assert o == p
        assert (o["query_desc_type"], o["bucket"], o["entry_count"], o["keyless_hash"]) == (abi.GroupByPerfectHash, day, 40000 + 2, 0)
        # a simple qual narrows min off the day grid; the entry count follows (max - min) / bucket
        o, p = _both_plans(sqlmini.parse("SELECT d, COUNT(*) FROM t WHERE d > %d GROUP BY d;" % (57990 * day + 5), t, ["d", "x"]), t)
        assert o == p and o["min_val"] == 57990 * day + 6 and o["entry_count"] == 9 + 2
    # TIMESTAMP of the same range: no bucket, too big => baseline
    t = _stats_only_table([(abi.kTIMESTAMP, False), (abi.kINT, True)], [(18000 * day, 58000 * day, True), (0, 9, False)])
    assert _both_plans(sqlmini.parse("SELECT d, COUNT(*) FROM t GROUP BY d;", t, ["d", "x"]), t) == [("error", abi.ERR_CARDINALITY_ESTIMATION_REQUIRED)] * 2


def test_count_distinct_refusals_agree(table):
    """Descriptors the reference serves with a std::set (CPU only) are refused by both planners."""
    for sql in ["SELECT COUNT(DISTINCT d) FROM test;", "SELECT x, COUNT(DISTINCT ofq) FROM test GROUP BY x;", "SELECT SUM(DISTINCT x) FROM test;"]:
        try:
            unit = sqlmini.parse(sql, table, rt.TEST_NAMES)
        except Exception:
            continue
        with pytest.raises(oracle_lib.OracleError) as ei:
            oracle_lib.plan(unit, table, entry_guess=48, has_card=True)
        assert ei.value.code == abi.ERR_UNSUPPORTED
        with pytest.raises(executor.UnsupportedOnThisPath):
            executor.Executor().plan(unit, table, max_groups_buffer_entry_guess=48, has_cardinality_estimation=True)
