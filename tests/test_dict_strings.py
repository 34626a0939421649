"""Dictionary-encoded string keys and TIME-family columns: the oracle against SQLite on the
ids, the known answers of Tests/GroupByTest.cpp, and planner parity."""
import numpy as np
import pytest

import oracle_lib
import order_queries as oq
import ref_tables as rt
import sqlmini
import str_tables as st
from heavydb_b200 import abi, executor
from test_order_by import assert_ordered_rows_match


@pytest.fixture(scope="module")
def env():
    table = st.str_table(4000, seed=12, frag_rows=900)
    cols = [(n, t, nn) for n, t, nn, _ in st.STR_COLS]
    return table, rt.make_sqlite(st.logical_rows(table), cols, "s")


@pytest.mark.parametrize("sql", st.STR_QUERIES)
def test_oracle_vs_sqlite(env, sql):
    table, con = env
    unit = sqlmini.parse(sql, table, st.STR_NAMES)
    res = oracle_lib.execute(unit, table, num_threads=3)
    ref = [tuple(r) for r in con.execute(oq.sqlite_sql(sql, unit, "s")).fetchall()]
    if unit.unit.num_order_entries:
        assert_ordered_rows_match(res.rows(), ref)
    else:
        rt.assert_rows_match(res.rows(), ref)
    want = oracle_lib.plan(unit, table).as_dict()
    assert executor.Executor().plan(unit, table).as_dict() == want


@pytest.mark.parametrize("sql", st.STR_REJECTED)
def test_rejected_on_both_sides(env, sql):
    table, _ = env
    unit = sqlmini.parse(sql, table, st.STR_NAMES)
    with pytest.raises(oracle_lib.OracleError) as ei:
        oracle_lib.execute(unit, table)
    assert ei.value.code == abi.ERR_UNSUPPORTED
    with pytest.raises(executor.UnsupportedOnThisPath):
        executor.Executor().plan(unit, table)


def test_groupbytest_known_answers():
    """Tests/GroupByTest.cpp:60-152 PerfectHashNoFallback: rows (1,'hi'), (2,'bye'); SELECT COUNT(*) FROM t WHERE x = 1
    GROUP BY str  =>  one row, value 1; :264-338 BaselineNoFilters => two rows, each 1."""
    t = abi.Table([(abi.kINT, True), (abi.kTEXT, False)])
    t.add_host_fragment([np.array([1, 2], dtype=np.int32), np.array([0, 1], dtype=np.int32)])   # 'hi' -> id 0, 'bye' -> id 1
    unit = sqlmini.parse("SELECT COUNT(*) FROM t WHERE x = 1 GROUP BY str;", t, ["x", "str"])
    res = oracle_lib.execute(unit, t)
    assert res.plan.query_desc_type == abi.GroupByPerfectHash and res.rows() == [(1,)] and res.row_count() == 1
    unit = sqlmini.parse("SELECT COUNT(*) FROM t GROUP BY str;", t, ["x", "str"])
    assert oracle_lib.execute(unit, t).rows() == [(1,), (1,)]


def high_cardinality_str_table():
    """Tests/GroupByTest.cpp:160-171 + :192-194: table (x INT, str TEXT ENCODING DICT) with rows (1,'hi'), (2,'bye') whose cached
    range of `str` is forced to [0, 134217728] — one value more than the perfect-hash buffer limit allows (setup_str_col_caching)."""
    t = abi.Table([(abi.kINT, True), (abi.kTEXT, False)])
    t.add_host_fragment([np.array([1, 2], dtype=np.int32), np.array([0, 1], dtype=np.int32)])   # 'hi' -> id 0, 'bye' -> id 1
    st = t.fragments[0].stats[1]
    st.int_min, st.int_max, st.has_nulls = 0, 134217728, 0
    return t


def test_groupbytest_baseline_fallback():
    """Tests/GroupByTest.cpp:173-262 BaselineFallbackTest: COUNT(*) WHERE x = 1 GROUP BY str over a dictionary column whose range is
    too big for a perfect-hash buffer, with a filter and no sort => baseline hash (GroupByAndAggregate.cpp:311-356):
    executeWorkUnit(max_groups_buffer_entry_guess = 1, has_cardinality_estimation = false) throws CardinalityEstimationRequired, the
    same call with has_cardinality_estimation = true returns ONE row whose value is 1.  Both sides, plans equal."""
    t = high_cardinality_str_table()
    unit = sqlmini.parse("SELECT COUNT(*) FROM t WHERE x = 1 GROUP BY str;", t, ["x", "str"])
    with pytest.raises(oracle_lib.OracleError) as ei:
        oracle_lib.execute(unit, t, entry_guess=1, has_card=False)
    assert ei.value.code == abi.ERR_CARDINALITY_ESTIMATION_REQUIRED
    with pytest.raises(executor.CardinalityEstimationRequired):
        executor.Executor().plan(unit, t, max_groups_buffer_entry_guess=1, has_cardinality_estimation=False)
    res = oracle_lib.execute(unit, t, entry_guess=1, has_card=True)
    assert res.plan.query_desc_type == abi.GroupByBaselineHash and res.plan.entry_count == 1
    assert res.row_count() == 1 and res.rows() == [(1,)]
    got = executor.Executor().plan(unit, t, max_groups_buffer_entry_guess=1, has_cardinality_estimation=True).as_dict()
    assert got == res.plan.as_dict()
    # without the filter the same range stays perfect hash (dictionary ids are dense: :311-316), as BaselineNoFilters relies on
    unit2 = sqlmini.parse("SELECT COUNT(*) FROM t GROUP BY str;", t, ["x", "str"])
    assert oracle_lib.plan(unit2, t).query_desc_type == abi.GroupByPerfectHash


def test_dictionary_range_rule_with_a_sort():
    """GroupByAndAggregate.cpp:313-356 read branch by branch on the forced range [0, 134217728] (too big for a perfect-hash buffer) with
    a filter: no sort => baseline (estimate first); a sort keeps the original range => perfect hash; a sort AND a COUNT(DISTINCT)
    target => baseline all the same (:331-338).  Planner and oracle, same decisions."""
    t = high_cardinality_str_table()

    def decide(sql, has_card):
        unit = sqlmini.parse(sql, t, ["x", "str"])
        try:
            o = oracle_lib.plan(unit, t, entry_guess=4, has_card=has_card)
        except oracle_lib.OracleError as e:
            with pytest.raises(executor.QueryExecutionError) as ei:
                executor.Executor().plan(unit, t, max_groups_buffer_entry_guess=4, has_cardinality_estimation=has_card)
            assert ei.value.code == e.code
            return e.code
        g = executor.Executor().plan(unit, t, max_groups_buffer_entry_guess=4, has_cardinality_estimation=has_card)
        assert g.as_dict() == o.as_dict()
        return o.query_desc_type

    assert decide("SELECT COUNT(*) FROM t WHERE x >= 1 GROUP BY str;", False) == abi.ERR_CARDINALITY_ESTIMATION_REQUIRED
    assert decide("SELECT COUNT(*) FROM t WHERE x >= 1 GROUP BY str;", True) == abi.GroupByBaselineHash
    assert decide("SELECT COUNT(*), SUM(x) FROM t WHERE x >= 1 GROUP BY str ORDER BY 2;", False) == abi.GroupByPerfectHash
    assert decide("SELECT COUNT(*), COUNT(DISTINCT x) FROM t WHERE x >= 1 GROUP BY str ORDER BY 1;", False) == abi.ERR_CARDINALITY_ESTIMATION_REQUIRED
    assert decide("SELECT COUNT(*), COUNT(DISTINCT x) FROM t WHERE x >= 1 GROUP BY str ORDER BY 1;", True) == abi.GroupByBaselineHash
    assert decide("SELECT COUNT(*), COUNT(DISTINCT x) FROM t GROUP BY str ORDER BY 1;", False) in (abi.GroupByPerfectHash, abi.ERR_UNSUPPORTED)

