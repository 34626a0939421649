"""Golden vectors of the reference's `test` table for the TIME-family and encoded columns (ExecuteTest.cpp:2040-2050,
:2844, :5057, :5317, :27956-27998): the oracle against SQLite like SQLiteComparator, and the product's planner
against the oracle's."""
import pytest

import oracle_lib
import order_queries as oq
import ref_tables as rt
import ref_time_table as tt
import sqlmini
from heavydb_b200 import abi, executor
from test_order_by import assert_ordered_rows_match


@pytest.fixture(scope="module")
def env():
    rows = tt.time_rows()
    return tt.make_table(rows), tt.make_sqlite(rows)


@pytest.mark.parametrize("sql", tt.TIME_QUERIES)
def test_time_columns_vs_sqlite(env, sql):
    table, con = env
    unit = sqlmini.parse(sql, table, tt.TIME_NAMES)
    res = oracle_lib.execute(unit, table, num_threads=2)
    ref = [tuple(r) for r in con.execute(oq.sqlite_sql(sql, unit, "test")).fetchall()]
    if unit.unit.num_order_entries:
        assert_ordered_rows_match(res.rows(), ref)
    else:
        rt.assert_rows_match(res.rows(), ref)
    assert executor.Executor().plan(unit, table).as_dict() == oracle_lib.plan(unit, table).as_dict()


def test_known_answers(env):
    """Counts that follow from the three row templates (10 / 5 / 5 rows; the middle one has NULL dates)."""
    table, _ = env
    def count(sql):
        return oracle_lib.execute(sqlmini.parse(sql, table, tt.TIME_NAMES), table).rows()[0][0]
    assert count(f"SELECT COUNT(*) FROM test WHERE o1 > {tt.D_1999_09_08};") == 15
    assert count(f"SELECT COUNT(*) FROM test WHERE o1 <= {tt.D_1999_09_08};") == 0
    assert count("SELECT COUNT(*) FROM test WHERE o1 = o2;") == 15
    assert count("SELECT COUNT(*) FROM test WHERE o1 <> o2;") == 0
    assert count("SELECT COUNT(*) FROM test WHERE fx IS NULL;") == 5
    res = oracle_lib.execute(sqlmini.parse(f"SELECT o, COUNT(*) FROM test WHERE o <= {tt.D_1999_09_09} GROUP BY o ORDER BY 2;", table, tt.TIME_NAMES), table)
    assert res.rows() == [(tt.D_1999_09_09, 15)]
    assert res.plan.query_desc_type == abi.GroupByPerfectHash and res.plan.bucket == 86400


def test_verbatim_strings_fold_to_the_same_queries(env):
    """The quoted-literal strings of ExecuteTest.cpp:2040-2050, folded like the analyzer folds them, are the queries above."""
    table, con = env
    for i, sql in enumerate(tt.VERBATIM):
        folded = tt.fold_time_literals(sql)
        assert "'" not in folded
        unit = sqlmini.parse(folded, table, tt.TIME_NAMES)
        got = oracle_lib.execute(unit, table).rows()
        assert got == [tuple(r) for r in con.execute(folded.rstrip(";")).fetchall()]
    assert tt.fold_time_literals("SELECT COUNT(*) FROM test WHERE m > '2014-12-13 22:23:15' AND n = '15:13:14';") == \
        f"SELECT COUNT(*) FROM test WHERE m > {tt.TS_A} AND n = {tt.T_151314};"
