"""Golden join vectors of the reference (`test` JOIN `test_inner`, Select.Joins_* of Tests/ExecuteTest.cpp): the oracle
against SQLite, the product's planner against the oracle's."""
import pytest

import oracle_lib
import order_queries as oq
import ref_join_tables as rj
import ref_tables as rt
import sqlmini
from heavydb_b200 import executor
from test_order_by import assert_ordered_rows_match


@pytest.fixture(scope="module")
def env():
    rows = rt.test_rows()
    return rt.make_table(rows), rj.inner_table(), rj.add_inner_to_sqlite(rt.make_sqlite(rows))


def _check(res, con, sql, unit, name):
    ref = [tuple(r) for r in con.execute(oq.sqlite_sql(sql, unit, name)).fetchall()]
    if unit.unit.num_order_entries:
        assert_ordered_rows_match(res.rows(), ref)
    else:
        rt.assert_rows_match(res.rows(), ref)


@pytest.mark.parametrize("sql", rj.JOIN_GOLDEN)
def test_join_golden(env, sql):
    test, inner, con = env
    unit = sqlmini.parse(sql, test, rt.TEST_NAMES, inner=(inner, rj.INNER_NAMES))
    res = oracle_lib.execute(unit, test, num_threads=2)
    _check(res, con, sql, unit, "test")
    assert executor.Executor().plan(unit, test).as_dict() == oracle_lib.plan(unit, test).as_dict()


@pytest.mark.parametrize("sql", rj.INNER_GOLDEN)
def test_inner_table_golden(env, sql):
    _, inner, con = env
    unit = sqlmini.parse(sql, inner, rj.INNER_NAMES)
    res = oracle_lib.execute(unit, inner)
    _check(res, con, sql, unit, "test_inner")
    assert executor.Executor().plan(unit, inner).as_dict() == oracle_lib.plan(unit, inner).as_dict()


def test_known_answers(env):
    """20 rows of `test` (x = 7 fifteen times, x = 8 five times) against inner keys {7, -9}."""
    test, inner, _ = env
    def rows(sql):
        return oracle_lib.execute(sqlmini.parse(sql, test, rt.TEST_NAMES, inner=(inner, rj.INNER_NAMES)), test).rows()
    assert rows(rj.JOIN_GOLDEN[0]) == [(15,)]
    assert rows(rj.JOIN_GOLDEN[1]) == [(7, 15)]
    assert rows(rj.JOIN_GOLDEN[4]) == [(10,)]       # y > 42: the 5 + 5 rows with y = 43, matched or not
