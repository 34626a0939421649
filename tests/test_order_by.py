"""ORDER BY / LIMIT / OFFSET — CPU side: the oracle's restatement of ResultSet::sort + dropFirstN /
keepFirstN (ResultSet.cpp:58-66, :781-849, :1310-1527) against SQLite, the way Tests/ExecuteTest.cpp compares
ordered results row by row (SQLiteComparator::compare_impl, ExecuteTest.cpp:383-520)."""
import pytest

import oracle_lib
import order_queries as oq
import ref_tables as rt
import sqlmini
from heavydb_b200 import abi, executor
from test_gpu_parity import RAND_COLS, RAND_NAMES, random_table


def assert_ordered_rows_match(ours, ref, eps=rt.EPS, fp_abs=0.0):
    assert len(ours) == len(ref), f"{len(ours)} rows vs {len(ref)}\nours={ours[:6]}\nref={ref[:6]}"
    for a, b in zip(ours, ref):
        assert len(a) == len(b)
        for va, vb in zip(a, b):
            if va is None or vb is None:
                assert va is None and vb is None, f"{a} vs {b}"
            elif isinstance(vb, float) or isinstance(va, float):
                assert va == vb or abs(va - vb) <= eps * abs(vb) + fp_abs, f"{a} vs {b}"
            else:
                assert va == vb, f"{a} vs {b}"


@pytest.fixture(scope="module")
def golden():
    rows = rt.test_rows()
    return rt.make_table(rows), rt.make_sqlite(rows)


@pytest.fixture(scope="module")
def rand():
    table = random_table(3000, seed=31, frag_rows=700)
    return table, rt.make_sqlite(oq.rows_of(table, RAND_COLS), RAND_COLS, "r")


@pytest.mark.parametrize("sql", oq.GOLDEN_ORDER_QUERIES)
def test_oracle_order_by_vs_sqlite_golden(golden, sql):
    table, con = golden
    unit = sqlmini.parse(sql, table, rt.TEST_NAMES)
    res = oracle_lib.execute(unit, table, entry_guess=64, has_card=True)
    ref = [tuple(r) for r in con.execute(oq.sqlite_sql(sql, unit, "test")).fetchall()]
    if unit.unit.num_order_entries:
        assert_ordered_rows_match(res.rows(), ref)
    else:   # LIMIT without ORDER BY: any `limit` rows of the unordered result
        full = [tuple(r) for r in con.execute(sql[:sql.upper().index(" LIMIT ")]).fetchall()]
        assert len(res.rows()) == len(ref) and all(r in full for r in res.rows())
    assert res.row_count() == len(ref)


@pytest.mark.parametrize("sql", oq.RAND_ORDER_QUERIES)
def test_oracle_order_by_vs_sqlite_random(rand, sql):
    table, con = rand
    unit = sqlmini.parse(sql, table, RAND_NAMES)
    res = oracle_lib.execute(unit, table, entry_guess=3001, has_card=True, num_threads=4)
    ref = [tuple(r) for r in con.execute(oq.sqlite_sql(sql, unit, "r")).fetchall()]
    assert_ordered_rows_match(res.rows(), ref)
    assert res.row_count() == len(ref)


def test_offset_without_limit_quirk(rand):
    table, _ = rand
    unit = sqlmini.parse(oq.OFFSET_WITHOUT_LIMIT_QUIRK, table, RAND_NAMES)
    res = oracle_lib.execute(unit, table)
    assert res.rows() == [] and res.row_count() == 0 and res.entry_count() == 7


def test_result_set_sort_api(rand):
    """ResultSet::sort / dropFirstN / keepFirstN called on a finished result set give what sort_info gives."""
    table, _ = rand
    sql = "SELECT nn32, SUM(a32), AVG(a32) FROM r GROUP BY nn32 ORDER BY 3 DESC NULLS LAST, 1 LIMIT 25 OFFSET 10;"
    want = oracle_lib.execute(sqlmini.parse(sql, table, RAND_NAMES), table).rows()
    res = oracle_lib.execute(sqlmini.parse(sql[:sql.index(" ORDER BY")] + ";", table, RAND_NAMES), table)
    res.sort([(3, True, False), (1, False, False)], top_n=35)
    res.drop_first_n(10)
    res.keep_first_n(25)
    assert res.rows() == want and res.row_count() == 25 and res.entry_count() == 35


def test_sort_info_is_validated(golden):
    table, _ = golden
    b = abi.UnitBuilder(table)
    b.group_by(0)
    b.target_col(0)
    b.target(b.agg(abi.kCOUNT))
    b.order_by(3)           # only two targets
    with pytest.raises(executor.QueryExecutionError):
        executor.Executor().plan(b.build(), table)
    with pytest.raises(oracle_lib.OracleError):
        oracle_lib.plan(b.build(), table)
