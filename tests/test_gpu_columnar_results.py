"""ColumnarResults / Arrow hand-off of a finished result set (QueryEngine/ColumnarResults.cpp:256-392,
ArrowResultSetConverter.cpp): the columns must be exactly the rows of ResultSet iteration — and therefore of the
oracle — in order, in the target type's own width, NULLs as inline sentinels."""
import numpy as np
import pytest

import gpu_util as gu
import oracle_lib
import order_queries as oq
import sqlmini
import str_tables as stt
from heavydb_b200 import abi, executor
from test_gpu_parity import RAND_NAMES, RAND_QUERIES, random_table

pytestmark = pytest.mark.gpu


def columns_of_rows(rows, col_types):
    """Oracle rows (None = NULL) -> the arrays ColumnarResults would hold."""
    out = []
    for c, ty in enumerate(col_types):
        dt = np.dtype(abi.NUMPY_OF[ty])
        vals = [abi.NULL_OF[ty] if r[c] is None else r[c] for r in rows]
        out.append(np.array(vals, dtype=dt) if rows else np.empty(0, dtype=dt))
    return out


def check(rs):
    cols = rs.columnarResults(num_threads=4)
    types = [rs.getColType(i)[0] for i in range(rs.colCount())]
    assert [t for t, _, _ in cols] == types
    assert all(a.dtype == np.dtype(abi.NUMPY_OF[t]) for t, _, a in cols)
    got_rows = rs.rows()
    assert all(a.size == len(got_rows) == rs.rowCount() for _, _, a in cols)
    # 1. exactly the iteration of the result set itself, bit for bit
    for (ty, _, a), w in zip(cols, columns_of_rows(got_rows, types)):
        assert np.array_equal(a.view(np.uint8), w.view(np.uint8))
    # 2. ... which run_both / run_sorted have already compared with the oracle's rows
    # 3. Arrow: validity bitmap == "is the inline sentinel"
    batch = rs.toArrow(names=[f"c{i}" for i in range(len(cols))])
    assert batch.num_rows == len(got_rows) and batch.num_columns == len(cols)
    for i, (ty, _, a) in enumerate(cols):
        col = batch.column(i)
        assert col.null_count == int((a == abi.NULL_OF[ty]).sum())
        if a.size:
            assert np.array_equal(col.fill_null(abi.NULL_OF[ty]).to_numpy(zero_copy_only=False).astype(a.dtype), a)


def test_columnar_results_of_random_queries():
    table = random_table(30000, seed=77, frag_rows=8000)
    dev = gu.DeviceTable(table)
    ran = 0
    for sql in RAND_QUERIES:
        unit = sqlmini.parse(sql, table, RAND_NAMES)
        rs, ref = gu.run_both(unit, table, dev_table=dev, entry_guess=4001, has_card=True)
        check(rs)
        ran += 1
    assert ran >= 10


def test_columnar_results_follow_sort_offset_and_limit():
    from test_gpu_order_by import run_sorted
    table = random_table(30000, seed=78, frag_rows=8000)
    dev = gu.DeviceTable(table)
    for sql in oq.RAND_ORDER_QUERIES + [oq.OFFSET_WITHOUT_LIMIT_QUIRK]:
        unit = sqlmini.parse(sql, table, RAND_NAMES)
        rs, ref = run_sorted(unit, table, dev, entry_guess=3001, has_card=True)
        cols = rs.columnarResults()
        got = rs.rows()
        types = [rs.getColType(i)[0] for i in range(rs.colCount())]
        for (ty, _, a), w in zip(cols, columns_of_rows(got, types)):
            assert np.array_equal(a.view(np.uint8), w.view(np.uint8)), sql
        assert cols[0][2].size == rs.rowCount()


def test_columnar_results_of_dictionary_and_date_targets():
    table = stt.str_table(20000, seed=5, frag_rows=6000)
    dev = gu.DeviceTable(table)
    for sql in stt.STR_QUERIES:
        unit = sqlmini.parse(sql, table, stt.STR_NAMES)
        if unit.unit.num_order_entries:
            continue
        rs, ref = gu.run_both(unit, table, dev_table=dev)
        check(rs)


def test_columnar_results_of_an_empty_result():
    table = random_table(1000, seed=3, frag_rows=400)
    unit = sqlmini.parse("SELECT k8, COUNT(*), AVG(d) FROM r WHERE k32 < -1000000 GROUP BY k8;", table, RAND_NAMES)
    rs, ref = gu.run_both(unit, table)
    cols = rs.columnarResults()
    assert len(cols) == 3 and all(a.size == 0 for _, _, a in cols)
    assert rs.toArrow().num_rows == 0
