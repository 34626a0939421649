"""The product's read-out half — rowCount / getNextRow / isRowAtEmpty / getColType / ColumnarResults of libb2q — on the CPU:
`b2q_rs_create_from_storage` (ResultSet + allocateStorage, ResultSet.h:183-217, the way Tests/ResultSetTest.cpp wraps filled
storage) over the ORACLE's result buffer must read exactly what the oracle's own restatement of ResultSetIteration reads.
Row-wise and columnar, keyless / keyed / baseline-hash layouts, every column type incl. DECIMAL with both settings of
decimal_to_double.  (On a GPU box the same accessors run over the kernels' buffers: tests/test_gpu_*.py.)"""
import numpy as np
import pytest

import dec_tables as dt
import oracle_lib
import ref_tables as rt
import ref_time_table as tt
import sqlmini
import str_tables as stt
from heavydb_b200 import abi, executor
from test_filter_lowering import emu, run_program  # noqa: F401  (emu is a fixture)
from test_gpu_parity import RAND_NAMES, RAND_QUERIES, random_table
from test_oracle_golden import COLUMNAR_EXTRA, COUNT_DISTINCT_QUERIES, FLOAT_QUERIES, MULTI_KEY_QUERIES, NULL_LOGIC_QUERIES, PATH_QUERIES, REFERENCE_QUERIES


def _cases():
    yield "golden", rt.make_table(rt.test_rows()), rt.TEST_NAMES, list(REFERENCE_QUERIES) + list(MULTI_KEY_QUERIES) + list(PATH_QUERIES) + list(NULL_LOGIC_QUERIES) + list(COLUMNAR_EXTRA) + list(COUNT_DISTINCT_QUERIES) + list(FLOAT_QUERIES)
    yield "random", random_table(1500, seed=21, frag_rows=400), RAND_NAMES, list(RAND_QUERIES)
    yield "strings", stt.str_table(1200, seed=5, frag_rows=500), stt.STR_NAMES, list(stt.STR_QUERIES)
    yield "time", tt.make_table(tt.time_rows()), tt.TIME_NAMES, list(tt.TIME_QUERIES)
    yield "decimal", dt.make_table(dt.mixed_rows(), fragment_size=170), dt.DEC_NAMES, dt.GOLDEN_QUERIES + dt.MORE_QUERIES


def _strip_order(sql):
    up = sql.upper()
    return sql[:up.index(" ORDER BY ")] + ";" if " ORDER BY " in up else sql


@pytest.mark.parametrize("case", list(_cases()), ids=lambda c: c[0])
def test_product_readout_over_the_oracles_buffer(case):
    _, table, names, sqls = case
    ex = executor.Executor()
    ran = 0
    for sql in sqls:
        sql = _strip_order(sql)
        if " LIMIT " in sql.upper() or " OFFSET " in sql.upper():
            continue
        unit = sqlmini.parse(sql, table, names)
        for columnar in (False, True):
            try:
                ref = oracle_lib.execute(unit, table, entry_guess=6000, has_card=True, output_columnar=columnar)
            except oracle_lib.OracleError:
                continue
            eo = executor.execution_options(output_columnar_hint=columnar)
            rs = ex.resultSetFromStorage(ref.buffer(), unit, table, eo=eo, max_groups_buffer_entry_guess=6000,
                                         has_cardinality_estimation=True)
            ctx = (sql, columnar)
            assert rs.getQueryMemDesc().as_dict() == ref.plan.as_dict(), ctx
            assert rs.entryCount() == ref.entry_count() and rs.rowCount() == ref.row_count() and rs.colCount() == ref.col_count(), ctx
            assert [rs.getColType(i) for i in range(rs.colCount())] == [ref.col_type(i) for i in range(ref.col_count())], ctx
            for d2d in (True, False):
                assert rs.rows(decimal_to_double=d2d) == ref.rows(decimal_to_double=d2d), ctx
            n = rs.entryCount()
            L = oracle_lib.lib()
            step = max(1, n // 2000)
            for e in list(range(0, n, step)) + [n, n + 7]:
                want = bool(L.oracle_result_is_row_at_empty(ref.h, e)) if e < n else True
                assert rs.isRowAtEmpty(e) == want, (ctx, e)
            # ColumnarResults: one array per target in the target type's width, scaled integers for DECIMALs, NULL sentinels inline
            rows = ref.rows(decimal_to_double=False)
            cols = rs.columnarResults(num_threads=3)
            assert len(cols) == rs.colCount()
            for c, (ty, _nn, arr) in enumerate(cols):
                assert arr.size == len(rows), ctx
                null = abi.NULL_OF[ty]
                want = np.array([null if r[c] is None else r[c] for r in rows], dtype=abi.NUMPY_OF[ty])
                assert np.array_equal(arr, want), (ctx, c)
            ran += 1
    assert ran >= 10


def test_refusals():
    table = rt.make_table(rt.test_rows())
    unit = sqlmini.parse("SELECT x, COUNT(*) FROM test GROUP BY x;", table, rt.TEST_NAMES)
    ref = oracle_lib.execute(unit, table)
    with pytest.raises(executor.QueryExecutionError):
        executor.Executor().resultSetFromStorage(ref.buffer()[:-8], unit, table)       # not the descriptor's size
    rs = executor.Executor().resultSetFromStorage(ref.buffer(), unit, table)
    assert sorted(rs.rows()) == [(7, 15), (8, 5)]


def test_arrow_hand_off_on_the_cpu():
    """ResultSet.toArrow over wrapped storage: validity bitmaps mark the inline sentinels, DECIMALs become decimal128(…, scale)
    fed from the scaled integers (ArrowResultSetConverter.cpp:1141, :1425-1440)."""
    import decimal
    import pyarrow as pa
    table = dt.make_table(dt.mixed_rows(), fragment_size=170)
    unit = sqlmini.parse("SELECT dd, COUNT(*), SUM(p), MIN(q), AVG(dd_notnull), MAX(y) FROM test GROUP BY dd;", table, dt.DEC_NAMES)
    ref = oracle_lib.execute(unit, table)
    rs = executor.Executor().resultSetFromStorage(ref.buffer(), unit, table)
    batch = rs.toArrow(names=["dd", "n", "sp", "mq", "avg", "my"])
    assert [str(f.type) for f in batch.schema] == ["decimal128(19, 2)", "int32", "decimal128(19, 3)", "decimal128(19, 2)", "double", "int32"]
    rows = ref.rows(decimal_to_double=False)
    assert batch.num_rows == len(rows)
    scales = {0: 2, 2: 3, 3: 2}
    for c in range(6):
        got = batch.column(c).to_pylist()
        for g, r in zip(got, rows):
            if r[c] is None:
                assert g is None
            elif c in scales:
                assert isinstance(g, decimal.Decimal) and g == decimal.Decimal(r[c]).scaleb(-scales[c])
            else:
                assert g == r[c]
    assert any(r[0] is None for r in rows) and any(r[2] is not None and r[2] < 0 for r in rows)   # NULL key group, negative sums


@pytest.mark.parametrize("seed", range(2))
def test_host_side_end_to_end_against_sqlite(emu, seed):  # noqa: F811
    """Everything of the product except the CUDA kernels, against an independent engine: SQL -> b2q_plan -> the lowered
    program read on the host (tests/cpp/filter_emulator.cpp stands in for the kernels) -> b2q_rs_create_from_storage ->
    getNextRow, compared with SQLite like the reference's SQLiteComparator."""
    import random
    import order_queries as oq
    from test_gpu_fuzz import rand_query
    from test_gpu_parity import RAND_COLS
    from test_oracle_fuzz import known_reference_quirk, sqlite_overflows
    rng = random.Random(31000 + seed)
    table = random_table([900, 2500][seed], seed=700 + seed, frag_rows=[250, 2500][seed])
    con = rt.make_sqlite(oq.rows_of(table, RAND_COLS), RAND_COLS, "r")
    ex = executor.Executor()
    checked = 0
    for i in range(90):
        sql = rand_query(rng, multi_key=(i % 3 == 0))
        if sqlite_overflows(sql):
            continue
        unit = sqlmini.parse(sql, table, RAND_NAMES)
        try:
            plan = ex.plan(unit, table, max_groups_buffer_entry_guess=6000, has_cardinality_estimation=True)
        except executor.QueryExecutionError:
            continue
        if plan.query_desc_type not in (abi.GroupByPerfectHash, abi.NonGroupedAggregate) or known_reference_quirk(unit, plan):
            continue
        rc, buf = run_program(emu, unit, table, entry_guess=6000, has_card=True)
        assert rc == 0, (sql, rc)
        rs = ex.resultSetFromStorage(buf, unit, table, max_groups_buffer_entry_guess=6000, has_cardinality_estimation=True)
        ref = [tuple(r) for r in con.execute(sql.rstrip(";")).fetchall()]
        try:
            rt.assert_rows_match(rs.rows(), ref)
        except AssertionError as e:
            raise AssertionError(f"query: {sql}\n{e}") from e
        checked += 1
    assert checked >= 40
