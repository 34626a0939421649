"""DECIMAL / NUMERIC columns on the CUDA path against the oracle (tests/dec_tables.py: the reference's golden `dd`
columns, ExecuteTest.cpp:1971-1986, :2823, :12022, and a table with NULLs and 32 / 16-bit FIXED decimal chunks) —
device-resident and host-resident chunks, row-wise and columnar output, and the scaled-integer read-out."""
import pytest

import dec_tables as dt
import gpu_util as gu
import oracle_lib
import sqlmini
from test_gpu_order_by import run_sorted

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("which", ["golden", "mixed"])
def test_decimal_queries_on_the_gpu(which):
    table = dt.make_table(dt.golden_rows()) if which == "golden" else dt.make_table(dt.mixed_rows(), fragment_size=170)
    dev = gu.DeviceTable(table)
    for sql in dt.GOLDEN_QUERIES + dt.MORE_QUERIES:
        unit = sqlmini.parse(sql, table, dt.DEC_NAMES)
        try:
            if unit.unit.num_order_entries:
                run_sorted(unit, table, dev)
            else:
                rs, ref = gu.run_both(unit, table, dev_table=dev)
                assert rs.rows(decimal_to_double=False) == ref.rows(decimal_to_double=False)
                gu.run_both(unit, table, device_resident=False)
                try:   # one columnar layout is refused on both sides (keyed output under a narrow first key)
                    oracle_lib.execute(unit, table, output_columnar=True)
                except oracle_lib.OracleError:
                    continue
                gu.run_both(unit, table, dev_table=dev, output_columnar=True)
        except Exception as e:
            raise AssertionError(f"query: {sql}\n{e}") from e


def test_decimal_columnar_results_keep_the_scaled_integers():
    """ColumnarResults reads rows with decimal_to_double = false (ColumnarResults.cpp:155, :550)."""
    import numpy as np
    from heavydb_b200 import abi
    table = dt.make_table(dt.golden_rows())
    unit = sqlmini.parse("SELECT dd, COUNT(*), SUM(dd), AVG(dd) FROM test GROUP BY dd;", table, dt.DEC_NAMES)
    rs, _ = gu.run_both(unit, table)
    cols = rs.columnarResults()
    assert [c[0] for c in cols] == [abi.kDECIMAL, abi.kINT, abi.kDECIMAL, abi.kDOUBLE]
    order = np.argsort(cols[0][2])
    assert cols[0][2][order].tolist() == [11110, 22220, 33330]
    assert cols[2][2][order].tolist() == [111100, 111100, 166650]
    assert np.allclose(cols[3][2][order], [111.1, 222.2, 333.3], rtol=1e-12)


def test_sort_over_wrapped_storage():
    """b2q_rs_create_from_storage + ResultSet::sort on the device: the ORACLE's buffer, wrapped, sorted by the CUDA radix path,
    against the oracle's own ResultSet::sort restatement (DECIMAL SUM / AVG and a nullable DECIMAL key among the order keys)."""
    import oracle_lib
    from heavydb_b200 import executor
    table = dt.make_table(dt.mixed_rows(n=3000, seed=3), fragment_size=800)
    unit = sqlmini.parse("SELECT dd, COUNT(*), SUM(p), AVG(q) FROM test GROUP BY dd;", table, dt.DEC_NAMES)
    for order, top_n, drop, keep in [([(3, True, False), (1, False, True)], 0, 0, 0),
                                     ([(2, True, False), (1, True, True)], 40, 5, 20),
                                     ([(4, False, True), (1, False, False)], 0, 17, 0),
                                     ([(1, True, True)], 25, 0, 25)]:
        ref = oracle_lib.execute(unit, table)
        rs = executor.Executor().resultSetFromStorage(ref.buffer(), unit, table)
        ref.sort(order, top_n)
        rs.sort(order, top_n)
        ref.drop_first_n(drop)
        ref.keep_first_n(keep)
        rs.dropFirstN(drop)
        rs.keepFirstN(keep)
        assert rs.entryCount() == ref.entry_count() and rs.rowCount() == ref.row_count()
        gu.rows_equal_ordered(rs.rows(), ref.rows())
        assert rs.rows(decimal_to_double=False) == ref.rows(decimal_to_double=False)


def test_decimal_attribute_of_a_joined_dimension_on_the_gpu():
    """A DECIMAL(7, 2) dimension attribute stored as FIXED(32), read through the join index (packed / 16-bit / shared-memory
    join-table variants are the planner's choice) as filter operand, group key and aggregate argument; INNER and LEFT."""
    fact, dim, *_ = dt.star_join()
    for sql in dt.JOIN_QUERIES:
        unit = sqlmini.parse(sql, fact, dt.FACT_NAMES, inner=(dim, dt.DIM_NAMES))
        try:
            rs, ref = gu.run_both(unit, fact)
            assert rs.rows(decimal_to_double=False) == ref.rows(decimal_to_double=False)
            gu.run_both(unit, fact, device_resident=False)
        except Exception as e:
            raise AssertionError(f"query: {sql}\n{e}") from e
