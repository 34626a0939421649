"""The reference's golden `test` table, TIME-family / encoded columns (tests/ref_time_table.py: ExecuteTest.cpp:2040-2050,
:2844, :5317, ...), on the CUDA path against the oracle — device-resident and host-resident chunks."""
import pytest

import gpu_util as gu
import ref_join_tables as rj
import ref_tables as rt
import ref_time_table as tt
import sqlmini
from test_gpu_order_by import run_sorted

pytestmark = pytest.mark.gpu


def test_time_golden_queries_on_the_gpu():
    table = tt.make_table(tt.time_rows())
    dev = gu.DeviceTable(table)
    for sql in tt.TIME_QUERIES:
        unit = sqlmini.parse(sql, table, tt.TIME_NAMES)
        try:
            if unit.unit.num_order_entries:
                run_sorted(unit, table, dev)
            else:
                gu.run_both(unit, table, dev_table=dev)
                gu.run_both(unit, table, device_resident=False)
        except Exception as e:
            raise AssertionError(f"query: {sql}\n{e}") from e


def test_join_golden_queries_on_the_gpu():
    """`test` JOIN `test_inner` (tests/ref_join_tables.py: Select.Joins_* of ExecuteTest.cpp), incl. days-encoded DATE
    columns of the inner table as group key and filter operands."""
    test, inner = rt.make_table(rt.test_rows()), rj.inner_table()
    dev = gu.DeviceTable(test)
    for sql in rj.JOIN_GOLDEN:
        unit = sqlmini.parse(sql, test, rt.TEST_NAMES, inner=(inner, rj.INNER_NAMES))
        try:
            if unit.unit.num_order_entries:
                run_sorted(unit, test, dev)
            else:
                gu.run_both(unit, test, dev_table=dev)
                gu.run_both(unit, test, device_resident=False)
        except Exception as e:
            raise AssertionError(f"query: {sql}\n{e}") from e
    dev_inner = gu.DeviceTable(inner)
    for sql in rj.INNER_GOLDEN:
        unit = sqlmini.parse(sql, inner, rj.INNER_NAMES)
        if unit.unit.num_order_entries:
            run_sorted(unit, inner, dev_inner)
        else:
            gu.run_both(unit, inner, dev_table=dev_inner)


def test_join_single_days_column_rides_in_the_join_table():
    """The only inner column read is a days-encoded DATE spanning < 65534 days: it travels as 16-bit value slots of the
    staged join table — in DAYS, while the chunk stats that bound it are in seconds."""
    import numpy as np
    from heavydb_b200 import abi
    rng = np.random.default_rng(4)
    dim = abi.Table([(abi.kINT, True), (abi.kDATE, False)], encoded_sizes=[0, -4])
    dim.add_host_fragment([np.arange(10, dtype=np.int32), np.array([18000, 18000, 18003, -2**31, 18001, 18000, 18002, 18003, 18000, 18001], dtype=np.int32)])
    fact = abi.Table([(abi.kINT, False), (abi.kBIGINT, True)])
    for _ in range(3):
        fact.add_host_fragment([rng.integers(-1, 13, 5000).astype(np.int32), rng.integers(0, 100, 5000).astype(np.int64)])
    for sql in ["SELECT d.dday, COUNT(*), SUM(t.v) FROM t JOIN d ON t.fk = d.id GROUP BY d.dday;",
                "SELECT COUNT(*), MIN(d.dday), MAX(d.dday), COUNT(d.dday) FROM t LEFT JOIN d ON t.fk = d.id WHERE d.dday >= 1555286400 OR d.dday IS NULL;"]:
        unit = sqlmini.parse(sql, fact, ["fk", "v"], inner=(dim, ["id", "dday"]))
        gu.run_both(unit, fact)
        gu.run_both(unit, fact, device_resident=False)


@pytest.mark.parametrize("entry_count", [1, 2, 3, 5, 13, 31, 63, 126, 241, 511, 1021])
def test_reduction_ladder_of_gpu_shared_memory_test(entry_count):
    """Tests/GpuSharedMemoryTest.cpp:457-617's ladder (tests/reduce_ladder.py): per-CTA shared-memory tables merged into
    the HBM table on the device, one launch over all fragments and one per fragment, against the oracle's host reduce."""
    import reduce_ladder as rl
    from heavydb_b200 import abi
    for i, step in enumerate(rl.STEPS[:3] if entry_count < 100 else rl.STEPS[3:]):
        num_buffers = [2, 8, 64][i]
        table, _ = rl.ladder_table(entry_count, step, num_buffers, rows_per_buffer=300, seed=entry_count * 100 + step)
        unit = sqlmini.parse(rl.QUERY, table, rl.NAMES)
        gu.run_both(unit, table)                                   # the planner's kernel (shared-memory table)
        gu.run_both(unit, table, device_resident=False)           # streamed from the host, slice by slice
        gu.run_both(unit, table, force_kernel=abi.KERNEL_PERFECT_GLOBAL)   # the HBM-table kernel on the same shape
