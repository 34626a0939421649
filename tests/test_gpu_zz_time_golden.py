"""The reference's golden `test` table, TIME-family / encoded columns (tests/ref_time_table.py: ExecuteTest.cpp:2040-2050,
:2844, :5317, ...), on the CUDA path against the oracle — device-resident and host-resident chunks."""
import pytest

import gpu_util as gu
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
