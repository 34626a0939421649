"""`-m gpu`: the reference's own golden query strings, verbatim (tests/golden/executetest_harvest.json — see
tests/test_oracle_golden_harvest.py), through the CUDA path: rows, plan and raw output buffers against the oracle, host- and
device-resident fragments of the golden table (fragment_size = 2, ExecuteTest.cpp:30063-30115)."""
import pytest

import gpu_util as gu
import ref_full_table as ft
import sqlmini
from test_oracle_golden_harvest import QUERIES

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def golden():
    table = ft.make_table(ft.full_rows())
    return table, gu.DeviceTable(table)


@pytest.mark.parametrize("sql", QUERIES)
def test_reference_query_verbatim_on_the_gpu(golden, sql):
    table, dev = golden
    unit = sqlmini.parse(sql, table, ft.FULL_NAMES, dicts=ft.DICTS)
    if unit.unit.num_order_entries or unit.unit.has_limit or unit.unit.offset:
        from test_gpu_order_by import run_sorted     # rows in order + the compact buffer against the oracle's permutation
        run_sorted(unit, table, dev, entry_guess=48, has_card=True)
        return
    gu.run_both(unit, table, entry_guess=48, has_card=True, dev_table=dev)
    gu.run_both(unit, table, entry_guess=48, has_card=True, device_resident=False)
