"""CPU: planner parity (product vs oracle) and oracle-vs-numpy checks for ENCODING FIXED columns and the deleted-rows column."""
import numpy as np
import pytest

import enc_tables as et
import oracle_lib
from heavydb_b200 import abi, executor, sqlmini


@pytest.mark.parametrize("sql", et.ENC_QUERIES)
def test_plan_parity_with_encodings(sql):
    table = et.enc_table(3000, seed=1, frag_rows=700)
    unit = sqlmini.parse(sql, table, et.ENC_NAMES)
    want = oracle_lib.plan(unit, table, entry_guess=8000, has_card=True).as_dict()
    got = executor.Executor().plan(unit, table, max_groups_buffer_entry_guess=8000, has_cardinality_estimation=True).as_dict()
    assert got == want


def test_oracle_decodes_fixed_encoding_and_skips_deleted():
    table = et.enc_table(5000, seed=2, frag_rows=1000)
    cols = [np.concatenate([f.host_cols[c] for f in table.fragments]) for c in range(len(et.ENC_COLS))]
    live = cols[7] <= 0
    a = cols[2].astype(np.int64)            # BIGINT ENCODING FIXED(8): physical NULL is -128
    nn = live & (a != -128)
    unit = sqlmini.parse("SELECT COUNT(*), COUNT(a_i64_f8), SUM(a_i64_f8), MIN(a_i64_f8), MAX(a_i64_f8) FROM e;", table, et.ENC_NAMES)
    got = oracle_lib.execute(unit, table).rows()[0]
    assert got == (int(live.sum()), int(nn.sum()), int(a[nn].sum()), int(a[nn].min()), int(a[nn].max()))
    # NULL key group: physical -32768 reads back as the logical INT NULL (None)
    unit = sqlmini.parse("SELECT k_i32_f16, COUNT(*) FROM e GROUP BY k_i32_f16;", table, et.ENC_NAMES)
    rows = dict(oracle_lib.execute(unit, table).rows())
    k = cols[0]
    assert rows[None] == int((live & (k == -2**15)).sum())
    assert rows[5] == int((live & (k == 5)).sum())
    # filter_on_deleted_column = false sees every row
    oracle_lib.lib().oracle_set_filter_on_deleted_column(0)
    try:
        unit = sqlmini.parse("SELECT COUNT(*) FROM e;", table, et.ENC_NAMES)
        assert oracle_lib.execute(unit, table).rows() == [(5000,)]
    finally:
        oracle_lib.lib().oracle_set_filter_on_deleted_column(1)


def test_bad_encodings_are_rejected():
    t = abi.Table([(abi.kDOUBLE, True)], encoded_sizes=[4])
    t.add_host_fragment([np.zeros(4, dtype=np.int32)])
    b = abi.UnitBuilder(t)
    b.target(b.agg(abi.kCOUNT))
    with pytest.raises(executor.UnsupportedOnThisPath):
        executor.Executor().plan(b.build(), t)
    t = abi.Table([(abi.kBOOLEAN, True)])          # BOOLEAN that is not the deleted column
    t.add_host_fragment([np.zeros(4, dtype=np.int8)])
    b = abi.UnitBuilder(t)
    b.target(b.agg(abi.kCOUNT))
    with pytest.raises(executor.UnsupportedOnThisPath):
        executor.Executor().plan(b.build(), t)
