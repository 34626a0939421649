"""Cardinality-estimation query on the device (ACC_NDV in the non-grouped scan kernel): the linear-counting bitmap must
equal the oracle's bit for bit, and the whole baseline-hash flow — CardinalityEstimationRequired -> estimator unit ->
2 x NDV entries -> the real query — runs on this path."""
import numpy as np
import pytest

import gpu_util as gu
import join_tables as jt
import oracle_lib
import sqlmini
from heavydb_b200 import abi, executor
from test_gpu_parity import RAND_NAMES, random_table
from test_ndv_estimator import EST_CASES, estimator_unit

pytestmark = pytest.mark.gpu


def check(unit, table, dev):
    ex = executor.Executor()
    ref = oracle_lib.execute(unit, table, num_threads=4)
    for t, lvl in ((dev.table, abi.GPU_LEVEL), (table, abi.CPU_LEVEL)):
        rs = ex.executeWorkUnit(1, True, t, unit, memory_level=lvl)
        assert rs.getQueryMemDesc().as_dict() == ref.plan.as_dict()
        assert np.array_equal(rs.getHostEstimatorBuffer(), ref.buffer().view(np.uint8))
        assert rs.getNDVEstimator() == ref.ndv_estimator()
        assert rs.rowCount() == 0 and rs.rows() == []
    return rs


@pytest.mark.parametrize("n,frag_rows", [(3, 2), (60000, 13000), (400000, 100000)])
def test_estimator_bitmaps(n, frag_rows):
    table = random_table(n, seed=4 + n, frag_rows=frag_rows)
    dev = gu.DeviceTable(table)
    for cols in EST_CASES:
        check(estimator_unit(table, RAND_NAMES, cols), table, dev)
    check(estimator_unit(table, RAND_NAMES, ["k32"], filt=("nn32", abi.kLT, 100)), table, dev)


def test_estimator_on_encoded_and_joined_columns():
    import enc_tables as et
    import str_tables as stt
    t = et.enc_table(30000, seed=3, frag_rows=8000)
    check(estimator_unit(t, et.ENC_NAMES, ["k_i32_f16", "a_i64_f8"]), t, gu.DeviceTable(t))       # FIXED(16) / FIXED(8), deleted rows
    s = stt.str_table(30000, seed=3, frag_rows=8000)
    check(estimator_unit(s, stt.STR_NAMES, ["s8", "str", "ts"]), s, gu.DeviceTable(s))              # DICT(8) ids, NULLs
    fact, dim = jt.fact_table(50000, seed=2, frag_rows=16000), jt.dim_table()
    check(estimator_unit(fact, jt.FACT_NAMES, [(jt.DIM_NAMES.index("big"), 1), "x"], inner=(dim, 0, 0)), fact, gu.DeviceTable(fact))


def test_large_estimator():
    table = random_table(20000, seed=9, frag_rows=20000)
    unit = estimator_unit(table, RAND_NAMES, ["big", "k64"], large=True)       # LargeNDVEstimator: 256 MiB bitmap
    rs = executor.Executor().executeWorkUnit(1, True, table, unit, memory_level=abi.CPU_LEVEL)
    assert rs.getQueryMemDesc().buffer_size == 256 << 20
    assert abs(rs.getNDVEstimator() - 20000) <= 2


def test_baseline_flow_stays_on_the_path():
    """RelAlgExecutor.cpp:4194-4233: CardinalityEstimationRequired -> getNDVEstimation (createNdvExecutionUnit) ->
    max_groups_buffer_entry_guess = 2 x NDV -> executeWorkUnit again, has_cardinality_estimation = true."""
    table = random_table(120000, seed=21, frag_rows=30000)
    dev = gu.DeviceTable(table)
    ex = executor.Executor()
    unit = sqlmini.parse("SELECT sparse, COUNT(*), SUM(a64) FROM r WHERE nn32 < 250 GROUP BY sparse;", table, RAND_NAMES)
    with pytest.raises(executor.CardinalityEstimationRequired):
        ex.executeWorkUnit(0, True, dev.table, unit, memory_level=abi.GPU_LEVEL)
    ndv_unit = estimator_unit(table, RAND_NAMES, ["sparse"], filt=("nn32", abi.kLT, 250))
    ndv = ex.executeWorkUnit(1, True, dev.table, ndv_unit, memory_level=abi.GPU_LEVEL).getNDVEstimator()
    assert abs(ndv - 2000) <= 3
    gu.run_both(unit, table, entry_guess=2 * ndv, has_card=True, dev_table=dev)
