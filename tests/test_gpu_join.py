"""INNER hash join on the device: one-to-one table built by b2q_k_join_build, probed inside the scan kernel (JOIN
instantiations of b2q_k_scan), inner-table columns gathered at the matching row — against the oracle, bit-exact."""
import numpy as np
import pytest

import gpu_util as gu
import join_tables as jt
import oracle_lib
import sqlmini
from heavydb_b200 import abi, executor
from test_gpu_order_by import run_sorted

pytestmark = pytest.mark.gpu


def parse(sql, fact, dim):
    return sqlmini.parse(sql, fact, jt.FACT_NAMES, inner=(dim, jt.DIM_NAMES))


@pytest.mark.parametrize("n,frag_rows", [(7, 3), (6000, 1700), (300000, 70000)])
def test_join_queries(n, frag_rows):
    fact = jt.fact_table(n, seed=3 + n, frag_rows=frag_rows)
    dim = jt.dim_table()
    dev = gu.DeviceTable(fact)
    for sql in jt.JOIN_QUERIES + jt.LEFT_JOIN_QUERIES:
        unit = parse(sql, fact, dim)
        try:
            if unit.unit.num_order_entries:
                run_sorted(unit, fact, dev, entry_guess=4000, has_card=True)
                continue
            gu.run_both(unit, fact, entry_guess=4000, has_card=True, dev_table=dev)          # fact table resident in HBM
            gu.run_both(unit, fact, entry_guess=4000, has_card=True, device_resident=False)  # fact table streamed from the host
            p = executor.Executor().plan(unit, fact, max_groups_buffer_entry_guess=4000, has_cardinality_estimation=True)
            if p.query_desc_type == abi.GroupByPerfectHash:
                gu.run_both(unit, fact, force_kernel=abi.KERNEL_PERFECT_GLOBAL, dev_table=dev)
        except Exception as e:
            raise AssertionError(f"query: {sql}\n{e}") from e


def test_join_not_one_to_one_is_refused():
    fact = jt.fact_table(1000, seed=1, frag_rows=400)
    dup = abi.Table([(abi.kINT, True), (abi.kINT, False)])
    dup.add_host_fragment([np.array([1, 2, 2, 3], dtype=np.int32), np.array([5, 6, 7, 8], dtype=np.int32)])
    unit = sqlmini.parse("SELECT COUNT(*) FROM t JOIN d ON t.fk32 = d.id;", fact, jt.FACT_NAMES, inner=(dup, ["id", "a"]))
    with pytest.raises(executor.UnsupportedOnThisPath):
        executor.Executor().executeWorkUnit(0, True, fact, unit, memory_level=abi.CPU_LEVEL)
    with pytest.raises(oracle_lib.OracleError):
        oracle_lib.execute(unit, fact)


def test_join_empty_dimension():
    fact = jt.fact_table(500, seed=1, frag_rows=200)
    empty_dim = abi.Table([(ty, nn) for _, ty, nn in jt.DIM_COLS])
    empty_dim.add_host_fragment([np.zeros(0, dtype=abi.NUMPY_OF[ty]) for _, ty, _ in jt.DIM_COLS])
    rs, _ = gu.run_both(parse("SELECT COUNT(*), SUM(t.v) FROM t JOIN d ON t.fk32 = d.id32;", fact, empty_dim), fact)
    assert rs.rows() == [(0, None)]
